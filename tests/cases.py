"""Shared case table: the same seeded inputs oracle/make_golden.py fed to the reference."""
from __future__ import annotations

from types import SimpleNamespace

import torch

from deltakd_b200 import heads as H
from deltakd_b200 import synth


def A(**kw):
    d = dict(lrkd_rank=32, lrkd_alpha=0.1, lrkd_beta=0.1, lrkd_gamma=0.1, saliency_method=1,
             saliency_mask_ratio=0.5, wasskd_type="l1", mgd_alpha=7e-5, mgd_mask_ratio=0.5,
             mixup=0.8, cutmix=1.0, cutmix_minmax=None, smoothing=0.1, current_epoch=0)
    d.update(kw)
    return SimpleNamespace(**d)


# name -> (distillation_type, args, B, C, kwargs)
CASES = {
    "none_b8_c1000": ("none", A(), 8, 1000, {}),
    "soft_b8_c1000": ("soft", A(), 8, 1000, {}),
    "soft_b8_c100": ("soft", A(), 8, 100, {}),
    "soft_b5_c1000_tau1": ("soft", A(), 5, 1000, dict(tau=1.0, alpha=0.5)),
    "soft_b8_c1000_intlabels": ("soft", A(), 8, 1000, dict(base_int_labels=True)),
    "hard_b8_c1000": ("hard", A(), 8, 1000, {}),
    "hard_b8_c100_intlabels": ("hard", A(), 8, 100, dict(base_int_labels=True)),
    "curkd_ep0": ("curkd", A(current_epoch=0), 2, 100, {}),
    "curkd_ep120": ("curkd", A(current_epoch=120), 2, 100, {}),
    "curkd_ep200": ("curkd", A(current_epoch=200), 2, 100, {}),
    "mgd_r05": ("mgd", A(), 2, 100, {}),
    "mgd_r03": ("mgd", A(mgd_mask_ratio=0.3, mgd_alpha=2e-5), 3, 100, {}),
    "salmgd_m1": ("saliency_mgd", A(saliency_method=1), 2, 100, {}),
    "salmgd_m2": ("saliency_mgd", A(saliency_method=2), 2, 100, {}),
    "salmgd_m3": ("saliency_mgd", A(saliency_method=3), 2, 100, {}),
    "salmgd_m1_r07": ("saliency_mgd", A(saliency_method=1, saliency_mask_ratio=0.7), 2, 100, {}),
    "lrkd_r32": ("lrkd", A(), 3, 100, {}),
    "lrkd_r64": ("lrkd", A(lrkd_rank=64, lrkd_alpha=0.2, lrkd_beta=0.2, lrkd_gamma=0.2), 3, 100, {}),
    "wass_l1": ("wasskd", A(), 2, 100, {}),
    "wass_sinkhorn": ("wasskd", A(wasskd_type="sinkhorn"), 2, 100, dict(feat_kw=dict(scale=0.5, t_shift=0.1))),
    "vitkd": ("vitkd", A(), 2, 100, {}),
}

RNG_SEED = 4321  # torch.manual_seed before the loss call (the reference draws torch.rand inside random_masking)


def build_case(name: str, dtype=torch.float32, device="cpu"):
    """Rebuild the inputs of a golden case.  Returns a namespace with student (heads attached),
    teacher (replaying its outputs), tensors and hyper-parameters."""
    kind, args, B, C, kw = CASES[name]
    args = SimpleNamespace(**vars(args))
    args.distillation_type = kind
    int_labels = kw.get("base_int_labels", False)
    args.mixup, args.cutmix = (0.0, 0.0) if int_labels else (0.8, 1.0)
    teacher = synth.FeatureReplayModel(384)
    student = synth.FeatureReplayModel(192)
    torch.manual_seed(0)
    name_s = "deit_tiny_distilled_patch16_224" if kind in ("soft", "hard") else "deit_tiny_patch16_224"
    H.attach_distillation_heads(student, teacher, args, name_s)
    if hasattr(student, "mask_token"):
        with torch.no_grad():
            student.mask_token.copy_(torch.randn(student.mask_token.shape,
                                                 generator=torch.Generator().manual_seed(7)) * 0.1)
    student = student.to(dtype).to(device)
    outputs, outputs_kd, t_logits, labels = synth.make_logits(B, C, 1234, int_labels=int_labels)
    outputs, outputs_kd, t_logits = (t.to(dtype).to(device) for t in (outputs, outputs_kd, t_logits))
    labels = labels.to(device) if int_labels else labels.to(dtype).to(device)
    needs_feats = kind not in ("soft", "hard", "none")
    s_feats = t_feats = None
    if needs_feats:
        s_feats, t_feats = synth.make_features(B, 1234, **kw.get("feat_kw", {}))
        s_feats = [f.to(dtype).to(device).requires_grad_(True) for f in s_feats]
        t_feats = [f.to(dtype).to(device) for f in t_feats]
    teacher.set_outputs(t_logits, t_feats)
    outputs.requires_grad_(True)
    outputs_kd.requires_grad_(True)
    torch.manual_seed(RNG_SEED)
    noise = torch.rand(B, 196)  # what the reference's random_masking draws first on CPU
    return SimpleNamespace(name=name, kind=kind, args=args, B=B, C=C, alpha=kw.get("alpha", 0.1),
                           tau=kw.get("tau", 3.0), int_labels=int_labels, teacher=teacher, student=student,
                           outputs=outputs, outputs_kd=outputs_kd, teacher_logits=t_logits, labels=labels,
                           s_feats=s_feats, t_feats=t_feats, noise=noise.to(device), needs_feats=needs_feats)
