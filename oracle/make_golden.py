"""TEST INFRASTRUCTURE ONLY — golden-vector generator.

Imports the UNMODIFIED reference (`/root/reference/model/{loss,misc,models}.py`)
through the shim packages in oracle/shims/ (timm and geomloss are not installed
in this image) and records, for seeded synthetic inputs,
what the reference's own `DistillationLoss.forward`, `random_masking` and
`saliency_masking` return (fp32 and fp64) plus gradient digests.  The result is
committed as tests/golden/reference_v1.npz; `/root/reference` does not exist on
the GPU box, so nothing else may read it.

    python oracle/make_golden.py          # run in the authoring container only

Inputs are regenerated from seeds by deltakd_b200/synth.py at test time; the
file stores input checksums to detect RNG drift.  Note `wasskd/sinkhorn` goes
through oracle/sinkhorn.py on both sides (geomloss absent -> parity unpinned),
so that case pins only the wrapping arithmetic of loss.py:200-226.
"""
from __future__ import annotations

import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle", "shims"), "/root/reference", ROOT]

import model.loss as ref_loss  # noqa: E402
import model.misc as ref_misc  # noqa: E402
import model.models as ref_models  # noqa: E402

from deltakd_b200 import synth  # noqa: E402

from oracle.util import digest  # noqa: E402


def build_models(dtype_name: str, args, dtype):
    """Reference head attachment (models.py:59-178) on replay stand-ins."""
    def fake_create_model(name, pretrained=False, **kw):
        return synth.FeatureReplayModel(384 if "small" in name else 192)

    ref_models.timm.create_model = fake_create_model
    args.distillation_type = dtype_name
    args.dataset = "imagenet-1k"
    torch.manual_seed(0)
    student_name = "deit_tiny_distilled_patch16_224" if dtype_name in ("soft", "hard") else "deit_tiny_patch16_224"
    teacher, student = ref_models.load_teacher_student_model(
        "deit_small_distilled_patch16_224", student_name, args=args)
    if hasattr(student, "mask_token"):  # zeros by default (models.py:147); make it non-trivial
        with torch.no_grad():
            student.mask_token.copy_(torch.randn(student.mask_token.shape, generator=torch.Generator().manual_seed(7)) * 0.1)
    return teacher.to(dtype), student.to(dtype)


def head_state(student) -> dict:
    return {k: v for k, v in student.state_dict().items() if not k.startswith("blocks")}


def run_case(name, dtype_name, args, B, C, base_int_labels=False, alpha=0.1, tau=3.0, seed=1234,
             feat_kw=None, rng_seed=4321, out=None):
    feat_kw = feat_kw or {}
    for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        a = SimpleNamespace(**vars(args))
        teacher, student = build_models(dtype_name, a, dt)
        outputs, outputs_kd, t_logits, labels = synth.make_logits(B, C, seed, int_labels=base_int_labels)
        outputs, outputs_kd, t_logits = outputs.to(dt), outputs_kd.to(dt), t_logits.to(dt)
        if not base_int_labels:
            labels = labels.to(dt)
        needs_feats = dtype_name not in ("soft", "hard", "none")
        s_feats = t_feats = None
        if needs_feats:
            s_feats, t_feats = synth.make_features(B, seed, **feat_kw)
            s_feats = [f.to(dt).requires_grad_(True) for f in s_feats]
            t_feats = [f.to(dt) for f in t_feats]
        teacher.set_outputs(t_logits, t_feats)
        outputs.requires_grad_(True)
        outputs_kd.requires_grad_(True)
        a.mixup, a.cutmix = (0.0, 0.0) if base_int_labels else (0.8, 1.0)
        base = ref_loss.call_base_loss(a)
        crit = ref_loss.DistillationLoss(base, teacher, dtype_name, alpha, tau)
        model_out = (outputs, outputs_kd) if dtype_name in ("soft", "hard") else outputs
        inputs = torch.zeros(B, 3, 2, 2, dtype=dt)
        torch.manual_seed(rng_seed)  # the reference draws torch.rand(N, L) inside random_masking
        loss = crit(inputs, model_out, student, s_feats, labels, a)
        loss.backward()
        out[f"{name}/{tag}/loss"] = np.float64(loss.item())
        out[f"{name}/{tag}/g_outputs"] = digest(outputs.grad)
        if outputs_kd.grad is not None:
            out[f"{name}/{tag}/g_outputs_kd"] = digest(outputs_kd.grad)
        if needs_feats:
            for i, f in enumerate(s_feats):
                if f.grad is not None and f.grad.abs().sum() > 0:
                    out[f"{name}/{tag}/g_sfeat{i}"] = digest(f.grad)
            for k, p in student.named_parameters():
                if p.grad is not None and not k.startswith("blocks"):
                    out[f"{name}/{tag}/g_head/{k}"] = digest(p.grad)
        if tag == "f32":
            out[f"{name}/in_digest"] = np.concatenate(
                [digest(outputs)[:3], digest(labels.double())[:3]]
                + ([digest(s_feats[0])[:3], digest(t_feats[11])[:3]] if needs_feats else []))
            if dtype_name == "lrkd":  # record the reference's SVD column orientation: sign of the largest |V| entry
                for j, ti in enumerate((0, 1, 11)):
                    T = t_feats[ti][:, 2:].reshape(-1, 384)
                    U, S, Vh = torch.linalg.svd(T, full_matrices=False)
                    V = Vh[: a.lrkd_rank]
                    idx = V.abs().argmax(dim=1)
                    out[f"{name}/svd_sign{j}"] = torch.sign(V[torch.arange(V.shape[0]), idx]).numpy()
                    out[f"{name}/svd_S{j}"] = S[: a.lrkd_rank].double().numpy()
    print("  case", name, "loss f32", out[f"{name}/f32/loss"], "f64", out[f"{name}/f64/loss"])


def main():
    out = {}
    A = lambda **kw: SimpleNamespace(**{**dict(  # loss-relevant defaults of tools/train.py:103-136,157-186
        lrkd_rank=32, lrkd_alpha=0.1, lrkd_beta=0.1, lrkd_gamma=0.1, saliency_method=1,
        saliency_mask_ratio=0.5, wasskd_type="l1", mgd_alpha=7e-5, mgd_mask_ratio=0.5,
        mixup=0.8, cutmix=1.0, cutmix_minmax=None, smoothing=0.1, current_epoch=0), **kw})

    # --- logit losses (config 1: B=8, C=1000; and the script's real C=100)
    run_case("none_b8_c1000", "none", A(), 8, 1000, out=out)
    run_case("soft_b8_c1000", "soft", A(), 8, 1000, out=out)
    run_case("soft_b8_c100", "soft", A(), 8, 100, out=out)
    run_case("soft_b5_c1000_tau1", "soft", A(), 5, 1000, tau=1.0, alpha=0.5, out=out)
    run_case("soft_b8_c1000_intlabels", "soft", A(), 8, 1000, base_int_labels=True, out=out)
    run_case("hard_b8_c1000", "hard", A(), 8, 1000, out=out)
    run_case("hard_b8_c100_intlabels", "hard", A(), 8, 100, base_int_labels=True, out=out)
    # --- feature losses, B=2 at the real token / channel shapes
    run_case("curkd_ep0", "curkd", A(current_epoch=0), 2, 100, out=out)
    run_case("curkd_ep120", "curkd", A(current_epoch=120), 2, 100, out=out)
    run_case("curkd_ep200", "curkd", A(current_epoch=200), 2, 100, out=out)
    run_case("mgd_r05", "mgd", A(), 2, 100, out=out)
    run_case("mgd_r03", "mgd", A(mgd_mask_ratio=0.3, mgd_alpha=2e-5), 3, 100, out=out)
    for m in (1, 2, 3):
        run_case(f"salmgd_m{m}", "saliency_mgd", A(saliency_method=m), 2, 100, out=out)
    run_case("salmgd_m1_r07", "saliency_mgd", A(saliency_method=1, saliency_mask_ratio=0.7), 2, 100, out=out)
    run_case("lrkd_r32", "lrkd", A(), 3, 100, out=out)
    run_case("lrkd_r64", "lrkd", A(lrkd_rank=64, lrkd_alpha=0.2, lrkd_beta=0.2, lrkd_gamma=0.2), 3, 100, out=out)
    run_case("wass_l1", "wasskd", A(), 2, 100, out=out)
    run_case("wass_sinkhorn", "wasskd", A(wasskd_type="sinkhorn"), 2, 100,
             feat_kw=dict(scale=0.5, t_shift=0.1), out=out)
    run_case("vitkd", "vitkd", A(), 2, 100, out=out)

    # --- random_masking (misc.py:5-32): integer outputs, bit-exact
    for ratio in (0.5, 0.3, 0.75, 0.0):
        x = torch.randn(4, 196, 8, generator=torch.Generator().manual_seed(5))
        torch.manual_seed(99)
        x_keep, mask, ids_restore, ids_masked = ref_misc.random_masking(x, ratio)
        torch.manual_seed(99)
        noise = torch.rand(4, 196)
        tag = f"random_masking/r{ratio}"
        out[f"{tag}/noise"] = noise.numpy()
        out[f"{tag}/mask"] = mask.numpy()
        out[f"{tag}/ids_restore"] = ids_restore.numpy()
        out[f"{tag}/ids_masked"] = ids_masked.numpy()
        out[f"{tag}/x_keep_digest"] = digest(x_keep)
    # ties: quantised noise so equal keys straddle the keep boundary
    x = torch.randn(3, 196, 8, generator=torch.Generator().manual_seed(5))

    class _Quant:
        def __enter__(self):
            self.orig = torch.rand
            torch.rand = lambda *s, **k: torch.floor(self.orig(*s, **k) * 16) / 16
        def __exit__(self, *e):
            torch.rand = self.orig
    torch.manual_seed(99)
    with _Quant():
        x_keep, mask, ids_restore, ids_masked = ref_misc.random_masking(x, 0.5)
    torch.manual_seed(99)
    noise = torch.floor(torch.rand(3, 196) * 16) / 16
    out["random_masking/ties/noise"] = noise.numpy()
    out["random_masking/ties/mask"] = mask.numpy()
    out["random_masking/ties/ids_restore"] = ids_restore.numpy()

    # --- saliency_masking (misc.py:38-165): scores, mask, ids_restore per method
    for m in (1, 2, 3):
        a = A(saliency_method=m)
        teacher, student = build_models("saliency_mgd", a, torch.float32)
        _, t_feats = synth.make_features(3, 77, layers=[11])
        sfeat = torch.randn(3, 196, 384, generator=torch.Generator().manual_seed(6))
        with torch.no_grad():
            x_keep, mask, ids_restore = ref_misc.saliency_masking(student, t_feats[11], sfeat, 0.5, m)
            if m == 1:
                score = student.saliency_attn(t_feats[11][:, 2:])
            elif m == 2:
                tf = torch.cat([t_feats[11][:, :1], t_feats[11][:, 2:]], 1)
                qk = student.saliency_attn.qk(tf)
                q, k = torch.chunk(qk, 2, dim=-1)
                q = q.reshape(3, 197, 8, 48).permute(0, 2, 1, 3)
                k = k.reshape(3, 197, 8, 48).permute(0, 2, 1, 3)
                score = ((q[:, :, 0:1] @ k.transpose(-2, -1)) * 48 ** -0.5).softmax(-1).mean(1).squeeze(1)[:, 1:]
            else:
                score = student.saliency_attn(t_feats[11][:, :1], t_feats[11][:, 2:]).squeeze(1)
        out[f"saliency_masking/m{m}/score"] = score.numpy()
        out[f"saliency_masking/m{m}/mask"] = mask.numpy()
        out[f"saliency_masking/m{m}/ids_restore"] = ids_restore.numpy()
        out[f"saliency_masking/m{m}/x_keep_digest"] = digest(x_keep)

    path = os.path.join(ROOT, "tests", "golden", "reference_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB,", len(out), "arrays")


if __name__ == "__main__":
    main()
