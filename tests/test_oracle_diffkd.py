"""CPU: the oracle's DiffKD restatement (oracle/losses.py:diffkd) against the golden produced by the unmodified
reference branch (model/loss.py:105-155) with its RNG draws replayed (oracle/make_golden_diffkd.py)."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from deltakd_b200 import heads as H
from deltakd_b200 import synth
from oracle import losses as O
from oracle.util import digest, rel_err

GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_diffkd_v1.npz")
B, C, SEED = 3, 100, 1234


def diffkd_case(dtype=torch.float32, device="cpu"):
    """The inputs oracle/make_golden_diffkd.py fed to the reference (same seeds, same head creation order)."""
    args = synth.default_args(distillation_type="diffkd")
    teacher, student = synth.FeatureReplayModel(384), synth.FeatureReplayModel(192)
    torch.manual_seed(0)
    H.attach_distillation_heads(student, teacher, args, "deit_tiny_patch16_224")
    student = student.to(dtype).to(device)
    student.denoise_fn.eval()
    outputs, _, t_logits, labels = synth.make_logits(B, C, SEED)
    s_feats, t_feats = synth.make_features(B, SEED)
    g = torch.Generator().manual_seed(2024)
    noises = [torch.randn(B, 196, 384, generator=g) for _ in range(3)]
    return SimpleNamespace(
        args=args, teacher=teacher, student=student, outputs=outputs.to(dtype).to(device).requires_grad_(True),
        teacher_logits=t_logits.to(dtype).to(device), labels=labels.to(dtype).to(device),
        s_feats=[f.to(dtype).to(device).requires_grad_(True) for f in s_feats], t_feats=[f.to(dtype).to(device) for f in t_feats],
        t=torch.tensor([5, 2, 7], device=device), noises=[n.to(device) for n in noises])


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_oracle_diffkd_matches_reference(dtype):
    gold = np.load(GOLD)
    c = diffkd_case(dtype)
    heads = H.head_tensors(c.student)
    loss = O.distillation_loss("diffkd", c.outputs, c.labels, c.teacher_logits, c.s_feats, c.t_feats, heads, c.args, 0.1, 3.0,
                               diff_t=c.t, diff_noises=c.noises)
    loss.backward()
    tol = 2e-6 if dtype == torch.float32 else 1e-5    # fp64 oracle vs the reference's fp32 run
    assert abs(loss.item() - float(gold["diffkd/f32/loss"])) <= tol * abs(loss.item())
    gt = 1e-4
    for i in (0, 1, 11):
        assert rel_err(digest(c.s_feats[i].grad), gold[f"diffkd/f32/g_sfeat{i}"]) < gt, i
    n = 0
    for k, p in heads.items():
        key = f"diffkd/f32/g_head/{k}"
        if key in gold.files:
            assert p.grad is not None, k
            assert rel_err(digest(p.grad), gold[key]) < gt, k
            n += 1
    assert n == 14   # 3 align (w, b) + denoise_fn: net.0, net.2, time_embed.0, time_embed.2 (w, b)
