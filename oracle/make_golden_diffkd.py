"""TEST INFRASTRUCTURE ONLY — golden vectors of the reference's DiffKD branch (model/loss.py:105-155).

Same method as oracle/make_golden.py (unmodified reference imported through the shims), kept in its own file and
fixture so that reference_v1.npz stays byte-stable.  The branch draws `torch.randint` (diffusion step per sample) and
`torch.randn_like` (noise per layer) from torch's global generator; the draws are recorded and replayed so that the
fixture is independent of RNG streams, and `denoise_fn` runs in eval mode (its Dropout mask is another RNG draw).

    python oracle/make_golden_diffkd.py       # authoring container only; writes tests/golden/reference_diffkd_v1.npz
"""
from __future__ import annotations

import os
import sys
from types import SimpleNamespace
from unittest import mock

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle", "shims"), "/root/reference", ROOT]

import model.loss as ref_loss  # noqa: E402
from deltakd_b200 import synth  # noqa: E402
from oracle.make_golden import build_models  # noqa: E402
from oracle.util import digest  # noqa: E402

B, C, SEED = 3, 100, 1234
DIFF_T = [5, 2, 7]          # one step below T/2 and two above; no t = 0 (sigma = 0 makes w_t = 1e8)


def draws():
    g = torch.Generator().manual_seed(2024)
    return torch.tensor(DIFF_T), [torch.randn(B, 196, 384, generator=g) for _ in range(3)]


def main():
    out = {}
    A = SimpleNamespace(lrkd_rank=32, lrkd_alpha=0.1, lrkd_beta=0.1, lrkd_gamma=0.1, saliency_method=1, saliency_mask_ratio=0.5,
                        wasskd_type="l1", mgd_alpha=7e-5, mgd_mask_ratio=0.5, mixup=0.8, cutmix=1.0, cutmix_minmax=None,
                        smoothing=0.1, current_epoch=0)
    # float32 only: the reference hard-codes `t.float()` into denoise_fn (models.py:118), so the branch cannot run in fp64;
    # the fp64 side of the parity tests is the oracle restatement, itself pinned to this fp32 fixture.
    for dt, tag in ((torch.float32, "f32"),):
        a = SimpleNamespace(**vars(A))
        teacher, student = build_models("diffkd", a, dt)
        student.denoise_fn.eval()
        outputs, _, t_logits, labels = synth.make_logits(B, C, SEED)
        outputs, t_logits, labels = outputs.to(dt).requires_grad_(True), t_logits.to(dt), labels.to(dt)
        s_feats, t_feats = synth.make_features(B, SEED)
        s_feats = [f.to(dt).requires_grad_(True) for f in s_feats]
        t_feats = [f.to(dt) for f in t_feats]
        teacher.set_outputs(t_logits, t_feats)
        crit = ref_loss.DistillationLoss(ref_loss.call_base_loss(a), teacher, "diffkd", 0.1, 3.0)
        t_draw, noises = draws()
        it = iter(noises)
        with mock.patch("torch.randint", side_effect=lambda *x, **k: t_draw.clone()), \
                mock.patch("torch.randn_like", side_effect=lambda x, **k: next(it).to(x.dtype)):
            loss = crit(torch.zeros(B, 3, 2, 2, dtype=dt), outputs, student, s_feats, labels, a)
        loss.backward()
        out[f"diffkd/{tag}/loss"] = np.float64(loss.item())
        out[f"diffkd/{tag}/g_outputs"] = digest(outputs.grad)
        for i, f in enumerate(s_feats):
            if f.grad is not None and f.grad.abs().sum() > 0:
                out[f"diffkd/{tag}/g_sfeat{i}"] = digest(f.grad)
        for k, p in student.named_parameters():
            if p.grad is not None and not k.startswith("blocks"):
                out[f"diffkd/{tag}/g_head/{k}"] = digest(p.grad)
        print(tag, "loss", loss.item(), "grads", sorted(k for k in out if k.startswith(f"diffkd/{tag}/g_"))[:4], "...")
    path = os.path.join(ROOT, "tests", "golden", "reference_diffkd_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays")


if __name__ == "__main__":
    main()
