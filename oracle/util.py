"""TEST INFRASTRUCTURE ONLY — helpers shared by oracle/make_golden.py and tests/."""
from __future__ import annotations

import numpy as np
import torch

GRAD_STRIDE = 97


def digest_stride(numel: int) -> int:
    """Sampling stride of `digest`: 97 for small tensors, ~1500 samples for large ones (kept odd)."""
    return max(GRAD_STRIDE, (numel // 1500) | 1)


def digest(t: torch.Tensor) -> np.ndarray:
    """[norm, sum, abs-sum] + strided sample of a tensor (fp64)."""
    f = t.detach().double().cpu().reshape(-1)
    head = torch.stack([f.norm(), f.sum(), f.abs().sum()])
    return torch.cat([head, f[::digest_stride(f.numel())]]).numpy()


def rel_err(a, b) -> float:
    """Norm-wise relative error |a-b| / max(|b|, tiny) in fp64."""
    a = torch.as_tensor(np.asarray(a)).double().reshape(-1) if not torch.is_tensor(a) else a.detach().double().cpu().reshape(-1)
    b = torch.as_tensor(np.asarray(b)).double().reshape(-1) if not torch.is_tensor(b) else b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / max(float(b.norm()), 1e-300))
