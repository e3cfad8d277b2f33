"""CPU: the Sinkhorn restatement (oracle/sinkhorn.py; geomloss is absent -> PARITY UNPINNED) is anchored on the
properties of the debiased divergence and on scipy's exact optimal assignment."""
import numpy as np
import pytest
import torch

from oracle.sinkhorn import epsilon_schedule, max_diameter, sinkhorn_divergence


def _clouds(n=48, d=16, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, d, generator=g, dtype=torch.float64) * 0.5, torch.randn(n, d, generator=g, dtype=torch.float64) * 0.5 + 0.1


def test_divergence_of_identical_clouds_is_zero():
    x, _ = _clouds()
    assert abs(sinkhorn_divergence(x, x.clone()).item()) < 1e-12


def test_symmetry_and_positivity():
    x, y = _clouds()
    a, b = sinkhorn_divergence(x, y).item(), sinkhorn_divergence(y, x).item()
    assert a > 0 and abs(a - b) <= 1e-9 * a


def test_converges_to_half_squared_w2():
    from scipy.optimize import linear_sum_assignment
    x, y = _clouds(196, 384, 3)   # the shape the reference feeds (196 tokens in R^384)
    C = 0.5 * torch.cdist(x, y).pow(2).numpy()
    r, c = linear_sum_assignment(C)
    w2 = C[r, c].mean()
    s = sinkhorn_divergence(x, y).item()
    assert abs(s - w2) <= 0.01 * w2, (s, w2)   # scaling=0.5 stops early: ~0.2 % below the exact value


def test_epsilon_schedule_shape():
    x, y = _clouds(196, 384, 1)
    d = max_diameter(x, y)
    eps = epsilon_schedule(2, d, 0.05, 0.5)
    assert eps[0] == d ** 2 and eps[-1] == 0.05 ** 2
    assert len(eps) == 2 + int(np.ceil((2 * np.log(0.05) - 2 * np.log(d)) / (2 * np.log(0.5))))
    assert all(eps[i] > eps[i + 1] for i in range(1, len(eps) - 1))


def test_gradient_is_finite_and_translation_consistent():
    x, y = _clouds(32, 8, 5)
    x.requires_grad_(True)
    sinkhorn_divergence(x, y).backward()
    assert torch.isfinite(x.grad).all()
    # moving x towards y decreases the divergence
    with torch.no_grad():
        x2 = x - 0.1 * x.grad / x.grad.norm()
    assert sinkhorn_divergence(x2, y).item() < sinkhorn_divergence(x.detach(), y).item()
