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


@pytest.mark.parametrize("scale,shift", [(0.25, 0.05), (0.5, 0.1), (1.0, 0.3), (2.0, -0.3)])
def test_independent_anchor_exact_ot_at_several_diameters(scale, shift):
    """Independent anchor for the unpinned restatement: at blur = 0.05 (eps = 0.0025, far below the
    squared point spacing) the debiased divergence must approach the EXACT optimal-transport cost 0.5 * W2^2 computed by
    scipy's Hungarian solver, at four different cloud diameters (different eps-ladder lengths), and the gradient must be
    the displacement to the assigned partner: dS/dx_i -> (x_i - y_sigma(i)) / N."""
    from scipy.optimize import linear_sum_assignment
    g = torch.Generator().manual_seed(11)
    x = torch.randn(196, 384, generator=g, dtype=torch.float64) * scale
    y = torch.randn(196, 384, generator=g, dtype=torch.float64) * scale + shift
    d = max_diameter(x, y)
    n_eps = len(epsilon_schedule(2, d, 0.05, 0.5))
    C = 0.5 * torch.cdist(x, y).pow(2).numpy()
    r, c = linear_sum_assignment(C)
    w2 = C[r, c].mean()
    xg = x.clone().requires_grad_(True)
    s = sinkhorn_divergence(xg, y)
    s.backward()
    assert abs(s.item() - w2) <= 0.01 * w2, (scale, d, n_eps, s.item(), w2)
    # Gradient: close to that of the exact cost, (x_i - y_sigma(i)) / N (not equal: geomloss stops the eps ladder after one
    # step per eps, so ~17 % of the plan rows are not one-hot yet), and a step along it must lower the EXACT OT cost.
    g_exact = (x - y[torch.as_tensor(c)]) / 196.0
    cos = torch.nn.functional.cosine_similarity(xg.grad.reshape(1, -1), g_exact.reshape(1, -1)).item()
    assert cos > 0.75, (scale, cos)
    x2 = (x - 0.25 * 196.0 * xg.grad).numpy()
    C2 = 0.5 * torch.cdist(torch.from_numpy(x2), y).pow(2).numpy()
    r2, c2 = linear_sum_assignment(C2)
    assert C2[r2, c2].mean() < 0.8 * w2, (scale, C2[r2, c2].mean(), w2)


def test_dual_objective_bounds_the_exact_cost():
    """The entropic dual potentials (f_ba, g_ab) of OT_eps(x, y) are feasible up to eps for the unregularised dual:
    f_i + g_j <= C_ij + O(eps log N); hence mean(f) + mean(g) <= exact cost + O(eps log N) (weak duality), and at
    eps = 0.0025 it is tight to a fraction of a percent."""
    from scipy.optimize import linear_sum_assignment
    from oracle.sinkhorn import _half_sqdist, _softmin
    g = torch.Generator().manual_seed(4)
    x = torch.randn(64, 32, generator=g, dtype=torch.float64) * 0.7
    y = torch.randn(64, 32, generator=g, dtype=torch.float64) * 0.7 + 0.2
    C = _half_sqdist(x, y)
    N = 64
    logw = torch.full((N,), -float(np.log(N)), dtype=torch.float64)
    f = torch.zeros(N, dtype=torch.float64)
    gq = torch.zeros(N, dtype=torch.float64)
    for eps in epsilon_schedule(2, max_diameter(x, y), 0.05, 0.5) + [0.0025] * 200:   # plain alternating Sinkhorn to convergence
        f = _softmin(eps, C, logw + gq / eps)
        gq = _softmin(eps, C.t(), logw + f / eps)
    r, c = linear_sum_assignment(C.numpy())
    exact = C.numpy()[r, c].mean()
    dual = (f.mean() + gq.mean()).item()
    slack = 0.0025 * np.log(N)
    assert dual <= exact + slack
    assert abs(dual - exact) <= 0.01 * exact + slack
    assert float((f.view(-1, 1) + gq.view(1, -1) - C).max()) <= 2 * slack   # approximate dual feasibility
