"""TEST INFRASTRUCTURE ONLY — CPU restatement of the Sinkhorn divergence that
`geomloss.SamplesLoss("sinkhorn", blur=0.05)` evaluates for the reference's
WassKD-sinkhorn branch (call sites /root/reference/model/loss.py:8,202,221).

PARITY UNPINNED: `geomloss` is an unlisted, unpinned third-party dependency of
the reference (absent from /root/reference/requirements.txt:1-38 and
setup.py:9-20) and is not installed in this image; there is no network.  This
file restates the published algorithm of upstream geomloss 0.2.x
(`sinkhorn_samples.sinkhorn_tensorized`, `sinkhorn_divergence.sinkhorn_loop`,
`epsilon_schedule`, `max_diameter`, `sinkhorn_cost`) for the defaults the
reference uses: p=2, blur=0.05, scaling=0.5, reach=None (balanced), debias=True,
uniform weights, tensorized backend (N*M <= 5000**2).  The reference holds no
test or golden vector for this term, so the restatement is anchored only on
its own mathematical properties (S(x,x)=0, symmetry, S -> 0.5*W2^2), which
tests/test_oracle_sinkhorn.py checks against scipy's exact assignment.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs
may import this module.
"""
from __future__ import annotations

import numpy as np
import torch


def max_diameter(x: torch.Tensor, y: torch.Tensor) -> float:
    """Length of the diagonal of the joint bounding box of x and y ([N,D],[M,D])."""
    mins = torch.minimum(x.min(dim=0)[0], y.min(dim=0)[0])
    maxs = torch.maximum(x.max(dim=0)[0], y.max(dim=0)[0])
    return (maxs - mins).norm().item()


def epsilon_schedule(p: float, diameter: float, blur: float, scaling: float):
    """[diam^p] + geometric ladder exp(arange(p ln diam, p ln blur, p ln scaling)) + [blur^p]."""
    return (
        [diameter ** p]
        + [float(np.exp(e)) for e in np.arange(p * np.log(diameter), p * np.log(blur), p * np.log(scaling))]
        + [blur ** p]
    )


def _half_sqdist(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """C(x,y) = 0.5*|x-y|^2 expanded as 0.5*(|x|^2 - 2 x.y + |y|^2) (geomloss cost for p=2)."""
    d_xx = (x * x).sum(-1).unsqueeze(1)
    d_xy = x @ y.t()
    d_yy = (y * y).sum(-1).unsqueeze(0)
    return (d_xx - 2 * d_xy + d_yy) / 2


def _softmin(eps: float, C: torch.Tensor, h: torch.Tensor) -> torch.Tensor:
    """softmin_i = -eps * logsumexp_j(h_j - C_ij/eps)."""
    return -eps * (h.view(1, -1) - C / eps).logsumexp(dim=1)


def sinkhorn_divergence(x: torch.Tensor, y: torch.Tensor, blur: float = 0.05, scaling: float = 0.5,
                        p: int = 2) -> torch.Tensor:
    """Debiased Sinkhorn divergence S_eps(x, y) with uniform weights; differentiable w.r.t. x."""
    assert p == 2
    N, M = x.shape[0], y.shape[0]
    a_log = torch.full((N,), -float(np.log(N)), dtype=x.dtype, device=x.device)
    b_log = torch.full((M,), -float(np.log(M)), dtype=x.dtype, device=x.device)
    a = torch.full((N,), 1.0 / N, dtype=x.dtype, device=x.device)
    b = torch.full((M,), 1.0 / M, dtype=x.dtype, device=x.device)

    C_xx = _half_sqdist(x, x.detach())
    C_yy = _half_sqdist(y, y.detach())
    C_xy = _half_sqdist(x, y.detach())
    C_yx = _half_sqdist(y, x.detach())

    diameter = max_diameter(x.detach().reshape(-1, x.shape[-1]), y.detach().reshape(-1, y.shape[-1]))
    eps_list = epsilon_schedule(p, diameter, blur, scaling)

    with torch.no_grad():
        eps = eps_list[0]
        g_ab = _softmin(eps, C_yx, a_log)
        f_ba = _softmin(eps, C_xy, b_log)
        f_aa = _softmin(eps, C_xx, a_log)
        g_bb = _softmin(eps, C_yy, b_log)
        for eps in eps_list:
            ft_ba = _softmin(eps, C_xy, b_log + g_ab / eps)
            gt_ab = _softmin(eps, C_yx, a_log + f_ba / eps)
            ft_aa = _softmin(eps, C_xx, a_log + f_aa / eps)
            gt_bb = _softmin(eps, C_yy, b_log + g_bb / eps)
            f_ba, g_ab = 0.5 * (f_ba + ft_ba), 0.5 * (g_ab + gt_ab)
            f_aa, g_bb = 0.5 * (f_aa + ft_aa), 0.5 * (g_bb + gt_bb)

    # last extrapolation, differentiable through the cost matrices only
    f_ba, g_ab = (
        _softmin(eps, C_xy, (b_log + g_ab / eps).detach()),
        _softmin(eps, C_yx, (a_log + f_ba / eps).detach()),
    )
    f_aa = _softmin(eps, C_xx, (a_log + f_aa / eps).detach())
    g_bb = _softmin(eps, C_yy, (b_log + g_bb / eps).detach())
    return (a * (f_ba - f_aa)).sum() + (b * (g_ab - g_bb)).sum()


class SamplesLoss:
    """Call-compatible stand-in for `geomloss.SamplesLoss` restricted to what
    /root/reference/model/loss.py:202,221 uses."""

    def __init__(self, loss: str = "sinkhorn", p: int = 2, blur: float = 0.05, scaling: float = 0.5, **kw):
        if loss != "sinkhorn" or p != 2 or kw:
            raise NotImplementedError("oracle restates only SamplesLoss('sinkhorn', p=2, blur=...)")
        self.blur, self.scaling = blur, scaling

    def __call__(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        if x.dim() != 2 or y.dim() != 2 or x.shape[1] != y.shape[1]:
            raise ValueError("expected x [N,D], y [M,D]")
        return sinkhorn_divergence(x, y, blur=self.blur, scaling=self.scaling)
