"""Tiny feature-loss invocations for `__graft_entry__.smoke()`: one tcgen05 path (CurKD hidden-state matching) and one
sort path (WassKD l1) at B = 2, checked against the CPU oracle (test infrastructure; smoke() is allowed to use it)."""
from __future__ import annotations

import torch


def run(dev) -> None:
    from deltakd_b200 import functional as Fn
    from deltakd_b200 import synth
    from oracle import losses as O
    s_feats, t_feats = synth.make_features(2, 11, layers=[0])
    torch.manual_seed(0)
    lin = torch.nn.Linear(192, 384)
    s = s_feats[0].to(dev).requires_grad_(True)
    t = t_feats[0].to(dev)
    lin_d = torch.nn.Linear(192, 384).to(dev)
    lin_d.load_state_dict(lin.state_dict())
    loss = Fn.align_mse_layers_loss([s], [t], [lin_d], scale=4e-5 / 2)
    loss.backward()
    s64 = s_feats[0].double().requires_grad_(True)
    y = O._align(s64, lin.weight.double(), lin.bias.double())
    ref = ((y - t_feats[0].double()[:, 2:]) ** 2).sum() * (4e-5 / 2)
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item()), (loss.item(), ref.item())
    err = (s.grad.double().cpu() - s64.grad).norm() / s64.grad.norm()
    assert float(err) < 1e-4, float(err)
    w = Fn.wass_l1_loss([s.detach()], [t], [lin_d], weight=1.0)
    a = O._align(s_feats[0].double(), lin.weight.double(), lin.bias.double())
    wref = (torch.sort(a, dim=1)[0] - torch.sort(t_feats[0].double()[:, 2:], dim=1)[0]).abs().mean()
    assert abs(w.item() - wref.item()) <= 1e-5 * abs(wref.item()), (w.item(), wref.item())
