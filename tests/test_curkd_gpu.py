"""GPU: layer-wise hidden-state matching (CurKD early / mid) — tcgen05 path vs oracle and reference goldens."""
import pytest
import torch

from oracle import losses as O
from oracle.util import digest, rel_err
from tests.cases import build_case
from deltakd_b200 import heads as H

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-5
GRAD_RTOL = 1e-4


def _run(c):
    from deltakd_b200 import DistillationLoss, call_base_loss
    crit = DistillationLoss(call_base_loss(c.args), c.teacher, c.kind, c.alpha, c.tau)
    torch.manual_seed(4321)
    loss = crit(torch.zeros(c.B, 3, 2, 2, device="cuda"), c.outputs, c.student, c.s_feats, c.labels, c.args)
    loss.backward()
    return loss


@pytest.mark.parametrize("name", ["curkd_ep0", "curkd_ep120"])
def test_curkd_hidden_matches_reference(golden, name):
    c = build_case(name, device="cuda")
    loss = _run(c)
    heads = H.head_tensors(c.student)
    for tag in ("f32", "f64"):
        ref = float(golden[f"{name}/{tag}/loss"])
        assert abs(loss.item() - ref) <= LOSS_RTOL * abs(ref), (tag, loss.item(), ref)
        for i, f in enumerate(c.s_feats):
            key = f"{name}/{tag}/g_sfeat{i}"
            if key in golden.files:
                assert rel_err(digest(f.grad), golden[key]) < GRAD_RTOL, key
            else:
                assert f.grad is None
        for k, p in heads.items():
            key = f"{name}/{tag}/g_head/{k}"
            if key in golden.files:
                assert p.grad is not None, k
                assert rel_err(digest(p.grad), golden[key]) < GRAD_RTOL, key
    # full tensors vs the fp64 oracle
    o = build_case(name, dtype=torch.float64)
    oh = H.head_tensors(o.student)
    ol = O.distillation_loss(o.kind, o.outputs, o.labels, o.teacher_logits, o.s_feats, o.t_feats, oh, o.args,
                             o.alpha, o.tau)
    ol.backward()
    assert abs(loss.item() - ol.item()) <= LOSS_RTOL * abs(ol.item())
    for f, g in zip(c.s_feats, o.s_feats):
        if g.grad is not None:
            assert rel_err(f.grad, g.grad) < GRAD_RTOL
            assert float(f.grad[:, 0].abs().max()) == 0.0   # CLS row gets exactly zero gradient
    for k in oh:
        if oh[k].grad is not None:
            assert rel_err(heads[k].grad, oh[k].grad) < GRAD_RTOL, k


@pytest.mark.parametrize("B", [1, 3, 37])
@pytest.mark.parametrize("dtype,prec", [(torch.float32, "bf16x3"), (torch.float32, "bf16"), (torch.bfloat16, "bf16")])
def test_align_mse_shapes(B, dtype, prec):
    """ragged M (B*196 not a multiple of the 128-row tile), both dtypes / precisions; stated tolerances."""
    from deltakd_b200 import functional as Fn
    from deltakd_b200 import synth
    Fn.set_matmul_precision(prec)
    try:
        s_feats, t_feats = synth.make_features(B, 7, layers=[0, 1])
        lins = [torch.nn.Linear(192, 384) for _ in range(2)]
        ref_l = 0.0
        sd = [s.double().requires_grad_(True) for s in s_feats[:2]]
        if dtype == torch.bfloat16:
            sd = [s.bfloat16().double().requires_grad_(True) for s in s_feats[:2]]
        td = [(t.bfloat16() if dtype == torch.bfloat16 else t).double() for t in t_feats[:2]]
        for s, t, lin in zip(sd, td, lins):
            y = s[:, 1:] @ lin.weight.double().t() + lin.bias.double()
            ref_l = ref_l + ((y - t[:, 2:]) ** 2).sum() * 1e-4
        ref_l.backward()
        sc = [s.to(dtype).cuda().requires_grad_(True) for s in s_feats[:2]]
        tc = [t.to(dtype).cuda() for t in t_feats[:2]]
        lc = [torch.nn.Linear(192, 384).cuda() for _ in range(2)]
        for a, b in zip(lc, lins):
            a.load_state_dict(b.state_dict())
        loss = Fn.align_mse_layers_loss(sc, tc, lc, 1e-4)
        loss.backward()
        if prec == "bf16x3":
            lt, gt = 1e-5, 1e-4
        elif dtype == torch.float32:
            lt, gt = 2e-4, 6e-3    # single bf16 pass on fp32 data: operands rounded to 8 bits
        else:
            lt, gt = 2e-4, 8e-3    # bf16 storage: gradients also rounded to bf16 on store
        assert abs(loss.item() - ref_l.item()) <= lt * abs(ref_l.item()), (loss.item(), ref_l.item())
        for a, b in zip(sc, sd):
            assert rel_err(a.grad.float(), b.grad) < gt
        for a, b in zip(lc, lins):
            gw = torch.autograd.grad
        # head gradients against fp64 autograd
        ws = [lin.weight.double().detach().requires_grad_(True) for lin in lins]
        bs = [lin.bias.double().detach().requires_grad_(True) for lin in lins]
        l2 = 0.0
        for s, t, w, b in zip(sd, td, ws, bs):
            l2 = l2 + (((s.detach()[:, 1:] @ w.t() + b) - t[:, 2:]) ** 2).sum() * 1e-4
        l2.backward()
        for a, w, b in zip(lc, ws, bs):
            assert rel_err(a.weight.grad, w.grad) < gt
            assert rel_err(a.bias.grad, b.grad) < gt
    finally:
        Fn.set_matmul_precision(None)
