"""GPU: WassKD 'l1' (sorted-L1 over the token axis) — on-chip bitonic sort + tcgen05 alignment GEMMs vs the
oracle and the reference goldens.

The gradient of |sort(a) - sort(t)| is a sign: it is discontinuous where a sorted student value meets its
teacher partner.  The kernel's aligned activations differ from fp64 by ~2^-16 relative (bf16x3), so an
element whose |diff| is below that level may legitimately take the other sign.  The tests therefore allow,
on top of the 1e-4 relative gate, the exact amount such "ambiguous" elements can move a gradient
(`flip budget`; zero or one element at these sizes) and report how many there were."""
import pytest
import torch

from oracle import losses as O
from oracle.util import rel_err
from tests.cases import build_case
from deltakd_b200 import heads as H

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-5
GRAD_RTOL = 1e-4
AMBIG_TAU = 2e-5   # |sorted diff| below this (relative to 1 + |a|) counts as ambiguous


def _flip_budget(o, heads64, weight=5.0):
    """Per-tensor bound (fp64 norms) on how far the gradients can move if every ambiguous sign flipped."""
    n = 3
    budget = {}
    for i in range(3):
        W = heads64[f"align_wasskd.{i}.weight"].detach()
        s = o.s_feats[i].detach()[:, 1:]
        a = s @ W.t() + heads64[f"align_wasskd.{i}.bias"].detach()
        sa, pi = torch.sort(a, dim=1)
        st, _ = torch.sort(o.t_feats[i][:, 2:], dim=1)
        d = sa - st
        amb = d.abs() < AMBIG_TAU * (1 + sa.abs())
        c = weight / (n * a.numel())
        nb = int(amb.sum())
        # a flip changes g_a by 2c at one (b, token, channel): bound each gradient's change
        gs = gw = gb = 0.0
        if nb:
            bb, rr, dd = amb.nonzero(as_tuple=True)
            tok = pi[bb, rr, dd]
            gs = float((2 * c * W[dd].norm(dim=1)).pow(2).sum().sqrt())
            gw = float((2 * c * s[bb, tok].norm(dim=1)).pow(2).sum().sqrt())
            gb = float(2 * c * nb ** 0.5)
        budget[i] = (nb, gs, gw, gb)
    return budget


def _close(ours, ref, extra, what):
    ours, ref = ours.detach().double().cpu(), ref.detach().double().cpu()
    err = float((ours - ref).norm())
    assert err <= GRAD_RTOL * float(ref.norm()) + 1.0001 * extra, (what, err / float(ref.norm()), extra)


def test_wass_l1_matches_reference(golden):
    name = "wass_l1"
    from deltakd_b200 import DistillationLoss, call_base_loss
    c = build_case(name, device="cuda")
    crit = DistillationLoss(call_base_loss(c.args), c.teacher, c.kind, c.alpha, c.tau)
    loss = crit(torch.zeros(c.B, 3, 2, 2, device="cuda"), c.outputs, c.student, c.s_feats, c.labels, c.args)
    loss.backward()
    for tag in ("f32", "f64"):
        ref = float(golden[f"{name}/{tag}/loss"])
        assert abs(loss.item() - ref) <= LOSS_RTOL * abs(ref), (tag, loss.item(), ref)
    o = build_case(name, dtype=torch.float64)
    oh = H.head_tensors(o.student)
    ol = O.distillation_loss(o.kind, o.outputs, o.labels, o.teacher_logits, o.s_feats, o.t_feats, oh, o.args, o.alpha, o.tau)
    ol.backward()
    assert abs(loss.item() - ol.item()) <= LOSS_RTOL * abs(ol.item())
    budget = _flip_budget(o, oh)
    heads = H.head_tensors(c.student)
    for i in range(3):
        nb, gs, gw, gb = budget[i]
        assert nb <= 64, f"layer {i}: {nb} ambiguous elements"
        _close(c.s_feats[i].grad, o.s_feats[i].grad, gs, f"g_sfeat{i} ({nb} ambiguous)")
        assert float(c.s_feats[i].grad[:, 0].abs().max()) == 0.0
        _close(heads[f"align_wasskd.{i}.weight"].grad, oh[f"align_wasskd.{i}.weight"].grad, gw, f"g_W{i}")
        _close(heads[f"align_wasskd.{i}.bias"].grad, oh[f"align_wasskd.{i}.bias"].grad, gb, f"g_b{i}")
    for i in range(3, 12):
        assert c.s_feats[i].grad is None


@pytest.mark.parametrize("B", [1, 5])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_wass_l1_sorted_properties(B, dtype):
    """Size-independent properties: (1) the loss is invariant under a permutation of each sample's tokens
    (of student and teacher independently); (2) if the student equals the teacher up to a token permutation
    the loss is ~0; (3) loss of the batch = mean of the per-sample losses."""
    from deltakd_b200 import functional as Fn
    from deltakd_b200 import synth
    s_feats, t_feats = synth.make_features(B, 11, layers=[0])
    lin = torch.nn.Linear(192, 384).cuda()
    s = s_feats[0].to(dtype).cuda()
    t = t_feats[0].to(dtype).cuda()
    base = Fn.wass_l1_loss([s], [t], [lin], weight=1.0).item()
    g = torch.Generator().manual_seed(B)
    ps = torch.cat([torch.zeros(1, dtype=torch.long), 1 + torch.randperm(196, generator=g)]).cuda()
    pt = torch.cat([torch.arange(2), 2 + torch.randperm(196, generator=g)]).cuda()
    perm = Fn.wass_l1_loss([s[:, ps].contiguous()], [t[:, pt].contiguous()], [lin], weight=1.0).item()
    assert abs(perm - base) <= 2e-6 * abs(base)
    if B > 1:
        parts = [Fn.wass_l1_loss([s[b:b + 1]], [t[b:b + 1]], [lin], weight=1.0).item() for b in range(B)]
        assert abs(sum(parts) / B - base) <= 1e-5 * abs(base)
    # student == teacher (up to permutation) through an identity-like head: W = [I; I], b = 0 on a doubled feature
    with torch.no_grad():
        lin.weight.zero_(); lin.bias.zero_()
        lin.weight[:192, :] = torch.eye(192); lin.weight[192:, :] = torch.eye(192)
    t2 = torch.zeros_like(t)
    t2[:, 2:] = torch.cat([s[:, 1:], s[:, 1:]], dim=-1)[:, torch.randperm(196, generator=g).cuda()]
    zero = Fn.wass_l1_loss([s], [t2], [lin], weight=1.0).item()
    assert abs(zero) <= (2e-5 if dtype == torch.float32 else 1e-2) * max(1.0, abs(base))  # bf16x3 identity: 2^-16
