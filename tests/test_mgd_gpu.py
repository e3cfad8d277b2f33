"""GPU: masked generative distillation family (MGD, CurKD late phase) — tcgen05 implicit-GEMM convolutions
vs the oracle and the reference goldens.  The reference draws its mask noise with torch.rand on the CPU; the
tests inject that same noise tensor so that mask selection is bit-identical."""
from unittest import mock

import pytest
import torch

from oracle import losses as O
from oracle.util import digest, rel_err
from tests.cases import build_case
from deltakd_b200 import heads as H

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-5
GRAD_RTOL = 1e-4


def _run(c, probe=None):
    """One step through DistillationLoss; `probe` (dict) receives the kernel's mask and hidden activations (ReLU gate)."""
    from deltakd_b200 import DistillationLoss, call_base_loss
    from deltakd_b200 import functional as Fn
    crit = DistillationLoss(call_base_loss(c.args), c.teacher, c.kind, c.alpha, c.tau)
    noise = c.noise
    real = Fn.masked_generation_loss
    spy = (lambda *a, **k: real(*a, probe=probe, **k)) if probe is not None else real
    with mock.patch("torch.rand", side_effect=lambda *a, **k: noise.clone()), mock.patch.object(Fn, "masked_generation_loss", spy):
        loss = crit(torch.zeros(c.B, 3, 2, 2, device="cuda"), c.outputs, c.student, c.s_feats, c.labels, c.args)
    loss.backward()
    return loss


@pytest.mark.parametrize("name", ["mgd_r05", "mgd_r03", "curkd_ep200", "vitkd"])   # vitkd = 2-layer mimic + generation (loss.py:251-311)
def test_masked_generation_matches_reference(golden, name):
    c = build_case(name, device="cuda")
    probe = {}
    loss = _run(c, probe)
    heads = H.head_tensors(c.student)
    for tag in ("f32", "f64"):
        ref = float(golden[f"{name}/{tag}/loss"])
        assert abs(loss.item() - ref) <= LOSS_RTOL * abs(ref), (tag, loss.item(), ref)
    # gradients: against the fp64 oracle evaluated with the gate the KERNEL used (read back, not searched); every
    # difference from the oracle's own gate must sit at an ambiguous pre-activation (tests/gate_search.py)
    from tests.gate_search import gate_from_hidden, kernel_gate_oracle
    ours = {k: p.grad.detach().double().cpu() for k, p in heads.items() if p.grad is not None}
    for i, f in enumerate(c.s_feats):
        if f is not None and f.grad is not None:
            ours[f"s{i}"] = f.grad.detach().double().cpu()

    def eval_fn(pr):
        o = build_case(name, dtype=torch.float64)
        oh = H.head_tensors(o.student)
        l = O.distillation_loss(o.kind, o.outputs, o.labels, o.teacher_logits, o.s_feats, o.t_feats, oh, o.args,
                                o.alpha, o.tau, noise=o.noise, probe=pr)
        l.backward()
        g = {k: v.grad for k, v in oh.items() if v.grad is not None}
        for i, f in enumerate(o.s_feats):
            if f.grad is not None:
                g[f"s{i}"] = f.grad
        return l, g

    ol, grads, n_amb, n_flip = kernel_gate_oracle(eval_fn, gate_from_hidden(probe["hidden"]).cpu())
    print(f"[{name}] ReLU gates: {n_amb} ambiguous, {n_flip} differ from the fp64 oracle's")
    assert abs(loss.item() - ol.item()) <= LOSS_RTOL * abs(ol.item())
    checked = 0
    for k, g in grads.items():
        if float(g.abs().sum()) == 0:
            continue
        assert k in ours, k
        assert rel_err(ours[k], g) < GRAD_RTOL, (k, n_amb, n_flip)
        checked += 1
    assert checked >= 8  # g_s, align w/b, mask_token, 2 x (conv w, b)
    assert float(c.s_feats[11].grad[:, 0].abs().max()) == 0.0
    if name == "vitkd":
        assert c.s_feats[0].grad is not None and c.s_feats[1].grad is not None and c.s_feats[5].grad is None
    if n_flip == 0:  # the reference's recorded gradients (its own gate)
        for tag in ("f32", "f64"):
            for k, p in heads.items():
                key = f"{name}/{tag}/g_head/{k}"
                if key in golden.files and p.grad is not None:
                    assert rel_err(digest(p.grad), golden[key]) < GRAD_RTOL, key


@pytest.mark.parametrize("B,ratio", [(1, 0.5), (4, 0.25), (7, 0.9)])
@pytest.mark.parametrize("mode", ["bf16x3", "bf16"])
def test_generation_shapes_and_modes(B, ratio, mode):
    """tile tails (B*14 image rows not a multiple of 9 / 8), mask ratios, both precisions (stated tolerances).
    Gradients are compared given the same ReLU gate (tests/gate_search.py)."""
    from deltakd_b200 import functional as Fn
    from deltakd_b200 import synth
    from types import SimpleNamespace
    from tests.gate_search import best_gate_oracle
    Fn.set_matmul_precision(mode)
    try:
        args = SimpleNamespace(distillation_type="mgd")
        teacher, student = synth.FeatureReplayModel(384), synth.FeatureReplayModel(192)
        torch.manual_seed(1)
        H.attach_distillation_heads(student, teacher, args)
        with torch.no_grad():
            student.mask_token.normal_(0, 0.1)
        s_feats, t_feats = synth.make_features(B, 3, layers=[11])
        noise = synth.make_noise(B, seed=B)
        host_heads = {k: v.detach().double() for k, v in H.head_tensors(student).items()}
        student = student.cuda()
        sc = s_feats[11].cuda().requires_grad_(True)
        loss = Fn.masked_generation_loss(sc, t_feats[11].cuda(), student.align, student.mask_token, student.generation,
                                         mask_ratio=ratio, noise=noise.cuda(), scale=7e-5 / (B * 196 * 384))
        loss.backward()
        ours = {k: p.grad.detach().cpu().double() for k, p in H.head_tensors(student).items()}
        ours["s"] = sc.grad.detach().cpu().double()

        def eval_fn(probe):
            heads64 = {k: v.clone().requires_grad_(True) for k, v in host_heads.items()}
            s64 = s_feats[11].double().requires_grad_(True)
            l = O.mgd([s64], [t_feats[11].double()], heads64, 7e-5, ratio, noise, probe=probe)
            l.backward()
            g = {k: v.grad for k, v in heads64.items()}
            g["s"] = s64.grad
            return l, g

        tau = 3e-5 if mode == "bf16x3" else 3e-2
        ref, grads, n_amb, n_flip = best_gate_oracle(eval_fn, ours, tau=tau, max_flips=160 if mode == "bf16x3" else 0)
        lt, gt = (1e-5, 1e-4) if mode == "bf16x3" else (5e-3, 6e-2)   # one bf16 pass through two K=3456 convolutions
        assert abs(loss.item() - ref.item()) <= lt * abs(ref.item()), (loss.item(), ref.item())
        for k in ours:
            assert rel_err(ours[k], grads[k]) < gt, (k, n_amb, n_flip)
    finally:
        Fn.set_matmul_precision(None)


def test_persistent_multi_tile_and_linearity():
    """B=160: 249 conv tiles > 148 CTAs, so every persistent loop (ring phases, TMEM hand-over) wraps.
    Size-independent checks: the loss is the sum of the per-sample losses, and scaling `scale` scales it."""
    from deltakd_b200 import functional as Fn
    from deltakd_b200 import synth
    from types import SimpleNamespace
    B = 160
    args = SimpleNamespace(distillation_type="mgd")
    teacher, student = synth.FeatureReplayModel(384), synth.FeatureReplayModel(192)
    torch.manual_seed(2)
    H.attach_distillation_heads(student, teacher, args)
    student = student.cuda()
    s_feats, t_feats = synth.make_features(B, 5, layers=[11])
    s, t = s_feats[11].cuda(), t_feats[11].cuda()
    noise = synth.make_noise(B, seed=8).cuda()

    def run(sl, scale):
        x = s[sl].clone().requires_grad_(True)
        for p in student.parameters():
            p.grad = None
        l = Fn.masked_generation_loss(x, t[sl], student.align, student.mask_token, student.generation,
                                      mask_ratio=0.5, noise=noise[sl], scale=scale)
        l.backward()
        return l.item(), x.grad, student.generation[0].weight.grad.clone()

    full, g_full, gw_full = run(slice(0, B), 1e-6)
    parts = [run(slice(i, i + 40), 1e-6) for i in range(0, B, 40)]
    assert abs(sum(p[0] for p in parts) - full) <= 2e-5 * abs(full)
    assert rel_err(torch.cat([p[1] for p in parts]), g_full) < 1e-4     # per-sample gradients are independent
    assert rel_err(sum(p[2] for p in parts), gw_full) < 1e-4           # weight gradients add up
    twice, g2, _ = run(slice(0, B), 2e-6)
    assert abs(twice - 2 * full) <= 1e-5 * abs(twice)
