"""GPU: host-side contract of the autograd wrappers: float16 inputs under the reference trainer's
autocast call, per-stream scratch, and a loud error on a second backward."""
import pytest
import torch

from oracle import losses as O
from oracle.util import rel_err
from deltakd_b200 import synth

pytestmark = pytest.mark.gpu


def test_float16_logits_under_reference_autocast():
    """tools/engine.py:23-29 runs the student under torch.cuda.amp.autocast(enabled=True) (float16) and the criterion
    outside it: float16 logits are upcast exactly at the boundary and the gradients come back as float16."""
    from deltakd_b200 import DistillationLoss, call_base_loss
    B, C = 16, 1000
    z, zk, zt, y = synth.make_logits(B, C, 5)
    lin = torch.nn.Linear(C, C).cuda()
    x = z.cuda()
    with torch.autocast("cuda", enabled=True):   # float16, as in the reference engine
        out = lin(x)
        out_kd = lin(zk.cuda())
    assert out.dtype == torch.float16
    args = synth.default_args()
    teacher = synth.FeatureReplayModel(384)
    teacher.set_outputs(zt.cuda(), None)
    crit = DistillationLoss(call_base_loss(args), teacher, "soft", 0.1, 3.0)
    loss = crit(torch.zeros(B, 3, 2, 2, device="cuda"), (out, out_kd), None, None, y.cuda(), args)
    loss.backward()
    assert lin.weight.grad is not None and torch.isfinite(lin.weight.grad).all()
    o1 = out.detach().double().cpu().requires_grad_(True)
    o2 = out_kd.detach().double().cpu().requires_grad_(True)
    ref = O.distillation_loss("soft", (o1, o2), y.double(), zt.double(), None, None, {}, args, 0.1, 3.0)
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())


def test_float16_features_are_accepted():
    from deltakd_b200 import functional as Fn
    s_feats, t_feats = synth.make_features(3, 9, layers=[0])
    lin = torch.nn.Linear(192, 384).cuda()
    s16 = s_feats[0].cuda().half().requires_grad_(True)
    t16 = t_feats[0].cuda().half()
    loss = Fn.align_mse_layers_loss([s16], [t16], [lin], 1e-4)
    loss.backward()
    assert s16.grad is not None and s16.grad.dtype == torch.float16
    s64 = s16.detach().double().cpu().requires_grad_(True)
    ref = (((s64[:, 1:] @ lin.weight.double().cpu().t() + lin.bias.double().cpu()) - t16.double().cpu()[:, 2:]) ** 2).sum() * 1e-4
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert rel_err(s16.grad.float(), s64.grad) < 2e-3   # the gradient is rounded to float16 on the way back


def test_unsupported_widths_raise_clearly():
    from deltakd_b200 import functional as Fn
    s = torch.randn(2, 197, 384, device="cuda", requires_grad=True)
    t = torch.randn(2, 198, 768, device="cuda")
    with pytest.raises(ValueError, match="192 -> 384"):
        Fn.align_mse_layers_loss([s], [t], [torch.nn.Linear(384, 768).cuda()], 1e-4)
    s2 = torch.randn(2, 577, 192, device="cuda", requires_grad=True)   # 384-px inputs: 24 x 24 patches
    t2 = torch.randn(2, 578, 384, device="cuda")
    with pytest.raises(ValueError, match="196 patch tokens"):
        Fn.wass_sinkhorn_loss([s2], [t2], [torch.nn.Linear(192, 384).cuda()])


def test_second_backward_is_refused():
    from deltakd_b200 import functional as Fn
    z, zk, zt, y = (t.cuda() for t in synth.make_logits(8, 100, 3))
    z.requires_grad_(True); zk.requires_grad_(True)
    loss = Fn.logit_kd_loss(z, zk, zt, y, kd_kind="soft", alpha=0.1, tau=3.0)
    loss.backward(retain_graph=True)
    with pytest.raises(RuntimeError, match="already backpropagated"):
        loss.backward()


def test_two_streams_do_not_share_scratch():
    """The same op on two streams at once: each stream has its own ticket counter / arenas (functional._WS is keyed by
    stream), so both results equal the single-stream result."""
    from deltakd_b200 import functional as Fn
    z, zk, zt, y = (t.cuda() for t in synth.make_logits(256, 1000, 11))
    want = Fn.logit_kd_loss(z, zk, zt, y, kd_kind="soft", alpha=0.1, tau=3.0).item()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    outs = []
    for _ in range(20):
        for st in (s1, s2):
            with torch.cuda.stream(st):
                outs.append(Fn.logit_kd_loss(z, zk, zt, y, kd_kind="soft", alpha=0.1, tau=3.0))
    torch.cuda.synchronize()
    assert all(o.item() == want for o in outs)
    keys = [k for k in Fn._WS if k[2] == "logit_kd"]
    assert len({k[1] for k in keys}) >= 3   # default stream + the two side streams


@pytest.mark.parametrize("entry", ["dkd_align_mse", "dkd_wass_l1"])
def test_layers_on_their_own_streams_give_the_same_result(entry, monkeypatch):
    """The multi-layer feature losses fork every layer onto its own stream (functional._LAYER_STREAMS): same kernels, same
    per-layer scratch sizes — the feature gradients must be bit-identical to the single-stream order, the loss and the
    (atomically reduced) weight gradients equal up to the order of the adds; the fork / join must also survive a CUDA-graph capture and two replays."""
    from deltakd_b200 import functional as Fn
    B = 5
    s_feats, t_feats = synth.make_features(B, 21, layers=[0, 1, 2])
    lins = [torch.nn.Linear(192, 384).cuda() for _ in range(3)]

    def run(multi):
        monkeypatch.setattr(Fn, "_LAYER_STREAMS", multi)
        s = [s_feats[i].cuda().requires_grad_(True) for i in range(3)]
        t = [t_feats[i].cuda() for i in range(3)]
        for lin in lins:
            lin.zero_grad(set_to_none=True)
        loss = Fn.align_mse_layers_loss(s, t, lins, scale=1e-3, _entry=entry)
        loss.backward()
        torch.cuda.synchronize()
        return loss.item(), [x.grad.clone() for x in s], [lin.weight.grad.clone() for lin in lins]

    l1, gs1, gw1 = run(True)
    l0, gs0, gw0 = run(False)
    assert abs(l1 - l0) <= 1e-6 * abs(l0)
    for a, b in zip(gs1, gs0):
        assert torch.equal(a, b)
    for a, b in zip(gw1, gw0):   # split-K weight gradients are reduced with fp32 atomics: equal up to the order of the adds
        assert rel_err(a, b) < 1e-5

    # under capture: fork / join become graph dependencies
    monkeypatch.setattr(Fn, "_LAYER_STREAMS", True)
    s = [s_feats[i].cuda().requires_grad_(True) for i in range(3)]
    t = [t_feats[i].cuda() for i in range(3)]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        Fn.align_mse_layers_loss(s, t, lins, scale=1e-3, _entry=entry).backward()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    for x in s:
        x.grad = None
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        loss = Fn.align_mse_layers_loss(s, t, lins, scale=1e-3, _entry=entry)
        loss.backward()
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    assert abs(loss.item() - l0) <= 1e-6 * abs(l0)
    for a, b in zip([x.grad for x in s], gs0):
        assert torch.equal(a, b)
