"""GPU: step epilogue (scaler + clip + AdamW + EMA), top-k accuracy, Mixup / CutMix with in-kernel soft labels — §8f
ranks 3 and 4 — against torch's own CPU optimizer / clipping and the restated timm pieces (oracle/step.py)."""
import copy

import numpy as np
import pytest
import torch

from oracle import losses as O
from oracle import step as S
from oracle.util import rel_err
from deltakd_b200 import synth

pytestmark = pytest.mark.gpu


def _model(seed=0):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Linear(37, 64), torch.nn.LayerNorm(64), torch.nn.GELU(), torch.nn.Linear(64, 10, bias=True),
                               torch.nn.Linear(10, 3, bias=False))


@pytest.mark.parametrize("clip,ema_decay,loss_scale", [(None, None, None), (0.5, 0.99, None), (1.0, 0.999, 1024.0)])
def test_step_epilogue_matches_torch(clip, ema_decay, loss_scale):
    from deltakd_b200 import FusedStepEpilogue
    ref = _model().double()
    ours = copy.deepcopy(ref).float().cuda()
    rp = list(ref.parameters())
    decay = [p for p in rp if p.ndim > 1]
    no_decay = [p for p in rp if p.ndim <= 1]
    opt = torch.optim.AdamW([{"params": decay, "weight_decay": 0.05}, {"params": no_decay, "weight_decay": 0.0}], lr=3e-3,
                            betas=(0.9, 0.999), eps=1e-8)
    ema = [p.detach().clone() for p in rp] if ema_decay else None
    sc = S.ScalerState(scale=loss_scale or 1.0, interval=3, dynamic=loss_scale is not None)
    epi = FusedStepEpilogue(ours.parameters(), lr=3e-3, weight_decay=0.05, clip_grad=clip, ema_decay=ema_decay,
                            loss_scale=loss_scale, growth_interval=3)
    g = torch.Generator().manual_seed(1)
    for it in range(7):
        x = torch.randn(16, 37, generator=g)
        inject_inf = loss_scale is not None and it == 4
        # reference
        loss_r = (ref(x.double()) ** 2).mean() * sc.scale
        loss_r.backward()
        if inject_inf:
            rp[0].grad[0, 0] = float("inf")
        skipped, norm = S.epilogue_step(rp, opt, sc, clip, ema, ema_decay)
        # ours
        loss_o = (ours(x.cuda()) ** 2).mean()
        epi.scale(loss_o).backward()
        if inject_inf:
            next(iter(ours.parameters())).grad[0, 0] = float("inf")
        epi.step()
        assert bool(epi.skipped.item()) == skipped, it
        if not skipped:
            assert abs(epi.grad_norm.item() - norm) <= 1e-5 * norm
        assert abs(epi.loss_scale.item() - sc.scale) <= 1e-6 * sc.scale, (it, epi.loss_scale.item(), sc.scale)
        for po, pr in zip(ours.parameters(), ref.parameters()):
            assert rel_err(po, pr) < 2e-6, it
            assert float(po.grad.abs().max()) == 0.0          # zero_grad fused into the update
    if ema_decay:
        by_id = {id(p): e for p, e in zip(epi.params, epi.ema_tensors())}
        for po, e_ref in zip(ours.parameters(), ema):
            assert rel_err(by_id[id(po)], e_ref) < 2e-6


def test_step_epilogue_large_flat_buffer():
    """8.4 M parameters (DeiT-Tiny + MGD heads scale): one step against torch.optim.AdamW run on the GPU in fp32."""
    from deltakd_b200 import FusedStepEpilogue
    torch.manual_seed(3)
    ws = [torch.nn.Parameter(torch.randn(1024, 2048, device="cuda") * 0.02) for _ in range(4)] + \
         [torch.nn.Parameter(torch.randn(2048, device="cuda") * 0.02) for _ in range(5)]
    ref = [torch.nn.Parameter(w.detach().clone()) for w in ws]
    opt = torch.optim.AdamW([{"params": ref[:4], "weight_decay": 0.05}, {"params": ref[4:], "weight_decay": 0.0}], lr=5e-4)
    epi = FusedStepEpilogue(ws, lr=5e-4, weight_decay=0.05, clip_grad=1.0)
    for _ in range(2):
        gs = [torch.randn_like(w) * 0.01 for w in ws]
        for w, r, g in zip(ws, ref, gs):
            w.grad.copy_(g)
            r.grad = g.clone()
        torch.nn.utils.clip_grad_norm_(ref, 1.0)
        opt.step()
        epi.step()
    for w, r in zip(ws, ref):
        assert rel_err(w, r) < 1e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_topk_accuracy(dtype):
    from deltakd_b200 import accuracy
    g = torch.Generator().manual_seed(5)
    z = torch.randn(257, 1000, generator=g).to(dtype)
    t = torch.randint(0, 1000, (257,), generator=g)
    z[torch.arange(0, 257, 3), t[::3]] += 4.0        # a good share of hits
    a1, a5 = accuracy(z.cuda(), t.cuda(), topk=(1, 5))
    r1, r5 = S.accuracy(z.float(), t, topk=(1, 5))
    # bf16 logits tie often and torch.topk's order among equal values is unspecified: the kernel ranks ties lower index
    # first; it must sit between the strict (ties lose) and the lenient (ties win) counts, and equal torch when no target ties
    zf = z.float()
    zt = zf.gather(1, t.view(-1, 1))
    hi_rank = (zf >= zt).sum(1) - 1          # ties all ahead of the target
    lo_rank = (zf > zt).sum(1)               # ties all behind
    for a, r, k in ((a1, r1, 1), (a5, r5, 5)):
        lo = (hi_rank < k).float().mean().item() * 100
        hi = (lo_rank < k).float().mean().item() * 100
        assert lo - 1e-4 <= a.item() <= hi + 1e-4, (k, lo, a.item(), hi)
        if lo == hi:
            assert abs(a.item() - r.item()) < 1e-4
    exact = ((zf > zt).sum(1) + ((zf == zt) & (torch.arange(1000).view(1, -1) < t.view(-1, 1))).sum(1))
    assert abs(a5.item() - (exact < 5).float().mean().item() * 100) < 1e-4
    # tuple outputs of the distilled student are reduced by the caller (engine.py:50-51); small C clamps k
    a1, a5 = accuracy(z[:, :3].contiguous().cuda(), (t % 3).cuda(), topk=(1, 5))
    assert a5.item() == 100.0


@pytest.mark.parametrize("use_cutmix", [False, True])
def test_mixup_batch_and_fused_labels(use_cutmix):
    """Images: one in-place kernel == timm's _mix_batch.  Labels: the logit kernel fed MixedLabels == the fp64 oracle fed
    timm's dense mixup_target, for the base CE alone and for soft-KD through DistillationLoss."""
    from deltakd_b200 import DistillationLoss, Mixup, call_base_loss
    B, C = 64, 1000
    g = torch.Generator().manual_seed(9)
    x = torch.randn(B, 3, 32, 32, generator=g)
    y = torch.randint(0, C, (B,), generator=g)
    mix = Mixup(mixup_alpha=0.0 if use_cutmix else 0.8, cutmix_alpha=1.0 if use_cutmix else 0.0, label_smoothing=0.1, num_classes=C)
    np.random.seed(4)
    xg = x.clone().cuda()
    xm, labels = mix(xg, y.cuda())
    lam = float(labels.lam.item())
    assert 0.0 < lam < 1.0
    # replay the host draws to get the box
    np.random.seed(4)
    lam0, cm = mix._params_per_batch()
    box = (0, 0, 0, 0)
    if cm:
        from deltakd_b200.mixup import rand_bbox
        box = rand_bbox(x.shape, lam0)
    assert cm == use_cutmix
    ref_x = S.mix_batch(x, lam if not cm else lam0, cm, box)
    assert torch.allclose(xm.cpu(), ref_x, rtol=1e-6, atol=1e-6)
    assert rel_err(labels.dense(), S.mixup_target(y, C, lam, 0.1)) < 1e-6
    # losses
    z, zk, zt, _ = synth.make_logits(B, C, 21)
    args = synth.default_args()
    dense = S.mixup_target(y, C, lam, 0.1)
    for kind in ("none", "soft"):
        teacher = synth.FeatureReplayModel(384)
        teacher.set_outputs(zt.cuda(), None)
        crit = DistillationLoss(call_base_loss(args), teacher, kind, 0.1, 3.0)
        zc, zkc = z.cuda().requires_grad_(True), zk.cuda().requires_grad_(True)
        out = zc if kind == "none" else (zc, zkc)
        loss = crit(torch.zeros(B, 3, 2, 2, device="cuda"), out, None, None, labels, args)
        loss.backward()
        zo, zko = z.double().requires_grad_(True), zk.double().requires_grad_(True)
        ref = O.distillation_loss(kind, zo if kind == "none" else (zo, zko), dense, zt.double(), None, None, {}, args, 0.1, 3.0)
        ref.backward()
        assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item()), kind
        assert rel_err(zc.grad, zo.grad) < 1e-4
        if kind == "soft":
            assert rel_err(zkc.grad, zko.grad) < 1e-4


def test_mixed_labels_large_batch_ring_kernel():
    """B = 2048 (the bulk-copy ring kernel path) with in-kernel mixed labels, bf16 logits."""
    from deltakd_b200 import functional as Fn
    B, C = 2048, 1000
    z, zk, zt, _ = synth.make_logits(B, C, 33)
    y = torch.randint(0, C, (B,), generator=torch.Generator().manual_seed(2))
    lam = torch.tensor([0.37], device="cuda")
    zc = z.bfloat16().cuda().requires_grad_(True)
    zkc = zk.bfloat16().cuda().requires_grad_(True)
    loss = Fn.logit_kd_loss(zc, zkc, zt.bfloat16().cuda(), y.cuda(), kd_kind="soft", smoothing=0.1, alpha=0.1, tau=3.0, mix_lam=lam)
    loss.backward()
    zo = zc.detach().double().cpu().requires_grad_(True)
    zko = zkc.detach().double().cpu().requires_grad_(True)
    ref = O.distillation_loss("soft", (zo, zko), S.mixup_target(y, C, 0.37, 0.1), zt.bfloat16().double(), None, None, {},
                              synth.default_args(), 0.1, 3.0)
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert rel_err(zc.grad.float(), zo.grad) < 6e-3 and rel_err(zkc.grad.float(), zko.grad) < 6e-3
