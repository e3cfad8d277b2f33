"""TEST INFRASTRUCTURE ONLY — CPU restatement of the step epilogue and input-side helpers around the criterion
(/root/reference/tools/engine.py:16-18, 53-69; tools/train.py:264, 288-301).

The arithmetic lives in third-party packages: torch (present: `torch.optim.AdamW`, `torch.nn.utils.clip_grad_norm_` are
CALLED here, not restated) and timm 0.9.12 (absent from this image, pinned in /root/reference/requirements.txt:28):
`NativeScaler.__call__` (= torch GradScaler scale / unscale_ / step / update), `ModelEma.update`
(ema = decay * ema + (1 - decay) * model), `accuracy` (top-k), `mixup_target` / `Mixup._mix_batch` are restated from
their published forms.  The reference holds no test for any of them: parity of the timm pieces is unpinned (low risk).
"""
from __future__ import annotations

import torch


class ScalerState:
    """torch.amp.GradScaler's scalar state (defaults: init 65536, growth 2, backoff 0.5, interval 2000)."""

    def __init__(self, scale=65536.0, growth=2.0, backoff=0.5, interval=2000, dynamic=True):
        self.scale, self.growth, self.backoff, self.interval, self.dynamic = float(scale), growth, backoff, interval, dynamic
        self.tracker = 0


def epilogue_step(params, opt: torch.optim.AdamW, scaler: ScalerState, clip_grad, ema, ema_decay):
    """One `loss_scaler(loss, optimizer, clip_grad, parameters)` + `model_ema.update` on CPU tensors whose .grad hold the
    gradients of the SCALED loss.  Returns (skipped, grad_norm)."""
    inv = 1.0 / scaler.scale
    found_inf = any(not torch.isfinite(p.grad).all() for p in params)
    for p in params:
        p.grad.mul_(inv)                                            # GradScaler.unscale_
    norm = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(p.grad.double()) for p in params]))
    if clip_grad:
        torch.nn.utils.clip_grad_norm_(params, clip_grad)           # timm dispatch_clip_grad(mode='norm')
    if not found_inf:
        opt.step()                                                  # GradScaler.step skips on inf / nan
        if ema is not None:
            with torch.no_grad():
                for e, p in zip(ema, params):
                    e.mul_(ema_decay).add_(p.detach(), alpha=1.0 - ema_decay)   # timm ModelEma._update
    if scaler.dynamic:                                              # GradScaler.update
        if found_inf:
            scaler.scale *= scaler.backoff
            scaler.tracker = 0
        else:
            scaler.tracker += 1
            if scaler.tracker == scaler.interval:
                scaler.scale *= scaler.growth
                scaler.tracker = 0
    opt.zero_grad(set_to_none=False)
    return found_inf, float(norm)


def accuracy(output: torch.Tensor, target: torch.Tensor, topk=(1, 5)):
    """timm.utils.accuracy: percentage of rows whose target is among the k largest logits."""
    maxk = min(max(topk), output.size(1))
    _, pred = output.topk(maxk, 1, True, True)
    correct = pred.t().eq(target.reshape(1, -1).expand_as(pred.t()))
    return [correct[:min(k, maxk)].reshape(-1).float().sum(0) * 100.0 / target.size(0) for k in topk]


def mixup_target(target: torch.Tensor, num_classes: int, lam: float, smoothing: float) -> torch.Tensor:
    """timm.data.mixup.mixup_target: y1 * lam + y2 * (1 - lam), smoothed one-hots of target and target.flip(0)."""
    off = smoothing / num_classes
    on = 1.0 - smoothing + off
    def one_hot(t):
        return torch.full((t.shape[0], num_classes), off, dtype=torch.float64).scatter_(1, t.view(-1, 1), on)
    return one_hot(target) * lam + one_hot(target.flip(0)) * (1.0 - lam)


def mix_batch(x: torch.Tensor, lam: float, use_cutmix: bool, box) -> torch.Tensor:
    """timm Mixup._mix_batch on a copy of x."""
    x = x.clone()
    if lam == 1.0:
        return x
    if use_cutmix:
        yl, yh, xl, xh = box
        x[:, :, yl:yh, xl:xh] = x.flip(0)[:, :, yl:yh, xl:xh]
    else:
        x_flipped = x.flip(0).mul_(1.0 - lam)
        x.mul_(lam).add_(x_flipped)
    return x
