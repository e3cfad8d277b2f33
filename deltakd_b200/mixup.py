"""Mixup / CutMix for the distillation step — the input side of the base criterion.

The reference builds `timm.data.Mixup(mixup_alpha, cutmix_alpha, cutmix_minmax, prob, switch_prob, mode, label_smoothing,
num_classes)` (/root/reference/tools/train.py:288-295) and calls `samples, targets = mixup_fn(samples, targets)` every
step (tools/engine.py:16-18); timm then materialises a [B, num_classes] soft-label tensor that the base criterion
(SoftTargetCrossEntropy) reads back.  Here:
  * the random draws (lam, CutMix box) stay host-side numpy calls in timm's order (mode 'batch');
  * the images are mixed in place by one kernel (`dkd_mix_batch`: each pair (b, B-1-b) read once, written once);
  * the targets come back as `MixedLabels` (int64 ids + lam): `DistillationLoss` / `SoftTargetCrossEntropy` pass them to
    the fused logit kernel, which generates lam*smooth_onehot(y[r]) + (1-lam)*smooth_onehot(y[B-1-r]) on the fly — the
    soft-label tensor never exists (`.dense()` builds it for callers that want it).
timm is absent from this image: the class restates timm 0.9.12's `Mixup` (mode 'batch', elementwise/pair modes are not
used by the reference) — parity unpinned on the RNG stream, pinned on the arithmetic by tests against this restatement.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .functional import _require_cuda, _stream


class MixedLabels:
    """Targets of a mixed batch: class ids [B] (int64), lam (fp32 device scalar), label smoothing, class count."""

    def __init__(self, target: torch.Tensor, lam: torch.Tensor, smoothing: float, num_classes: int):
        self.target, self.lam, self.smoothing, self.num_classes = target, lam, float(smoothing), int(num_classes)

    @property
    def device(self):
        return self.target.device

    @property
    def shape(self):
        return (self.target.shape[0], self.num_classes)

    def to(self, *a, **k):
        return MixedLabels(self.target.to(*a, **k), self.lam.to(*a, **k), self.smoothing, self.num_classes)

    def dense(self) -> torch.Tensor:
        """timm mixup_target: y1 * lam + y2 * (1 - lam) with smoothed one-hots of target and target.flip(0)."""
        off = self.smoothing / self.num_classes
        on = 1.0 - self.smoothing + off
        def one_hot(t):
            return torch.full((t.shape[0], self.num_classes), off, device=t.device).scatter_(1, t.view(-1, 1), on)
        lam = self.lam.reshape(())
        return one_hot(self.target) * lam + one_hot(self.target.flip(0)) * (1.0 - lam)


def rand_bbox(img_shape, lam, margin=0.0):
    """timm.data.mixup.rand_bbox: box of area ratio (1 - lam), centre uniform, clipped."""
    ratio = np.sqrt(1 - lam)
    img_h, img_w = img_shape[-2:]
    cut_h, cut_w = int(img_h * ratio), int(img_w * ratio)
    margin_y, margin_x = int(margin * cut_h), int(margin * cut_w)
    cy = np.random.randint(0 + margin_y, img_h - margin_y)
    cx = np.random.randint(0 + margin_x, img_w - margin_x)
    yl = int(np.clip(cy - cut_h // 2, 0, img_h))
    yh = int(np.clip(cy + cut_h // 2, 0, img_h))
    xl = int(np.clip(cx - cut_w // 2, 0, img_w))
    xh = int(np.clip(cx + cut_w // 2, 0, img_w))
    return yl, yh, xl, xh


class Mixup:
    """Call-compatible with timm.data.Mixup for the reference's use (mode 'batch'): `x, y = mixup_fn(x, y)`."""

    def __init__(self, mixup_alpha=1.0, cutmix_alpha=0.0, cutmix_minmax=None, prob=1.0, switch_prob=0.5, mode="batch",
                 correct_lam=True, label_smoothing=0.1, num_classes=1000):
        if mode != "batch":
            raise NotImplementedError("deltakd_b200.Mixup implements timm's default mode 'batch' (the reference's)")
        if cutmix_minmax is not None:
            raise NotImplementedError("cutmix_minmax is not used by any exp/*.sh script of the reference")
        self.mixup_alpha, self.cutmix_alpha = mixup_alpha, cutmix_alpha
        self.mix_prob, self.switch_prob = prob, switch_prob
        self.label_smoothing, self.num_classes = label_smoothing, num_classes
        self.correct_lam = correct_lam
        self.mixup_enabled = True

    def _params_per_batch(self):
        lam, use_cutmix = 1.0, False
        if self.mixup_enabled and np.random.rand() < self.mix_prob:
            if self.mixup_alpha > 0.0 and self.cutmix_alpha > 0.0:
                use_cutmix = np.random.rand() < self.switch_prob
                lam_mix = np.random.beta(self.cutmix_alpha, self.cutmix_alpha) if use_cutmix else \
                    np.random.beta(self.mixup_alpha, self.mixup_alpha)
            elif self.mixup_alpha > 0.0:
                lam_mix = np.random.beta(self.mixup_alpha, self.mixup_alpha)
            elif self.cutmix_alpha > 0.0:
                use_cutmix = True
                lam_mix = np.random.beta(self.cutmix_alpha, self.cutmix_alpha)
            else:
                raise ValueError("One of mixup_alpha > 0., cutmix_alpha > 0., cutmix_minmax not None should be true.")
            lam = float(lam_mix)
        return lam, use_cutmix

    def __call__(self, x: torch.Tensor, target: torch.Tensor):
        assert len(x) % 2 == 0, "Batch size should be even when using this"
        _require_cuda(x, target)
        if x.dtype != torch.float32 or not x.is_contiguous() or x.dim() != 4:
            raise ValueError("Mixup expects a contiguous fp32 image batch [B, C, H, W] on the GPU")
        lam, use_cutmix = self._params_per_batch()
        box = (0, 0, 0, 0)
        if lam != 1.0 and use_cutmix:
            box = rand_bbox(x.shape, lam)
            if self.correct_lam:
                lam = 1.0 - (box[1] - box[0]) * (box[3] - box[2]) / float(x.shape[-2] * x.shape[-1])
        lam_t = torch.tensor([lam], dtype=torch.float32, device=x.device)
        if lam != 1.0:
            B, CH, H, W = x.shape
            _lib.call("dkd_mix_batch", x.data_ptr(), B, CH, H, W, lam_t.data_ptr(), int(use_cutmix), *box, _stream())
        return x, MixedLabels(target, lam_t, self.label_smoothing, self.num_classes)
