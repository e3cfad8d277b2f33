"""`forward_with_features` — captures block MLP outputs during one forward pass — and `FrozenTeacher`.

Same contract as the reference's helper (/root/reference/model/models.py:181-199): returns
`(model(x), [mlp_out_0, ..., mlp_out_{depth-1}])`, or `(None, None)` for a module without
`.blocks`.  DistributedDataParallel wrappers are unwrapped first (the reference returns
`(None, None)` for them, SURVEY.md D6, which breaks feature KD under DDP).

SURVEY 8f rank 1: the reference hooks all 12 blocks of student AND teacher every step although a
distillation type reads 1-4 of them.  `layers=` restricts the hooks to the blocks that will be read
(`needed_layers(distillation_type, args)`); the returned list keeps its length (None at the blocks that
were not captured) so `features[i]` indexing is unchanged.  The loss kernels take the captured tensors in
place (token offsets `[:, 1:]` / `[:, 2:]` are kernel arguments, not slices), so nothing is copied.
"""
from __future__ import annotations

import torch
import torch.nn as nn


def unwrap(model: nn.Module) -> nn.Module:
    return model.module if isinstance(model, nn.parallel.DistributedDataParallel) else model


def needed_layers(distillation_type: str, args=None, depth: int = 12):
    """Block indices whose MLP outputs DistillationLoss.forward reads for this type (reference model/loss.py:80-236,
    251-451), or None for "all" (unknown types keep the reference behaviour)."""
    kind = (distillation_type or "").lower()
    last = depth - 1
    if kind in ("none", "soft", "hard"):
        return ()
    if kind in ("vitkd", "diffkd"):
        return (0, 1, last)                       # loss.py:258-263, :112-121
    if kind == "lrkd":
        return (0, 1, 11, last)                   # student -1, teacher 11 (loss.py:91, 98)
    if kind in ("mgd", "saliency_mgd"):
        return (last,)                            # loss.py:426-428, :338-339
    if kind == "wasskd":
        return (0, 1, 2)                          # loss.py:188-194
    if kind == "curkd":
        epoch = getattr(args, "current_epoch", None)
        if epoch is None:
            return (0, 1, 2, 3, 4, 5, 6, 11)
        return (0, 1, 2) if epoch < 100 else (3, 4, 5, 6) if epoch < 151 else (11,)   # loss.py:376-397
    return None


def forward_with_features(model: nn.Module, x, layers=None):
    """(model(x), features).  `layers=None` hooks every block (reference behaviour); an iterable of block indices
    (negative = from the end) hooks only those — the other entries of the returned list are None."""
    inner = unwrap(model)
    if not hasattr(inner, "blocks"):
        return None, None
    mlps = [blk.mlp for blk in inner.blocks if hasattr(blk, "mlp")]
    captured = [None] * len(mlps)
    if layers is None:
        wanted = range(len(mlps))
    else:
        wanted = sorted({int(i) % len(mlps) for i in layers}) if len(mlps) else ()

    def make_hook(slot):
        def hook(_module, _inp, out):
            captured[slot] = out
        return hook

    handles = [mlps[i].register_forward_hook(make_hook(i)) for i in wanted]
    try:
        output = model(x)
    finally:
        for h in handles:
            h.remove()
    return output, captured


class FrozenTeacher(nn.Module):
    """Frozen teacher in a half-width dtype (the reference runs the fp32 teacher twice per step,
    loss.py:44-52).  Wraps any teacher exposing `.blocks[i].mlp` / `.embed_dim`: parameters are cast ONCE to `dtype`
    (no per-step autocast weight casts), gradients are off, the module stays in eval mode, and the forward runs under
    autocast so LayerNorm / softmax keep their fp32 internals.  Drop-in for `teacher_model` in
    `DistillationLoss(base_criterion, teacher_model, ...)`: `forward_with_features` finds `.blocks` through it."""

    def __init__(self, teacher: nn.Module, dtype: torch.dtype = torch.bfloat16):
        super().__init__()
        self.inner = teacher.eval()
        for p in self.inner.parameters():
            p.requires_grad_(False)
        if dtype is not None and next(self.inner.parameters(), torch.empty(0)).is_cuda:
            self.inner.to(dtype)
        self.dtype = dtype
        self.embed_dim = getattr(teacher, "embed_dim", None)

    @property
    def blocks(self):
        return self.inner.blocks

    def train(self, mode: bool = True):   # a frozen teacher never leaves eval mode (reference: models.py:72-74)
        return super().train(False)

    @torch.no_grad()
    def forward(self, x):
        if self.dtype is None or not x.is_cuda:
            return self.inner(x)
        with torch.autocast("cuda", dtype=self.dtype):
            return self.inner(x.to(self.dtype))
