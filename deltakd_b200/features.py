"""`forward_with_features` — captures every block's MLP output during one forward pass.

Same contract as the reference's helper (/root/reference/model/models.py:181-199): returns
`(model(x), [mlp_out_0, ..., mlp_out_{depth-1}])`, or `(None, None)` for a module without
`.blocks`.  DistributedDataParallel wrappers are unwrapped first (the reference returns
`(None, None)` for them, SURVEY.md D6, which breaks feature KD under DDP).
"""
from __future__ import annotations

import torch.nn as nn


def unwrap(model: nn.Module) -> nn.Module:
    return model.module if isinstance(model, nn.parallel.DistributedDataParallel) else model


def forward_with_features(model: nn.Module, x):
    inner = unwrap(model)
    if not hasattr(inner, "blocks"):
        return None, None
    mlps = [blk.mlp for blk in inner.blocks if hasattr(blk, "mlp")]
    captured = [None] * len(mlps)

    def make_hook(slot):
        def hook(_module, _inp, out):
            captured[slot] = out
        return hook

    handles = [m.register_forward_hook(make_hook(i)) for i, m in enumerate(mlps)]
    try:
        output = model(x)
    finally:
        for h in handles:
            h.remove()
    return output, captured
