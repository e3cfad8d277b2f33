"""Data-parallel plumbing of the loss path: samples are independent, so the batch is
partitioned across ranks with NO data-path collective; the only collectives of a training step are DDP's
gradient all-reduce (reference: tools/train.py:308) and one scalar all-reduce of the loss for logging
(the reference logs the local `loss.item()`, tools/engine.py:71).  torch.distributed (NCCL on the GPUs,
gloo in the CPU tests) carries both.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world() -> tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous equal shards; the reference's loaders drop the ragged tail (dataset/datasets.py:162), so
    `n` must divide evenly — a ragged split would change the batch-mean semantics under DDP averaging."""
    if n % world_size:
        raise ValueError(f"batch {n} does not split evenly over {world_size} ranks (drop_last semantics)")
    per = n // world_size
    return rank * per, (rank + 1) * per


def shard_batch(x, rank: int, world_size: int):
    """Shard a tensor, or a (nested) list/tuple of tensors (None entries kept), along dim 0."""
    if x is None:
        return None
    if isinstance(x, (list, tuple)):
        return type(x)(shard_batch(t, rank, world_size) for t in x)
    lo, hi = shard_bounds(x.shape[0], rank, world_size)
    return x[lo:hi]


def reduce_mean_scalar(loss: torch.Tensor) -> torch.Tensor:
    """Mean over ranks of a 0-dim loss (4-byte all-reduce); detached, for logging."""
    out = loss.detach().clone()
    rank, ws = world()
    if ws > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM)
        out /= ws
    return out


def average_gradients(params) -> None:
    """What DDP does to the student + head gradients (mean over ranks), for harnesses that do not wrap in DDP.
    Parameters without a gradient on this rank (heads unused in a CurKD phase, saliency_attn: SURVEY D6)
    contribute zeros so that every rank joins every all-reduce."""
    rank, ws = world()
    if ws == 1:
        return
    for p in params:
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
        g /= ws
        p.grad = g


def max_over_ranks(values, device) -> list[float]:
    """Element-wise max over ranks of a list of floats (device timings are reported as the max)."""
    rank, ws = world()
    if ws == 1:
        return list(values)
    t = torch.tensor(list(values), device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()
