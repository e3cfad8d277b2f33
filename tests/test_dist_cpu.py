"""CPU, world_size 2, gloo: the loss path shards over the batch with no data-path collective.
Each rank evaluates the (oracle) loss on its shard; averaging the losses / gradients over ranks — what DDP and
the logging all-reduce do — reproduces the single-process full-batch result for every batch-mean loss.
LRKD is NOT shard-invariant by reference semantics (per-rank SVD basis): checked to differ."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deltakd_b200 import dist as D
from deltakd_b200 import heads as H
from oracle import losses as O
from oracle.util import rel_err
from tests.cases import build_case

WORLD = 2


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _loss_and_grads(c, rank=0, world=1):
    heads = H.head_tensors(c.student)
    for p in heads.values():
        p.grad = None
    sh = lambda x: D.shard_batch(x, rank, world)
    s_feats = None if c.s_feats is None else [f.detach()[slice(*D.shard_bounds(c.B, rank, world))].requires_grad_(True) for f in c.s_feats]
    t_feats = sh(c.t_feats)
    outs = c.outputs.detach()[slice(*D.shard_bounds(c.B, rank, world))].requires_grad_(True)
    outs_kd = c.outputs_kd.detach()[slice(*D.shard_bounds(c.B, rank, world))].requires_grad_(True)
    outputs = (outs, outs_kd) if c.kind in ("soft", "hard") else outs
    noise = sh(c.noise)
    loss = O.distillation_loss(c.kind, outputs, sh(c.labels), sh(c.teacher_logits), s_feats, t_feats, heads, c.args,
                               c.alpha, c.tau, noise=noise)
    loss.backward()
    return loss.detach(), outs.grad, s_feats, heads


def _worker(rank, port, name, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        torch.set_num_threads(1)
        c = build_case(name, dtype=torch.float64)
        loss, g_out, s_feats, heads = _loss_and_grads(c, rank, WORLD)
        assert D.world() == (rank, WORLD)
        mean_loss = D.reduce_mean_scalar(loss)
        params = [p for p in heads.values()]
        D.average_gradients(params)
        tmax = D.max_over_ranks([float(rank), 1.0], "cpu")
        assert tmax == [float(WORLD - 1), 1.0]
        # gather the sharded input gradients (each rank owns its samples; DDP divides by world through the loss mean)
        parts = [torch.zeros_like(g_out) for _ in range(WORLD)]
        dist.all_gather(parts, g_out)
        if rank == 0:
            torch.save((mean_loss.item(), torch.cat(parts) / WORLD, {k: p.grad.clone() for k, p in heads.items() if p.grad is not None}), out)
    finally:
        dist.destroy_process_group()


def _run(name):
    import tempfile
    ctx = mp.get_context("spawn")
    port = _free_port()
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "rank0.pt")
        procs = [ctx.Process(target=_worker, args=(r, port, name, path)) for r in range(WORLD)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        return torch.load(path)


@pytest.mark.parametrize("name", ["soft_b8_c1000", "curkd_ep0", "mgd_r05"])
def test_sharded_equals_full_batch(name):
    loss2, g_out2, hg2 = _run(name)
    c = build_case(name, dtype=torch.float64)
    loss1, g_out1, _, heads = _loss_and_grads(c)
    assert abs(loss2 - loss1.item()) <= 1e-12 * abs(loss1.item())
    assert rel_err(g_out2, g_out1) < 1e-12
    for k, g in hg2.items():
        if heads[k].grad is None:   # head unused in this phase: every rank contributed zeros
            assert float(g.abs().max()) == 0.0, k
        else:
            assert rel_err(g, heads[k].grad) < 1e-10, k


def test_lrkd_is_not_shard_invariant():
    """Reference semantics (loss.py:318-321): the SVD is of the LOCAL [B_loc*196, 384] matrix, so the target
    basis — and the loss — depend on how the batch is split.  Kept as is, only documented here."""
    c = build_case("lrkd_r32", dtype=torch.float64)   # B = 3: ragged over 2 ranks -> refused (drop_last semantics)
    with pytest.raises(ValueError):
        D.shard_bounds(c.B, 0, 2)
    heads = H.head_tensors(c.student)
    coef = (c.args.lrkd_alpha, c.args.lrkd_beta, c.args.lrkd_gamma)
    with torch.no_grad():
        full = O.lrkd([f[:2] if f is not None else None for f in c.s_feats], [f[:2] for f in c.t_feats], heads, 32, coef).item()
        halves = [O.lrkd([f[r:r + 1] for f in c.s_feats], [f[r:r + 1] for f in c.t_feats], heads, 32, coef).item() for r in range(2)]
    assert abs(sum(halves) / 2 - full) > 1e-4 * abs(full)


def test_shard_helpers():
    x = torch.arange(8).reshape(8, 1)
    assert D.shard_batch(x, 1, 2).flatten().tolist() == [4, 5, 6, 7]
    assert D.shard_batch([x, None, (x, x)], 0, 4)[2][1].flatten().tolist() == [0, 1]
    assert D.world() == (0, 1)
    assert D.reduce_mean_scalar(torch.tensor(3.0)).item() == 3.0
