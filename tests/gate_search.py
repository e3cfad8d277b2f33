"""ReLU-gate-aware gradient comparison for the masked-generation losses.

The generator's ReLU makes the gradient discontinuous where a pre-activation is ~0.  Two correct
implementations with different rounding (fp32 CPU, TF32 cuDNN, bf16x3 tcgen05 ...) can gate such an element
differently; like mask selection "given the same scores", gradient parity is defined given the same gate.
`best_gate_oracle` evaluates the fp64 oracle with the reference gate and then greedily flips only the
AMBIGUOUS gates (|pre-activation| < tau) while that brings the oracle gradient closer to the implementation's.
"""
from __future__ import annotations

import torch


def best_gate_oracle(eval_fn, ours: dict, tau: float = 3e-5, max_flips: int = 64):
    """eval_fn(probe) -> (loss, {name: grad}) runs the fp64 oracle; ours = {name: grad tensor (cpu, fp64)}.
    Returns (loss, grads, n_ambiguous, n_flipped)."""
    def err(g):
        return max(float((ours[k] - g[k]).norm() / g[k].norm().clamp_min(1e-300)) for k in ours)

    probe = {}
    loss, grads = eval_fn(probe)
    pre = probe["pre"]
    base_gate = (pre > 0).to(pre.dtype)
    amb = (pre.abs() < tau).nonzero()
    best, best_e = (loss, grads), err(grads)
    flipped = 0
    if amb.shape[0] == 0 or amb.shape[0] > max_flips:
        return best[0], best[1], int(amb.shape[0]), 0
    order = sorted(range(amb.shape[0]), key=lambda i: float(pre[tuple(amb[i])].abs()))
    gate = base_gate.clone()
    for i in order:
        idx = tuple(amb[i])
        gate[idx] = 1 - gate[idx]
        l2, g2 = eval_fn({"gate": gate})
        e2 = err(g2)
        if e2 < best_e:
            best, best_e, flipped = (l2, g2), e2, flipped + 1
        else:
            gate[idx] = 1 - gate[idx]
    return best[0], best[1], int(amb.shape[0]), flipped


def case_gate_oracle(name: str, c, tau: float = 3e-5, max_flips: int = 48):
    """fp64 oracle of golden case `name` compared, gate-aware, with the gradients held by the GPU case `c`
    (tests.cases.build_case(..., device='cuda') after backward).  Returns (loss, grads, ours, n_amb, n_flip);
    keys: head parameter names and 's<i>' for student feature i."""
    from oracle import losses as O
    from tests.cases import build_case
    from deltakd_b200 import heads as H

    ours = {k: p.grad.detach().double().cpu() for k, p in H.head_tensors(c.student).items() if p.grad is not None}
    for i, f in enumerate(c.s_feats):
        if f is not None and f.grad is not None:
            ours[f"s{i}"] = f.grad.detach().double().cpu()

    def eval_fn(probe):
        o = build_case(name, dtype=torch.float64)
        oh = H.head_tensors(o.student)
        l = O.distillation_loss(o.kind, o.outputs, o.labels, o.teacher_logits, o.s_feats, o.t_feats, oh, o.args,
                                o.alpha, o.tau, noise=o.noise, probe=probe)
        l.backward()
        g = {k: v.grad for k, v in oh.items() if v.grad is not None}
        for i, f in enumerate(o.s_feats):
            if f.grad is not None:
                g[f"s{i}"] = f.grad
        return l, g

    loss, grads, n_amb, n_flip = best_gate_oracle(eval_fn, {k: v for k, v in ours.items()}, tau=tau, max_flips=max_flips)
    return loss, grads, ours, n_amb, n_flip
