"""KD heads the loss path reads off the (unwrapped) student model.

The attribute names and shapes are part of the drop-in contract: the reference
attaches them in `load_teacher_student_model` (/root/reference/model/models.py:76-176),
they are checkpointed with the student's state_dict (tools/train.py:352), and
`DistillationLoss.forward` looks them up by name (model/loss.py:89-91, 191, 338,
381, 390, 397, 426).  `attach_distillation_heads` creates the same modules in the
same order, so a given torch seed yields the same initial values.
"""
from __future__ import annotations

import torch
import torch.nn as nn


class SimpleAttention(nn.Module):
    """Saliency scorer for methods 1 and 2 (reference models.py:38-56): a single
    `qk = Linear(dim, 2*dim)`; forward returns, per token, the head-mean of the
    diagonal of softmax(Q K^T * head_dim^-0.5).  The score is computed by the
    `dkd_saliency_score` kernel (no gradient reaches this module in the loss path)."""

    def __init__(self, dim: int, num_heads: int = 8):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qk = nn.Linear(dim, dim * 2, bias=True)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from . import functional as Fn
        return Fn.saliency_score_selfdiag(x, self.qk.weight, self.qk.bias, self.num_heads)


class SimpleCrossAttention(nn.Module):
    """Saliency scorer for method 3 (reference models.py:14-35): separate q / k
    projections; forward(x_query [B,Nq,C], x_key [B,Nk,C]) -> head-mean attention [B,Nq,Nk]."""

    def __init__(self, dim: int, num_heads: int = 8):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.q = nn.Linear(dim, dim, bias=True)
        self.k = nn.Linear(dim, dim, bias=True)

    def forward(self, x_query: torch.Tensor, x_key: torch.Tensor) -> torch.Tensor:
        from . import functional as Fn
        return Fn.saliency_score_cross(x_query, x_key, self.q.weight, self.q.bias,
                                       self.k.weight, self.k.bias, self.num_heads)


class DenoisingNetwork(nn.Module):
    """DiffKD noise predictor (reference models.py:103-121; same sub-module names -> same state_dict keys)."""

    def __init__(self, dims: int):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(dims, dims * 2), nn.GELU(), nn.Linear(dims * 2, dims), nn.Dropout(0.1))
        self.time_embed = nn.Sequential(nn.Linear(1, dims), nn.GELU(), nn.Linear(dims, dims))

    def forward(self, x, t):
        t_emb = self.time_embed(t.float().view(-1, 1))
        return self.net(x + t_emb.unsqueeze(1))


def _generation(dim: int) -> nn.Sequential:
    return nn.Sequential(
        nn.Conv2d(dim, dim, kernel_size=3, padding=1),
        nn.ReLU(inplace=True),
        nn.Conv2d(dim, dim, kernel_size=3, padding=1),
    )


def _linears(n: int, d_in: int, d_out: int) -> nn.ModuleList:
    return nn.ModuleList([nn.Linear(d_in, d_out, bias=True) for _ in range(n)])


def attach_distillation_heads(student_model: nn.Module, teacher_model: nn.Module, args,
                              student_model_name: str = "") -> nn.Module:
    """Attach the per-method heads (same names, shapes, creation order as models.py:76-176)."""
    kind = args.distillation_type.lower()
    ds = getattr(student_model, "embed_dim", None)
    dt = getattr(teacher_model, "embed_dim", None)
    if kind == "vitkd":
        student_model.align2 = _linears(2, ds, dt)
        student_model.align = nn.Linear(ds, dt, bias=True)
        student_model.mask_token = nn.Parameter(torch.zeros(1, 1, dt))
        student_model.generation = _generation(dt)
    elif kind == "lrkd":
        student_model.align = _linears(3, ds, args.lrkd_rank)
    elif "deit" in student_model_name and kind in ("soft", "hard"):
        student_model.set_distilled_training(enable=True)
    elif kind == "diffkd":
        student_model.denoise_fn = DenoisingNetwork(dt)
        student_model.align = _linears(3, ds, dt)
    elif kind == "saliency_mgd":
        student_model.align = nn.Linear(ds, dt, bias=True)
        student_model.mask_token = nn.Parameter(torch.zeros(1, 1, dt))
        student_model.generation = _generation(dt)
        if args.saliency_method in (1, 2):
            student_model.saliency_attn = SimpleAttention(dt, num_heads=8)
        elif args.saliency_method == 3:
            student_model.saliency_attn = SimpleCrossAttention(dt, num_heads=8)
    elif kind == "mgd":
        student_model.align = nn.Linear(ds, dt, bias=True)
        student_model.mask_token = nn.Parameter(torch.zeros(1, 1, dt))
        student_model.generation = _generation(dt)
    elif kind == "curkd":
        student_model.curkd_align_early = _linears(3, ds, dt)
        student_model.curkd_align_mid = _linears(4, ds, dt)
        student_model.curkd_align_last = nn.Linear(ds, dt, bias=True)
        student_model.mask_token = nn.Parameter(torch.zeros(1, 1, dt))
        student_model.generation = _generation(dt)
    elif kind == "wasskd":
        student_model.align_wasskd = _linears(3, ds, dt)
    return student_model


def head_tensors(student_model: nn.Module) -> dict:
    """name -> tensor for every attached head parameter (state_dict naming, `blocks.*` excluded)."""
    return {k: v for k, v in student_model.named_parameters() if not k.startswith("blocks")}
