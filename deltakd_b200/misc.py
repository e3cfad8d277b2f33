"""Token masking for masked-generation distillation — B200 versions of the reference's
`random_masking` / `saliency_masking` (/root/reference/model/misc.py:5-32, 38-165).

Both keep the reference's call signatures and return tuples.  The double argsort is
replaced by one rank-by-counting kernel (`dkd_mask_rank`); ties rank lower-index-first
(torch.argsort(stable=False) leaves them unspecified; CUDA radix sort behaves this way).
"""
from __future__ import annotations

import torch

from . import functional as Fn


def len_keep_of(L: int, mask_ratio: float) -> int:
    """Python float arithmetic then truncation, exactly as misc.py:12 (e.g. r=0.3 -> 137)."""
    return int(L * (1 - mask_ratio))


def random_masking(x: torch.Tensor, mask_ratio: float, noise: torch.Tensor | None = None):
    """x [N,L,D] -> (x_keep [N,len_keep,D], mask [N,L] 0=keep/1=masked, ids_restore, ids_masked).

    `noise` defaults to torch.rand(N, L, device=x.device) — the same draw from torch's generator
    as misc.py:14 — and can be passed in for reproducible parity checks."""
    N, L, D = x.shape
    len_keep = len_keep_of(L, mask_ratio)
    if noise is None:
        noise = torch.rand(N, L, device=x.device)
    mask, ids_restore, ids_shuffle = Fn.mask_rank(noise, len_keep)
    ids_keep = ids_shuffle[:, :len_keep]
    ids_masked = ids_shuffle[:, len_keep:L]
    # API-parity output only: the fused loss path never materialises x_keep (it uses `mask`)
    x_keep = torch.gather(x, dim=1, index=ids_keep.unsqueeze(-1).expand(-1, -1, D))
    return x_keep, mask, ids_restore, ids_masked


def saliency_scores(student_model, teacher_feat: torch.Tensor, method: int) -> torch.Tensor:
    """fp32 [B, 196] scores whose ascending order picks the kept tokens (misc.py:62-162).
    teacher_feat is the teacher's last-block feature [B,198,D] with CLS, DIST at 0, 1."""
    attn = student_model.saliency_attn
    if method == 1:
        return attn(teacher_feat[:, 2:])
    if method == 2:
        return Fn.saliency_score_cls_row(teacher_feat, attn.qk.weight, attn.qk.bias, attn.num_heads)
    if method == 3:
        w = attn(teacher_feat[:, :1], teacher_feat[:, 2:])
        return w.squeeze(1) if w.dim() == 3 and w.size(1) == 1 else w
    raise ValueError(f"Invalid saliency masking method: {method}")


def saliency_masking(student_model, teacher_feat, student_feat, mask_ratio, method):
    """(x_keep, mask, ids_restore): keeps the LOWEST-score tokens (misc.py:72-81)."""
    with torch.no_grad():
        score = saliency_scores(student_model, teacher_feat, method)
    L = score.shape[1]
    len_keep = len_keep_of(L, mask_ratio)
    mask, ids_restore, ids_shuffle = Fn.mask_rank(score, len_keep)
    D = student_feat.shape[-1]
    x_keep = torch.gather(student_feat, dim=1, index=ids_shuffle[:, :len_keep].unsqueeze(-1).expand(-1, -1, D))
    return x_keep, mask, ids_restore
