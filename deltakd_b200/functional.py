"""torch.autograd.Function wrappers over the C ABI (include/deltakd.h).

PyTorch is plumbing here: it owns device memory and streams; every number is
produced by libdeltakd_sm100's kernels.  All losses are fused forward+backward:
the forward launch also writes the gradients for d(loss)=1 and `backward` only
rescales them on the device when the incoming grad is not 1
(`dkd_scale_if_not_one`, no host sync).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

_DT = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("deltakd_b200 runs on CUDA (sm_100) tensors only; there is no CPU path")


def _dtype_code(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported dtype {t.dtype}: deltakd_b200 takes float32 or bfloat16") from None


# zero-initialised scratch, cached per (device, tag, size); kernels leave their counters reset
_WS: dict = {}


def _workspace(device: torch.device, tag: str, nbytes: int) -> torch.Tensor:
    key = (device.index, tag)
    ws = _WS.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(nbytes, 1024), dtype=torch.uint8, device=device)
        _WS[key] = ws
    return ws


def _rescale_(g: torch.Tensor, grad_out: torch.Tensor) -> torch.Tensor:
    """g *= grad_out on the device, skipped when grad_out == 1 (decided on the device)."""
    if g is None:
        return None
    go = grad_out
    if go.dtype != torch.float32 or not go.is_contiguous():
        go = go.to(torch.float32).contiguous()
    _lib.call("dkd_scale_if_not_one", _ptr(g), g.numel(), _dtype_code(g), _ptr(go), _stream())
    return g


# --------------------------------------------------------------------------- logit losses
class _LogitKD(torch.autograd.Function):
    @staticmethod
    def forward(ctx, outputs, outputs_kd, teacher_logits, labels, label_kind, kd_kind, smoothing, alpha, tau, parts_out):
        ref = outputs if outputs is not None else outputs_kd
        B, Cn = ref.shape
        dt = _dtype_code(ref)
        need_g0 = label_kind >= 0 and ctx.needs_input_grad[0]
        need_g1 = kd_kind != 0 and ctx.needs_input_grad[1]
        g0 = torch.empty_like(outputs) if need_g0 else None
        g1 = torch.empty_like(outputs_kd) if need_g1 else None
        loss3 = torch.empty(3, dtype=torch.float32, device=ref.device)
        nbytes = _lib.lib.dkd_logit_kd_workspace_bytes(B)
        ws = _workspace(ref.device, "logit_kd", nbytes)
        _lib.call("dkd_logit_kd_fwdbwd", _ptr(outputs), _ptr(outputs_kd), _ptr(teacher_logits), _ptr(labels),
                  label_kind, kd_kind, B, Cn, dt, float(smoothing), float(alpha), float(tau),
                  _ptr(g0), _ptr(g1), _ptr(loss3), _ptr(ws), ws.numel(), _stream())
        ctx.grads = (g0, g1)
        parts_out.append(loss3)  # {total, base, kd}, detached side channel for logging
        return loss3[0]

    @staticmethod
    def backward(ctx, grad_total):
        g0, g1 = ctx.grads
        ctx.grads = None
        return (_rescale_(g0, grad_total), _rescale_(g1, grad_total)) + (None,) * 8


def logit_kd_loss(outputs, outputs_kd, teacher_logits, labels, *, kd_kind: str, smoothing: float = 0.1,
                  alpha: float = 0.0, tau: float = 1.0, return_parts: bool = False):
    """Fused base CE (+ soft / hard KD) on logits.

    labels: float [B,C] soft targets -> SoftTargetCrossEntropy; int64 [B] -> LabelSmoothingCrossEntropy(smoothing);
    None -> KD term only (returns alpha*kd).  kd_kind in {"none","soft","hard"}.
    Returns the 0-dim fp32 total `base*(1-alpha) + kd*alpha` (or base alone for "none").
    """
    kk = {"none": 0, "soft": 1, "hard": 2}[kd_kind]
    ref = outputs if outputs is not None else outputs_kd
    _require_cuda(outputs, outputs_kd, teacher_logits, labels)
    if ref.dim() != 2:
        raise ValueError(f"logits must be [B, C], got {tuple(ref.shape)}")
    if labels is None:
        lk = -1
        outputs = None
    elif labels.dtype == torch.int64:
        lk = 1
        if labels.shape != (ref.shape[0],):
            raise ValueError(f"int labels must be [B], got {tuple(labels.shape)}")
    else:
        lk = 0
        if labels.shape != ref.shape:
            raise ValueError(f"soft labels must be [B, C] like the logits, got {tuple(labels.shape)}")
        if labels.dtype != ref.dtype:
            labels = labels.to(ref.dtype)
    if kk:
        if outputs_kd is None or teacher_logits is None:
            raise ValueError("KD needs outputs_kd and teacher_logits")
        if outputs_kd.shape != ref.shape or teacher_logits.shape != ref.shape:
            raise ValueError("outputs_kd / teacher_logits must have the shape of outputs")
        if teacher_logits.dtype != ref.dtype:
            teacher_logits = teacher_logits.to(ref.dtype)
        if outputs_kd.dtype != ref.dtype:
            raise TypeError("outputs and outputs_kd must share a dtype")
        teacher_logits = teacher_logits.detach().contiguous()
        outputs_kd = outputs_kd.contiguous()
    else:
        outputs_kd = teacher_logits = None
    if outputs is not None:
        outputs = outputs.contiguous()
    if labels is not None:
        labels = labels.detach().contiguous()
    parts = []
    total = _LogitKD.apply(outputs, outputs_kd, teacher_logits, labels, lk, kk, smoothing, alpha, tau, parts)
    return (total, parts[0]) if return_parts else total


# --------------------------------------------------------------------------- masking
def mask_rank(score: torch.Tensor, len_keep: int, want_shuffle: bool = True):
    """(mask fp32 [B,L], ids_restore int64 [B,L], ids_shuffle int64 [B,L] | None) from fp32 scores."""
    _require_cuda(score)
    if score.dim() != 2:
        raise ValueError("score must be [B, L]")
    score = score.detach().to(torch.float32).contiguous()
    B, L = score.shape
    mask = torch.empty(B, L, dtype=torch.float32, device=score.device)
    ids_restore = torch.empty(B, L, dtype=torch.int64, device=score.device)
    ids_shuffle = torch.empty(B, L, dtype=torch.int64, device=score.device) if want_shuffle else None
    _lib.call("dkd_mask_rank", _ptr(score), B, L, int(len_keep), _ptr(mask), _ptr(ids_restore),
              _ptr(ids_shuffle), _stream())
    return mask, ids_restore, ids_shuffle
