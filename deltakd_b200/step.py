"""Step epilogue of the distillation step on B200 — what /root/reference/tools/engine.py:53-69 does after the criterion:

    acc1, acc5 = accuracy(student_logits, targets, topk=(1, 5))              # engine.py:53-56  -> `accuracy`
    optimizer.zero_grad()                                                    # engine.py:58
    loss_scaler(loss, optimizer, clip_grad=clip_grad, parameters=...)        # engine.py:61-62  (timm NativeScaler)
    model_ema.update(student_model)                                          # engine.py:68-69  (timm ModelEma)

`FusedStepEpilogue` keeps the parameters, their gradients, both AdamW moments and the EMA copy in FLAT fp32 buffers
(the modules' parameters become views into them) and runs scaler-unscale + inf/nan check + gradient-norm clipping +
AdamW + EMA + zero_grad as two launches of libdeltakd_sm100 (`dkd_step_epilogue`), with the skip decision, the clip
coefficient, the step count and the loss scale on the device (no host read-back: the whole training step stays
graph-capturable).  There is no CPU path.
"""
from __future__ import annotations

import torch

from . import _lib
from .functional import _dtype_code, _require_cuda, _stream


def accuracy(output: torch.Tensor, target: torch.Tensor, topk=(1, 5)):
    """timm.utils.accuracy (engine.py:53-56): [acc@k0, acc@k1] in percent, as 0-dim device tensors (one kernel)."""
    _require_cuda(output, target)
    if output.dim() != 2 or target.shape != (output.shape[0],):
        raise ValueError("accuracy expects logits [B, C] and int64 targets [B]")
    if len(topk) != 2:
        raise ValueError("topk must have two entries, e.g. (1, 5)")
    B, C = output.shape
    k0, k1 = (min(int(k), C) for k in topk)
    hits = torch.empty(2, dtype=torch.float32, device=output.device)
    out = output.detach()
    if out.dtype == torch.float16:
        out = out.float()
    out = out.contiguous()
    _lib.call("dkd_topk_hits", out.data_ptr(), target.contiguous().data_ptr(), B, C, _dtype_code(out), k0, k1, hits.data_ptr(), _stream())
    acc = hits * (100.0 / B)
    return [acc[0], acc[1]]


class FusedStepEpilogue:
    """AdamW + gradient clipping + loss scaling + EMA over flat buffers.

        epi = FusedStepEpilogue(student.parameters(), lr=5e-4, weight_decay=0.05, clip_grad=None, ema_decay=None)
        ...
        loss = criterion(...)
        epi.backward_and_step(loss)          # == optimizer.zero_grad(); loss_scaler(loss, optimizer, clip_grad, ...); ema.update()

    Parameters with ndim <= 1 (biases, norm scales) and names in `no_decay` form timm's no-weight-decay group
    (timm.optim.create_optimizer: filter_bias_and_bn) and are placed after the decayed ones in the flat buffer.
    `loss_scale=None` (bf16 / fp32 training) keeps the scale at 1; a number enables torch GradScaler's dynamic scaling
    (init `loss_scale`, growth 2 every `growth_interval` good steps, backoff 0.5 on non-finite gradients).
    Build it BEFORE wrapping the model in DistributedDataParallel (DDP then all-reduces the flat gradient views)."""

    def __init__(self, params, lr: float = 5e-4, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.05,
                 clip_grad: float | None = None, ema_decay: float | None = None, loss_scale: float | None = None,
                 growth_factor: float = 2.0, backoff_factor: float = 0.5, growth_interval: int = 2000):
        params = [p for p in params if p.requires_grad]
        if not params:
            raise ValueError("no trainable parameters")
        _require_cuda(*params)
        if any(p.dtype != torch.float32 for p in params):
            raise TypeError("FusedStepEpilogue keeps fp32 master parameters")
        dev = params[0].device
        decay = [p for p in params if p.ndim > 1]
        no_decay = [p for p in params if p.ndim <= 1]
        self.params = decay + no_decay
        pad = lambda n: (n + 3) // 4 * 4                  # every tensor starts 16-byte aligned
        self.n_decay = sum(pad(p.numel()) for p in decay)
        self.n = self.n_decay + sum(pad(p.numel()) for p in no_decay)
        self.flat_p = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(self.n, dtype=torch.float32, device=dev)
        off = 0
        self.offsets = []
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                self.flat_p[off:off + k].copy_(p.reshape(-1))
                p.data = self.flat_p[off:off + k].view_as(p)            # the module now owns a view of the flat buffer
                p.grad = self.flat_g[off:off + k].view_as(p)
                self.offsets.append((off, k))
                off += pad(k)
        self.ema = self.flat_p.clone() if ema_decay else None
        self.lr = torch.tensor([lr], dtype=torch.float32, device=dev)
        self.betas, self.eps, self.weight_decay = betas, eps, weight_decay
        self.clip_grad = float(clip_grad) if clip_grad else 0.0
        self.ema_decay = float(ema_decay) if ema_decay else 0.0
        self.dynamic = loss_scale is not None
        self.growth_factor, self.backoff_factor, self.growth_interval = growth_factor, backoff_factor, int(growth_interval)
        self.state = torch.zeros(8, dtype=torch.float32, device=dev)
        self.state[0] = float(loss_scale) if loss_scale is not None else 1.0
        self.ws = torch.zeros(int(_lib.lib.dkd_step_workspace_bytes()), dtype=torch.uint8, device=dev)

    # ---- the reference's loss_scaler(...) call --------------------------------------------------------------
    def scale(self, loss: torch.Tensor) -> torch.Tensor:
        return loss * self.state[0] if self.dynamic else loss

    def set_lr(self, lr: float) -> None:
        """Scheduler hook (timm schedulers set param_group['lr']): updates the device scalar in place."""
        self.lr.fill_(float(lr))

    def zero_grad(self) -> None:
        """No-op by design: `step()` clears the flat gradient buffer in the same pass that consumed it."""

    def step(self) -> None:
        for p, (off, k) in zip(self.params, self.offsets):     # autograd may have replaced .grad (e.g. set_to_none): re-bind
            g = p.grad
            if g is None:
                p.grad = self.flat_g[off:off + k].view_as(p)
            elif g.data_ptr() != self.flat_g.data_ptr() + 4 * off:
                self.flat_g[off:off + k].copy_(g.reshape(-1))
                p.grad = self.flat_g[off:off + k].view_as(p)
        _lib.call("dkd_step_epilogue", self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.exp_avg.data_ptr(),
                  self.exp_avg_sq.data_ptr(), None if self.ema is None else self.ema.data_ptr(), self.n, self.n_decay,
                  self.lr.data_ptr(), float(self.betas[0]), float(self.betas[1]), float(self.eps), float(self.weight_decay),
                  self.clip_grad, self.ema_decay, int(self.dynamic), float(self.growth_factor), float(self.backoff_factor),
                  self.growth_interval, 1, self.state.data_ptr(), self.ws.data_ptr(), self.ws.numel(), _stream())

    def backward_and_step(self, loss: torch.Tensor) -> None:
        self.scale(loss).backward()
        self.step()

    # ---- views for logging / checkpoints ---------------------------------------------------------------------
    @property
    def grad_norm(self) -> torch.Tensor:
        return self.state[4]

    @property
    def loss_scale(self) -> torch.Tensor:
        return self.state[0]

    @property
    def skipped(self) -> torch.Tensor:
        return self.state[3]

    def ema_tensors(self):
        """EMA copy as a list of tensors shaped like the parameters (timm ModelEma.module's values)."""
        if self.ema is None:
            return None
        return [self.ema[off:off + k].view_as(p) for p, (off, k) in zip(self.params, self.offsets)]

    def state_dict(self):
        return {"exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq, "ema": self.ema, "state": self.state, "lr": self.lr}
