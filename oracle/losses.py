"""TEST INFRASTRUCTURE ONLY — CPU oracle (plain torch, fp32 or fp64) for the
DeltaKD distillation-loss hot path.

Every function restates one piece of /root/reference/model/loss.py,
model/misc.py or model/models.py in closed form (cited per function) and is
pinned against golden vectors produced by importing the unmodified reference
(oracle/make_golden.py -> tests/golden/*.npz, checked by
tests/test_oracle_golden.py).  The Sinkhorn term is the exception: its
arithmetic lives in the absent third-party `geomloss`, see oracle/sinkhorn.py
("parity unpinned").

Conventions: `s_feats[i]` is the student's block-i MLP output [B,197,Ds] with
the CLS token at index 0; `t_feats[i]` the teacher's [B,198,Dt] with CLS, DIST
at 0, 1 (reference models.py:181-199).  `heads` is a dict of plain tensors
named like the attributes models.py:76-176 attaches to the student.
Gradients come from torch autograd over these closed forms.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

from oracle.sinkhorn import sinkhorn_divergence


# --------------------------------------------------------------------------- base CE
def soft_target_ce(z: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """timm SoftTargetCrossEntropy: mean_b sum_c -y*log_softmax(z) (loss.py:35,247)."""
    return (-(y * F.log_softmax(z, dim=-1)).sum(-1)).mean()


def label_smoothing_ce(z: torch.Tensor, labels: torch.Tensor, smoothing: float = 0.1) -> torch.Tensor:
    """timm LabelSmoothingCrossEntropy (loss.py:249): (1-e)*nll + e*mean_c(-logp), batch mean."""
    logp = F.log_softmax(z, dim=-1)
    nll = -logp.gather(1, labels.view(-1, 1)).squeeze(1)
    return ((1.0 - smoothing) * nll + smoothing * (-logp.mean(-1))).mean()


def base_criterion_for(args) -> str:
    """call_base_loss (loss.py:244-249): truthiness of mixup/cutmix/cutmix_minmax picks the soft-target CE."""
    active = args.mixup > 0 or args.cutmix > 0.0 or args.cutmix_minmax
    return "soft_target" if active else "label_smoothing"


def base_loss(z, labels, kind: str, smoothing: float = 0.1):
    return soft_target_ce(z, labels) if kind == "soft_target" else label_smoothing_ce(z, labels, smoothing)


# --------------------------------------------------------------------------- logit KD
def soft_kd(z_s: torch.Tensor, z_t: torch.Tensor, T: float) -> torch.Tensor:
    """loss.py:57-64: sum p_t (log p_t - log p_s) * T^2 / numel, p = softmax(z/T)."""
    lp_s = F.log_softmax(z_s / T, dim=1)
    lp_t = F.log_softmax(z_t / T, dim=1)
    return F.kl_div(lp_s, lp_t, reduction="sum", log_target=True) * (T * T) / z_s.numel()


def hard_kd(z_s: torch.Tensor, z_t: torch.Tensor) -> torch.Tensor:
    """loss.py:66-67: CE(z_s, argmax z_t); argmax = first maximal index."""
    idx = torch.from_numpy(np.argmax(z_t.detach().cpu().numpy(), axis=1)).to(z_s.device)
    return F.nll_loss(F.log_softmax(z_s, dim=1), idx)


# --------------------------------------------------------------------------- masking
def len_keep_of(L: int, mask_ratio: float) -> int:
    """misc.py:12 — Python float arithmetic then truncation."""
    return int(L * (1 - mask_ratio))


def mask_from_scores(score: torch.Tensor, len_keep: int):
    """misc.py:17-30 restated as rank-by-counting (integer work in numpy):
    ids_restore[b,i] = rank of score[b,i] ascending (ties -> lower index first,
    torch CPU argsort behaviour), mask = 1.0 where rank >= len_keep."""
    sc = score.detach().cpu().numpy()
    ids_shuffle = np.argsort(sc, axis=1, kind="stable")
    ids_restore = np.argsort(ids_shuffle, axis=1, kind="stable")
    mask = (ids_restore >= len_keep).astype(np.float32)
    return (torch.from_numpy(mask), torch.from_numpy(ids_restore.astype(np.int64)),
            torch.from_numpy(ids_shuffle.astype(np.int64)))


def random_masking(x: torch.Tensor, mask_ratio: float, noise: torch.Tensor):
    """misc.py:5-32 with the noise tensor passed in (the reference draws torch.rand(N,L))."""
    N, L, D = x.shape
    lk = len_keep_of(L, mask_ratio)
    mask, ids_restore, ids_shuffle = mask_from_scores(noise, lk)
    ids_keep = ids_shuffle[:, :lk]
    x_keep = torch.gather(x, 1, ids_keep.unsqueeze(-1).expand(-1, -1, D))
    return x_keep, mask.to(x.dtype), ids_restore, ids_shuffle[:, lk:]


# --------------------------------------------------------------------------- saliency scores
def saliency_score(method: int, t_feat: torch.Tensor, heads: dict, num_heads: int = 8) -> torch.Tensor:
    """Scores whose ascending argsort picks the kept tokens (misc.py:62-162, models.py:14-56).
    method 1: head-mean diagonal of softmax(QK^T*scale) over the 196 patches;
    method 2: CLS-query row over [CLS]+patches, head-mean, CLS column dropped;
    method 3: cross attention CLS -> patches (separate q / k projections)."""
    B, _, D = t_feat.shape
    hd = D // num_heads
    scale = hd ** -0.5
    if method == 1:
        x = t_feat[:, 2:]
        qk = x @ heads["saliency_attn.qk.weight"].t() + heads["saliency_attn.qk.bias"]
        L = x.shape[1]
        q = qk[..., :D].reshape(B, L, num_heads, hd).permute(0, 2, 1, 3)
        k = qk[..., D:].reshape(B, L, num_heads, hd).permute(0, 2, 1, 3)
        attn = torch.softmax((q @ k.transpose(-2, -1)) * scale, dim=-1)
        return attn.mean(1).diagonal(dim1=-2, dim2=-1)
    if method == 2:
        x = torch.cat([t_feat[:, :1], t_feat[:, 2:]], dim=1)
        qk = x @ heads["saliency_attn.qk.weight"].t() + heads["saliency_attn.qk.bias"]
        L = x.shape[1]
        q = qk[..., :D].reshape(B, L, num_heads, hd).permute(0, 2, 1, 3)
        k = qk[..., D:].reshape(B, L, num_heads, hd).permute(0, 2, 1, 3)
        logits = (q[:, :, 0:1] @ k.transpose(-2, -1)) * scale
        attn = torch.softmax(logits, dim=-1).mean(1).squeeze(1)
        return attn[:, 1:]
    if method == 3:
        cls, patches = t_feat[:, :1], t_feat[:, 2:]
        q = cls @ heads["saliency_attn.q.weight"].t() + heads["saliency_attn.q.bias"]
        k = patches @ heads["saliency_attn.k.weight"].t() + heads["saliency_attn.k.bias"]
        L = patches.shape[1]
        q = q.reshape(B, 1, num_heads, hd).permute(0, 2, 1, 3)
        k = k.reshape(B, L, num_heads, hd).permute(0, 2, 1, 3)
        attn = torch.softmax((q @ k.transpose(-2, -1)) * scale, dim=-1)
        return attn.mean(1).squeeze(1)
    raise ValueError(f"Invalid saliency masking method: {method}")


# --------------------------------------------------------------------------- feature losses
def _align(s: torch.Tensor, W: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    return s[:, 1:] @ W.t() + b


def curkd_hidden(s_feats, t_feats, heads, epoch: int) -> torch.Tensor:
    """curkd_loss early/mid (loss.py:376-393): sum-MSE over the selected layers * 4e-5/(nL*B)."""
    B = next(f for f in s_feats if f is not None).shape[0]
    if epoch < 100:
        layers, name, off = range(3), "curkd_align_early", 0
    else:
        layers, name, off = range(3, 7), "curkd_align_mid", 3
    total = 0.0
    for i in layers:
        y = _align(s_feats[i], heads[f"{name}.{i - off}.weight"], heads[f"{name}.{i - off}.bias"])
        total = total + ((y - t_feats[i][:, 2:]) ** 2).sum()
    return total / float(len(layers)) / B * 4e-5


def generator(x_tokens: torch.Tensor, heads: dict, probe: dict | None = None) -> torch.Tensor:
    """student.generation = Conv3x3 -> ReLU -> Conv3x3 on the hw x hw token grid (models.py:148-151);
    the [B,N,D] token layout is NHWC, the reference permutes to NCHW (loss.py:444-446).
    `probe` (tests only): probe["pre"] receives the ReLU pre-activations [B,D,hw,hw]; if probe["gate"] is
    set it replaces the ReLU's own 0/1 gate (used to study gate flips of near-zero pre-activations)."""
    B, N, D = x_tokens.shape
    hw = int(N ** 0.5)
    img = x_tokens.reshape(B, hw, hw, D).permute(0, 3, 1, 2)
    pre = F.conv2d(img, heads["generation.0.weight"], heads["generation.0.bias"], padding=1)
    if probe is not None:
        probe["pre"] = pre.detach()
    h = pre * probe["gate"] if (probe is not None and probe.get("gate") is not None) else F.relu(pre)
    g = F.conv2d(h, heads["generation.2.weight"], heads["generation.2.bias"], padding=1)
    return g.flatten(2).transpose(1, 2)


def masked_generation_sse(x_aligned, mask, t_patch, heads, probe=None) -> torch.Tensor:
    """Shared core of mgd / saliency_mgd / curkd-late / vitkd-gen:
    where(mask, mask_token, x) -> generator -> sum m*(g - t)^2 (loss.py:436-450 et al.)."""
    m = mask.unsqueeze(-1).to(x_aligned.dtype)
    x_m = x_aligned * (1 - m) + heads["mask_token"].reshape(1, 1, -1) * m
    g = generator(x_m, heads, probe)
    return (m * (g - t_patch) ** 2).sum()


def mgd(s_feats, t_feats, heads, mgd_alpha: float, mask_ratio: float, noise, probe=None) -> torch.Tensor:
    """mgd_loss (loss.py:422-451): mean over B*N*D, times mgd_alpha."""
    x = _align(s_feats[-1], heads["align.weight"], heads["align.bias"])
    _, mask, _, _ = random_masking(x, mask_ratio, noise)
    t = t_feats[-1][:, 2:]
    return masked_generation_sse(x, mask, t, heads, probe) / t.numel() * mgd_alpha


def saliency_mgd(s_feats, t_feats, heads, mask_ratio: float, method: int, probe=None) -> torch.Tensor:
    """saliency_mgd_loss (loss.py:335-360): mask = highest-score tokens; mean-MSE * 4."""
    x = _align(s_feats[-1], heads["align.weight"], heads["align.bias"])
    with torch.no_grad():
        score = saliency_score(method, t_feats[-1], heads)
    mask, _, _ = mask_from_scores(score, len_keep_of(x.shape[1], mask_ratio))
    t = t_feats[-1][:, 2:]
    return masked_generation_sse(x, mask.to(x.dtype), t, heads, probe) / t.numel() * 4


def curkd_late(s_feats, t_feats, heads, noise, probe=None) -> torch.Tensor:
    """curkd_loss epoch>=151 (loss.py:394-420): layer 11, ratio fixed 0.5, sum-MSE * 5e-5/B."""
    x = _align(s_feats[11], heads["curkd_align_last.weight"], heads["curkd_align_last.bias"])
    _, mask, _, _ = random_masking(x, 0.5, noise)
    B = x.shape[0]
    return masked_generation_sse(x, mask, t_feats[11][:, 2:], heads, probe) / B * 5e-5


def vitkd(s_feats, t_feats, heads, noise, alpha_v=0.00003, beta_v=0.000003, lambda_v=0.5, probe=None) -> torch.Tensor:
    """vitkd_loss (loss.py:251-311): 2-layer mimic (align2) + masked generation on the last layer."""
    B = s_feats[0].shape[0]
    lr = 0.0
    for i in range(2):
        y = _align(s_feats[i], heads[f"align2.{i}.weight"], heads[f"align2.{i}.bias"])
        lr = lr + ((y - t_feats[i][:, 2:]) ** 2).sum()
    lr = lr / B * alpha_v
    x = _align(s_feats[-1], heads["align.weight"], heads["align.bias"])
    _, mask, _, _ = random_masking(x, lambda_v, noise)
    gen = masked_generation_sse(x, mask, t_feats[-1][:, 2:], heads, probe) / B * beta_v / lambda_v
    return lr + gen


def lrkd_targets(t_patch: torch.Tensor, rank: int):
    """Rank-k target A = U_k diag(S_k) = T V_k of T=[B*196, Dt] (loss.py:318-324).  Returns (A, V_k, S_k).
    Column signs are LAPACK-arbitrary; callers align signs before comparing."""
    T = t_patch.reshape(-1, t_patch.shape[-1])
    U, S, Vh = torch.linalg.svd(T, full_matrices=False)
    return U[:, :rank] * S[:rank], Vh[:rank].t(), S[:rank]


def lrkd(s_feats, t_feats, heads, rank: int, coef, signs=None) -> torch.Tensor:
    """lrkd branch (loss.py:80-103, 314-330): student (0,1,-1) vs teacher (0,1,11), mean-MSE per layer.
    `signs` (optional list of 3 [rank] +-1 tensors) flips target columns for sign-aligned comparison."""
    total = 0.0
    for j, (si, ti) in enumerate(((0, 0), (1, 1), (-1, 11))):
        A, _, _ = lrkd_targets(t_feats[ti][:, 2:].detach(), rank)
        if signs is not None:
            A = A * signs[j]
        sp = _align(s_feats[si], heads[f"align.{j}.weight"], heads[f"align.{j}.bias"]).reshape(-1, rank)
        total = total + coef[j] * ((A - sp) ** 2).mean()
    return total


def wass_l1(s_feats, t_feats, heads) -> torch.Tensor:
    """wasskd 'l1' (loss.py:187-199): per-(b,channel) sort over tokens, mean |diff|, average of 3 layers."""
    total = 0.0
    for i in range(3):
        a = _align(s_feats[i], heads[f"align_wasskd.{i}.weight"], heads[f"align_wasskd.{i}.bias"])
        sa, _ = torch.sort(a, dim=1)
        st, _ = torch.sort(t_feats[i][:, 2:], dim=1)
        total = total + (sa - st).abs().mean()
    return total / 3.0


def wass_sinkhorn(s_feats, t_feats, heads, blur: float = 0.05) -> torch.Tensor:
    """wasskd 'sinkhorn' (loss.py:200-225): sum_b S(a[b], t[b]) / (B*N), average of 3 layers."""
    total = 0.0
    for i in range(3):
        a = _align(s_feats[i], heads[f"align_wasskd.{i}.weight"], heads[f"align_wasskd.{i}.bias"])
        t = t_feats[i][:, 2:]
        if a.shape != t.shape:
            raise ValueError(f"Feature shape mismatch at layer {i}: aligned student {a.shape} vs teacher {t.shape}")
        B, N, _ = a.shape
        layer = 0.0
        for b in range(B):
            layer = layer + sinkhorn_divergence(a[b], t[b], blur=blur)
        total = total + layer / (B * N)
    return total / 3.0


def denoise_fn(x: torch.Tensor, t: torch.Tensor, heads: dict) -> torch.Tensor:
    """student.denoise_fn = DenoisingNetwork(Dt) (models.py:103-121): time embedding Linear(1,D)-GELU-Linear(D,D)
    added to every token, then Linear(D,2D)-GELU-Linear(2D,D)-Dropout(0.1).  Dropout is evaluated in eval mode
    (identity): its mask is a torch RNG draw, outside what parity can pin."""
    te = F.linear(t.to(x.dtype).view(-1, 1), heads["denoise_fn.time_embed.0.weight"], heads["denoise_fn.time_embed.0.bias"])
    te = F.linear(F.gelu(te), heads["denoise_fn.time_embed.2.weight"], heads["denoise_fn.time_embed.2.bias"])
    h = F.linear(x + te.unsqueeze(1), heads["denoise_fn.net.0.weight"], heads["denoise_fn.net.0.bias"])
    return F.linear(F.gelu(h), heads["denoise_fn.net.2.weight"], heads["denoise_fn.net.2.bias"])


def diffkd(s_feats, t_feats, heads, t: torch.Tensor, noises, T: int = 8) -> torch.Tensor:
    """diffkd branch (loss.py:105-155): layers (0, 1, -1); L2-normalised tokens; noise = randn * sigma_t with
    sigma_t = (1 - cos(pi t / T)) * (0.3 if t < T/2 else 0.7); per layer mse(denoise_fn(t_hat + noise, t), noise)
    + mean_b(1 / (sigma_t^2 + 1e-8)) * mse(s_hat, t_hat); mean of the 3 layers, times lambda_feat = 5e-5.
    `t` (int [B], the reference's torch.randint draw) and `noises` (3 x randn_like draws, BEFORE the sigma scaling)
    are inputs so that parity does not depend on RNG streams."""
    dt = t_feats[0].dtype
    sigma_max = torch.where(t < T // 2, torch.tensor(0.3), torch.tensor(0.7))
    sigma_t = ((1 - torch.cos(math.pi * t.float() / T)) * sigma_max)
    total = 0.0
    for j, (si, ti) in enumerate(((0, 0), (1, 1), (-1, -1))):
        a = _align(s_feats[si], heads[f"align.{j}.weight"], heads[f"align.{j}.bias"])
        tf = t_feats[ti][:, 2:]
        tn = tf / torch.norm(tf, p=2, dim=-1, keepdim=True)
        sn = a / torch.norm(a, p=2, dim=-1, keepdim=True)
        noise = noises[j].to(dt) * sigma_t.view(-1, 1, 1)      # fp32 sigma promotes as in the reference
        pred = denoise_fn(tn + noise, t, heads)
        total = total + F.mse_loss(pred, noise.to(pred.dtype))
        w_t = 1 / (sigma_t ** 2 + 1e-8)
        total = total + w_t.mean() * F.mse_loss(sn, tn)
    return total / 3 * 5e-5


# --------------------------------------------------------------------------- dispatcher
def distillation_loss(dtype_: str, outputs, labels, teacher_logits, s_feats, t_feats, heads, args,
                      alpha: float, tau: float, base_kind: str = "soft_target", noise=None,
                      lrkd_signs=None, probe=None, diff_t=None, diff_noises=None) -> torch.Tensor:
    """DistillationLoss.forward (loss.py:29-242) with the teacher outputs passed in."""
    outputs_kd = None
    if not isinstance(outputs, torch.Tensor):
        outputs, outputs_kd = outputs
    base = base_loss(outputs, labels, base_kind, getattr(args, "smoothing", 0.1))
    t = dtype_.lower()
    if dtype_ == "none":
        return base
    if outputs_kd is None and dtype_ in ("soft", "hard"):
        raise ValueError("soft/hard distillation needs (outputs, outputs_kd)")
    if t == "soft":
        kd = soft_kd(outputs_kd, teacher_logits, tau)
    elif t == "hard":
        kd = hard_kd(outputs_kd, teacher_logits)
    elif t == "vitkd":
        return base + vitkd(s_feats, t_feats, heads, noise, probe=probe)
    elif t == "lrkd":
        kd = lrkd(s_feats, t_feats, heads, args.lrkd_rank,
                  (args.lrkd_alpha, args.lrkd_beta, args.lrkd_gamma), lrkd_signs)
    elif t == "diffkd":
        kd = diffkd(s_feats, t_feats, heads, diff_t, diff_noises)
    elif t == "curkd":
        if args.current_epoch < 151:
            return base + curkd_hidden(s_feats, t_feats, heads, args.current_epoch)
        return base + curkd_late(s_feats, t_feats, heads, noise, probe)
    elif t == "saliency_mgd":
        return base + saliency_mgd(s_feats, t_feats, heads, args.saliency_mask_ratio, args.saliency_method, probe)
    elif t == "wasskd":
        if args.wasskd_type == "l1":
            w = wass_l1(s_feats, t_feats, heads)
        elif args.wasskd_type == "sinkhorn":
            w = wass_sinkhorn(s_feats, t_feats, heads)
        else:
            w = 0.0  # loss.py:186-226: any other value leaves loss_wass = 0.0
        return base + w * 5.0
    elif t == "mgd":
        return base + mgd(s_feats, t_feats, heads, args.mgd_alpha, args.mgd_mask_ratio, noise, probe)
    else:
        raise ValueError(f"Invalid distillation type: {dtype_}")
    return base * (1 - alpha) + kd * alpha


def default_args(**kw) -> SimpleNamespace:
    """Loss-relevant argparse defaults of tools/train.py:103-136,157-186."""
    d = dict(lrkd_rank=32, lrkd_alpha=0.1, lrkd_beta=0.1, lrkd_gamma=0.1, saliency_method=1,
             saliency_mask_ratio=0.5, wasskd_type="l1", mgd_alpha=7e-5, mgd_mask_ratio=0.5,
             mixup=0.8, cutmix=1.0, cutmix_minmax=None, smoothing=0.1, current_epoch=0)
    d.update(kw)
    return SimpleNamespace(**d)
