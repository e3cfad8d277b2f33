"""GPU: token-stream row ops (dkd_layernorm_fwd / _bwd, dkd_colsum) against fp64 torch references of the same op,
and the bench-harness DeiT built on them against the same model on ATen's LayerNorm / Linear."""
import pytest
import torch
import torch.nn.functional as F

from oracle.util import rel_err

pytestmark = pytest.mark.gpu

LN_RTOL = 2e-6          # fp32 in / fp32 out, vs fp64
BF16_RTOL = 4e-3        # outputs rounded to bf16 on store (2^-9 relative per element)


def _ref_ln(x, w, b, eps, dy):
    x64 = x.detach().double().cpu().requires_grad_(True)
    w64 = w.detach().double().cpu().requires_grad_(True)
    b64 = b.detach().double().cpu().requires_grad_(True)
    y = F.layer_norm(x64, (x64.shape[-1],), w64, b64, eps)
    y.backward(dy.detach().double().cpu())
    return y.detach(), x64.grad, w64.grad, b64.grad


@pytest.mark.parametrize("M,D", [(1, 192), (7, 8), (777, 192), (50432, 192), (1001, 384), (333, 512), (65, 132)])
@pytest.mark.parametrize("xdt,pdt,ydt", [(torch.float32, torch.float32, torch.float32),
                                         (torch.float32, torch.float32, torch.bfloat16),
                                         (torch.bfloat16, torch.bfloat16, torch.bfloat16)])
def test_layernorm_matches_fp64(M, D, xdt, pdt, ydt):
    from deltakd_b200 import functional as Fn
    g = torch.Generator(device="cuda").manual_seed(M * 31 + D)
    x = (torch.randn(M, D, device="cuda", generator=g) * 1.7 + 0.3).to(xdt).requires_grad_(True)
    w = (1.0 + 0.2 * torch.randn(D, device="cuda", generator=g)).to(pdt).requires_grad_(True)
    b = (0.1 * torch.randn(D, device="cuda", generator=g)).to(pdt).requires_grad_(True)
    dy = torch.randn(M, D, device="cuda", generator=g).to(ydt)
    y = Fn.layer_norm(x, w, b, 1e-6, ydt)
    assert y.dtype == ydt and y.shape == x.shape
    y.backward(dy)
    ry, rdx, rdw, rdb = _ref_ln(x, w, b, 1e-6, dy)
    tol_y = LN_RTOL if ydt == torch.float32 else BF16_RTOL
    tol_dx = LN_RTOL if xdt == torch.float32 else BF16_RTOL
    tol_p = 5e-6 if pdt == torch.float32 else BF16_RTOL
    assert rel_err(y, ry) < tol_y
    assert x.grad.dtype == xdt and rel_err(x.grad, rdx) < tol_dx
    assert w.grad.dtype == pdt and rel_err(w.grad, rdw) < tol_p
    assert b.grad.dtype == pdt and rel_err(b.grad, rdb) < tol_p


def test_layernorm_inference_and_3d():
    from deltakd_b200 import functional as Fn
    x = torch.randn(4, 198, 384, device="cuda").bfloat16()
    w = torch.ones(384, device="cuda").bfloat16()
    b = torch.zeros(384, device="cuda").bfloat16()
    with torch.no_grad():
        y = Fn.layer_norm(x, w, b, 1e-6)
    ref = F.layer_norm(x.double(), (384,), w.double(), b.double(), 1e-6)
    assert y.shape == x.shape and rel_err(y, ref) < BF16_RTOL
    with pytest.raises(RuntimeError):
        Fn.layer_norm(torch.randn(4, 6), torch.ones(6), torch.zeros(6))          # CPU tensors: no CPU path
    with pytest.raises(RuntimeError):
        Fn.layer_norm(torch.randn(4, 6, device="cuda"), torch.ones(6, device="cuda"), None)   # D % 4


@pytest.mark.parametrize("M,N", [(1, 8), (5, 192), (1024, 576), (50432, 192), (50432, 768), (4097, 1000), (300, 2048)])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_colsum_matches_fp64_and_is_deterministic(M, N, dt):
    from deltakd_b200 import functional as Fn
    g = torch.Generator(device="cuda").manual_seed(M + N)
    a = torch.randn(M, N, device="cuda", generator=g).to(dt)
    out = Fn.column_sum(a)
    ref = a.double().sum(0)
    scale = a.double().abs().sum(0).max().item()          # sums of zero-mean data: compare against the magnitude summed
    assert out.dtype == torch.float32 and out.shape == (N,)
    assert (out.double() - ref).abs().max().item() <= 2e-6 * scale
    assert torch.equal(out, Fn.column_sum(a))


def test_deit_harness_fused_row_ops_match_aten():
    """Same weights, same input: the harness model on dkd LayerNorm / colsum vs plain nn.LayerNorm / nn.Linear."""
    import torch.nn as nn
    from deltakd_b200 import deit
    torch.manual_seed(0)
    m = deit.create_model("deit_tiny_distilled_patch16_224", num_classes=100, row_ops="dkd").cuda().train()
    m.set_distilled_training(True)
    ref = deit.create_model("deit_tiny_distilled_patch16_224", num_classes=100, row_ops="aten").cuda().train()
    ref.set_distilled_training(True)
    ref.load_state_dict(m.state_dict())            # same parameter names and shapes
    assert isinstance(m.blocks[0].norm1, deit.LayerNorm) and type(ref.blocks[0].norm1) is nn.LayerNorm
    with pytest.raises(RuntimeError):
        m.blocks[0].norm1.cpu()(torch.randn(2, 197, 192))   # dkd modules have no CPU path
    m.cuda()
    x = torch.randn(8, 3, 224, 224, device="cuda")
    outs = []
    for model in (m, ref):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            a, b = model(x)
        (a.float().square().mean() + b.float().square().mean()).backward()
        outs.append((a, b))
    assert rel_err(outs[0][0], outs[1][0]) < 3e-2 and rel_err(outs[0][1], outs[1][1]) < 3e-2   # bf16 model, 12 blocks
    for (n1, p1), (n2, p2) in zip(m.named_parameters(), ref.named_parameters()):
        assert n1 == n2 and p1.grad is not None and p2.grad is not None, n1
        assert p1.grad.dtype == p2.grad.dtype
    for name in ("blocks.11.norm2.weight", "blocks.0.norm1.bias", "blocks.5.mlp.fc1.bias", "blocks.0.attn.qkv.bias", "norm.weight",
                 "blocks.3.mlp.fc2.weight", "pos_embed"):
        g1, g2 = dict(m.named_parameters())[name].grad, dict(ref.named_parameters())[name].grad
        assert rel_err(g1, g2) < 8e-2, (name, rel_err(g1, g2))


@pytest.mark.parametrize("B,H,N,hd,dt", [(2, 3, 197, 64, torch.bfloat16), (3, 6, 198, 64, torch.bfloat16), (1, 1, 5, 8, torch.float32),
                                         (4, 3, 197, 64, torch.float32)])
def test_attention_head_relayouts_are_exact(B, H, N, hd, dt):
    """split_qkv / merge_heads (dkd_head_copy) are pure relayouts: forward and backward bit-identical to ATen's."""
    from deltakd_b200 import functional as Fn
    g = torch.Generator(device="cuda").manual_seed(B * 100 + N)
    qkv = torch.randn(B, N, 3 * H * hd, device="cuda", generator=g).to(dt)
    a = qkv.clone().requires_grad_(True)
    b = qkv.clone().requires_grad_(True)
    q1, k1, v1 = Fn.split_qkv(a, H)
    q2, k2, v2 = b.reshape(B, N, 3, H, hd).permute(2, 0, 3, 1, 4).unbind(0)
    for u, w in ((q1, q2), (k1, k2), (v1, v2)):
        assert u.shape == (B, H, N, hd) and torch.equal(u, w)
    # a [B,H,N,hd]-contiguous "attention output" built from the three parts, merged back to token-major
    o1 = (q1 * 2 + k1).contiguous() + v1
    o2 = (q2 * 2 + k2).contiguous() + v2
    y1 = Fn.merge_heads(o1)
    y2 = o2.transpose(1, 2).reshape(B, N, H * hd)
    assert y1.shape == (B, N, H * hd) and y1.is_contiguous() and torch.equal(y1, y2)
    dy = torch.randn(B, N, H * hd, device="cuda", generator=g).to(dt)
    y1.backward(dy)
    y2.backward(dy)
    assert torch.equal(a.grad, b.grad)
    # merge of an already token-major tensor is a view
    z = torch.randn(B, N, H, hd, device="cuda").to(dt).transpose(1, 2)
    assert Fn.merge_heads(z).data_ptr() == z.data_ptr()
