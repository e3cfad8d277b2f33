"""Full-size launches of the token-stream row ops (shapes of the DeiT-Tiny / DeiT-Small KD step at B = 256), for
`ncu -k regex:layernorm|colsum_kernel|head_copy|fold_columns --launch-skip 9 -c 9` and for CUDA-event timing.  Dev tool.
    python tools/rowops_probe.py"""
import sys, torch
sys.path.insert(0, '.')
from deltakd_b200 import functional as Fn

dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)
M = 256 * 197
xs = torch.randn(M, 192, device=dev, generator=g).requires_grad_(True)              # student residual stream (fp32)
ws, bs = (torch.randn(192, device=dev, generator=g).requires_grad_(True) for _ in range(2))
dys = torch.randn(M, 192, device=dev, generator=g).bfloat16()
xt = torch.randn(256 * 198, 384, device=dev, generator=g).bfloat16()                 # teacher stream (bf16, inference)
wt, bt = (torch.randn(384, device=dev, generator=g).bfloat16() for _ in range(2))
d768 = torch.randn(M, 768, device=dev, generator=g).bfloat16()                       # fc1 bias gradient
att = torch.randn(256, 3, 197, 64, device=dev, generator=g).bfloat16()               # attention output [B,H,N,hd]


def one_pass():
    xs.grad = ws.grad = bs.grad = None
    y = Fn.layer_norm(xs, ws, bs, 1e-6, torch.bfloat16)      # layernorm_fwd <float,float,bf16>
    y.backward(dys)                                           # layernorm_bwd + fold_columns
    with torch.no_grad():
        Fn.layer_norm(xt, wt, bt, 1e-6)                       # layernorm_fwd <bf16,bf16,bf16>
    Fn.column_sum(d768)                                       # colsum + fold_columns
    Fn.column_sum(dys)                                        # colsum + fold_columns
    Fn.merge_heads(att)                                       # head_copy


one_pass()
torch.cuda.synchronize()
one_pass()
torch.cuda.synchronize()
# CUDA-event timing of each op (20 reps, inputs > L2 in total are rotated by the other ops in the pass)
def t(fn, n=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
GB = 1e9
print("ln fwd student  us %.1f  GB/s %.0f" % ((u := t(lambda: Fn.layer_norm(xs.detach(), ws.detach(), bs.detach(), 1e-6, torch.bfloat16))), M * 192 * 6 / u / 1e3))
def fb():
    xs.grad = ws.grad = bs.grad = None
    Fn.layer_norm(xs, ws, bs, 1e-6, torch.bfloat16).backward(dys)
u2 = t(fb)
print("ln fwd+bwd student us %.1f  (bwd ~%.1f us, %.0f GB/s)" % (u2, u2 - u, M * 192 * 10 / (u2 - u) / 1e3))
with torch.no_grad():
    print("ln fwd teacher  us %.1f  GB/s %.0f" % ((v := t(lambda: Fn.layer_norm(xt, wt, bt, 1e-6))), xt.numel() * 4 / v / 1e3))
print("colsum 768      us %.1f  GB/s %.0f" % ((c := t(lambda: Fn.column_sum(d768))), d768.numel() * 2 / c / 1e3))
print("colsum 192      us %.1f  GB/s %.0f" % ((c := t(lambda: Fn.column_sum(dys))), dys.numel() * 2 / c / 1e3))
print("merge heads     us %.1f  GB/s %.0f" % ((h := t(lambda: Fn.merge_heads(att))), att.numel() * 4 / h / 1e3))
