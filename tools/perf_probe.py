"""Per-kernel timing of the feature losses at full size (B=512) via torch.profiler (CUPTI). Dev tool."""
import sys, torch
sys.path.insert(0, '.')
from types import SimpleNamespace
from torch.profiler import profile, ProfilerActivity
from deltakd_b200 import functional as Fn, synth, heads as H

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
which = sys.argv[2] if len(sys.argv) > 2 else "all"
dev = torch.device("cuda")


def feats(layers, dtype=torch.float32):
    g = torch.Generator(device="cuda").manual_seed(0)
    s = {i: torch.randn(B, 197, 192, device=dev, generator=g).to(dtype) for i in layers}
    t = {i: torch.randn(B, 198, 384, device=dev, generator=g).to(dtype) for i in layers}
    return s, t


def prof(name, fn, iters=3):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    print(f"== {name}: {e0.elapsed_time(e1) / iters * 1e3:.1f} us / call (B={B})")
    with profile(activities=[ProfilerActivity.CUDA]) as p:
        fn(); torch.cuda.synchronize()
    rows = sorted(p.key_averages(), key=lambda r: -r.device_time_total)[:14]
    for r in rows:
        print(f"   {r.device_time_total / r.count:10.1f} us x{r.count:<3d} {r.key[:110]}")


for mode, dtype in (("bf16x3", torch.float32), ("bf16", torch.float32), ("bf16", torch.bfloat16)):
    Fn.set_matmul_precision(mode)
    tag = f"{mode}/{str(dtype).split('.')[-1]}"
    if which in ("all", "curkd"):
        s, t = feats([0])
        s, t = {0: s[0].to(dtype).requires_grad_(True)}, {0: t[0].to(dtype)}
        lin = torch.nn.Linear(192, 384).cuda()
        def f():
            l = Fn.align_mse_layers_loss([s[0]], [t[0]], [lin], 1e-6); l.backward()
        prof(f"align_mse 1 layer {tag}", f)
    if which in ("all", "mgd"):
        args = SimpleNamespace(distillation_type="mgd")
        teacher, student = synth.FeatureReplayModel(384), synth.FeatureReplayModel(192)
        H.attach_distillation_heads(student, teacher, args)
        student = student.cuda()
        s, t = feats([11])
        s11 = s[11].to(dtype).requires_grad_(True); t11 = t[11].to(dtype)
        def f():
            l = Fn.masked_generation_loss(s11, t11, student.align, student.mask_token, student.generation,
                                          mask_ratio=0.5, scale=1e-9); l.backward()
        prof(f"mgd {tag}", f, iters=2)
