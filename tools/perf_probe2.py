"""Per-kernel timing (torch.profiler / CUPTI) of one fused loss op at full size. Dev tool.
    python tools/perf_probe2.py lrkd|wass|sinkhorn|curkd|salmgd [B]"""
import sys, torch
sys.path.insert(0, '.')
from types import SimpleNamespace
from torch.profiler import profile, ProfilerActivity
from deltakd_b200 import functional as Fn, synth, heads as H

which = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)
mk = lambda *sh: torch.randn(*sh, device=dev, generator=g)


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
    for r in sorted(p.key_averages(), key=lambda r: -r.device_time_total)[:16]:
        print(f"   {r.device_time_total / r.count:10.1f} us x{r.count:<3d} {r.key[:120]}")


if which == "lrkd":
    s = [mk(B, 197, 192).requires_grad_(True) for _ in range(3)]
    t = [mk(B, 198, 384) for _ in range(3)]
    lins = [torch.nn.Linear(192, 64).cuda() for _ in range(3)]
    basis = {}
    def f():
        l = Fn.lrkd_layers_loss(s, t, lins, 64, (0.2, 0.2, 0.2), basis_out=basis); l.backward()
    prof("lrkd rank 64, 3 layers", f)
    print("sweeps", basis["sweeps"].tolist())
elif which == "wass":
    s = [mk(B, 197, 192).requires_grad_(True) for _ in range(3)]
    t = [mk(B, 198, 384) for _ in range(3)]
    lins = [torch.nn.Linear(192, 384).cuda() for _ in range(3)]
    def f():
        l = Fn.wass_l1_loss(s, t, lins); l.backward()
    prof("wass l1, 3 layers", f)
elif which == "sinkhorn":
    s = [(0.5 * mk(B, 197, 192)).requires_grad_(True) for _ in range(3)]
    t = [0.5 * mk(B, 198, 384) + 0.1 for _ in range(3)]
    lins = [torch.nn.Linear(192, 384).cuda() for _ in range(3)]
    def f():
        l = Fn.wass_sinkhorn_loss(s, t, lins); l.backward()
    prof("wass sinkhorn, 3 layers", f, iters=2)
elif which == "curkd":
    s = [mk(B, 197, 192).requires_grad_(True) for _ in range(3)]
    t = [mk(B, 198, 384) for _ in range(3)]
    lins = [torch.nn.Linear(192, 384).cuda() for _ in range(3)]
    def f():
        l = Fn.align_mse_layers_loss(s, t, lins, 1e-6); l.backward()
    prof("curkd early, 3 layers", f)
elif which == "curkd_bf16":
    s = [mk(B, 197, 192).bfloat16().requires_grad_(True) for _ in range(3)]
    t = [mk(B, 198, 384).bfloat16() for _ in range(3)]
    lins = [torch.nn.Linear(192, 384).cuda() for _ in range(3)]
    def f():
        l = Fn.align_mse_layers_loss(s, t, lins, 1e-6); l.backward()
    prof("curkd early bf16, 3 layers", f)
elif which == "salmgd":
    from deltakd_b200.misc import saliency_scores
    for m in (1, 2, 3):
        args = SimpleNamespace(distillation_type="saliency_mgd", saliency_method=m)
        teacher, student = synth.FeatureReplayModel(384), synth.FeatureReplayModel(192)
        H.attach_distillation_heads(student, teacher, args)
        student = student.cuda()
        t = mk(B, 198, 384)
        prof(f"saliency score method {m}", lambda: saliency_scores(student, t, m))
