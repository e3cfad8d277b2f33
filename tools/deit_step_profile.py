"""Kernel-level profile (torch.profiler) of the end-to-end DeiT-Tiny KD step of bench.py.  Dev tool.
    python tools/deit_step_profile.py [soft|curkd]"""
import sys, torch
sys.path.insert(0, '.')
import bench
from torch.profiler import profile, ProfilerActivity

kind = sys.argv[1] if len(sys.argv) > 1 else "soft"
cls = bench.WORKLOADS[f"deit_tiny_kd_step_{kind}_b256_bf16"]
w = cls(torch.device("cuda"), 0)
w.setup()
host = w.host_sets(2)
ds = [w.to_device(h) for h in host]
for i in range(5):
    w.step(ds[i % 2])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(10):
    w.step(ds[i % 2])
e1.record(); torch.cuda.synchronize()
print("ms/step", e0.elapsed_time(e1) / 10)
import time
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(10):
    w.step(ds[i % 2])
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("CPU enqueue ms/step", (t1 - t0) * 100, " wall ms/step", (t2 - t0) * 100)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as p:
    w.step(ds[0]); torch.cuda.synchronize()
evs = [e for e in p.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"] or p.key_averages()
tot = sum(e.self_device_time_total for e in p.key_averages())
from torch.autograd import DeviceType
kern = [e for e in p.events() if e.device_type == DeviceType.CUDA]
print("GPU kernel time per step (us):", sum(e.device_time for e in kern), " kernels:", len(kern))
for r in sorted(p.key_averages(), key=lambda r: -r.self_device_time_total)[:40]:
    if r.self_device_time_total <= 0: break
    print(f"{r.self_device_time_total:10.0f} us  x{r.count:<4d} {r.key[:130]}")
