"""Dev tool: CUPTI kernel breakdown of the end-to-end DeiT KD step (bench.py DeiTKDStep).  python tools/deit_probe.py"""
import sys, torch
sys.path.insert(0, '.')
import bench
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda", 0)
w = bench.DeiTKDStep(dev, 0)
w.setup()
ds = w.to_device(w.host_sets(1)[0])
for _ in range(3):
    w.step(ds)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    w.step(ds)
e1.record(); torch.cuda.synchronize()
print("ms/step", e0.elapsed_time(e1) / 5)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as p:
    w.step(ds); torch.cuda.synchronize()
tot = sum(r.device_time_total for r in p.key_averages() if r.device_time_total)
rows = sorted(p.key_averages(), key=lambda r: -r.self_device_time_total)[:22]
for r in rows:
    print(f"{r.self_device_time_total/1e3:9.2f} ms x{r.count:<4d} {r.key[:100]}")
