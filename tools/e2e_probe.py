"""Where the end-to-end time of the headline step (configs[1], host buffers) goes.  Dev tool.
    python tools/e2e_probe.py"""
import sys, time, torch
sys.path.insert(0, '.')
import bench

dev = torch.device("cuda")
w = bench.LogitKD(dev, 0)
w.setup()
host = w.host_sets(8)
for i in range(20):
    w.step(w.to_device(host[i % 8])).item()
torch.cuda.synchronize()

def timed(fn, n=200):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n):
        fn(i)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6

print("full step (h2d + loss + backward + item): %.1f us" % timed(lambda i: w.step(w.to_device(host[i % 8])).item()))
print("h2d only + sync                         : %.1f us" % timed(lambda i: (host[i % 8].to(dev, non_blocking=True), torch.cuda.synchronize())))
ds = w.to_device(host[0])
print("step on resident inputs + item          : %.1f us" % timed(lambda i: w.step(ds).item()))
def fwd_only(i):
    z, zk, zt, y = ds
    w.teacher.set_outputs(zt, None)
    return w.crit(w.inputs, (z, zk), None, None, y, w.args)
print("forward only (no sync per step)         : %.1f us" % timed(fwd_only))
print("forward + item                          : %.1f us" % timed(lambda i: fwd_only(i).item()))
def fb(i):
    l = fwd_only(i); l.backward(); return l
print("forward + backward (no sync per step)   : %.1f us" % timed(fb))
print("to_device only (no sync per step)       : %.1f us" % timed(lambda i: w.to_device(host[i % 8])))

if len(sys.argv) > 1 and sys.argv[1] == "cprofile":
    import cProfile, pstats
    pr = cProfile.Profile()
    pr.enable()
    for i in range(2000):
        w.step(w.to_device(host[i % 8])).item()
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(35)
