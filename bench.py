#!/usr/bin/env python
"""bench.py — DeltaKD distillation-loss hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

A "step" is one pass of the hot path (loss forward + backward through
`deltakd_b200.DistillationLoss`) over one batch of synthetic teacher/student outputs.
Prints ONE JSON line (rank 0).  Keys follow the driver contract; see DESIGN.md §Measurement.

  value        samples/s, whole job, inputs resident in HBM, the step's launches replayed from a
               CUDA graph (no Python between launches); input sets rotate through > L2-size data
  e2e          same metric through the public API with HOST (pinned) inputs: per step H2D of the
               step's tensors, criterion(...) + backward, D2H read of the loss
  roofline     dominant kernel: algorithmic bytes (or flops) per launch / its mean duration,
               measured with CUDA events around that kernel's launches in a separate timed loop
  cpu_baseline the CPU oracle (restatement of the reference's PyTorch loss, `kind: "port"`) timed on
               this box's host cores on a bounded sample of the same workload
  --impl reference : the reference arm = the same CPU oracle with all host threads (the reference is
               pure PyTorch; /root/reference does not exist on the GPU box)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

L2_BYTES = 126 * 1024 * 1024


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sus=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sus=1400.0, src="fallback")


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [s for t, s in self.samples if t0 - 0.05 <= t <= t1 + 0.15] or [s for _, s in self.samples[-3:]]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _pin(t):
    return t.pin_memory() if torch.cuda.is_available() else t


# ----------------------------------------------------------------------------- workloads
class Workload:
    """One BASELINE.json config: builds device inputs, the step, its algorithmic cost and the CPU oracle."""
    name = ""
    dtype = "f32"

    def __init__(self, device, rank):
        self.device, self.rank = device, rank

    # number of rotating input sets so the working set exceeds L2
    def nsets(self):
        return max(2, int(1.5 * L2_BYTES / max(1, self.bytes_per_set())) + 1)


class LogitKD(Workload):
    """configs[1]: base SoftTargetCE + soft-KD (tau 3, alpha 0.1) on DeiT-Tiny logits, B=256, C=1000, bf16."""
    name = "soft_kd_logits_b256_c1000_bf16"
    dtype = "bf16"
    B, C, kind, alpha, tau = 256, 1000, "soft", 0.1, 3.0
    tdtype = torch.bfloat16
    launches_per_step = 3   # fused fwd+bwd kernel + 2 conditional-rescale kernels in backward
    dominant = "logit_kd_kernel"

    def bytes_per_set(self):
        return 4 * self.B * self.C * self.tdtype.itemsize

    def algorithmic_bytes(self):  # 4 reads + 2 gradient writes (6*B*C*elt)
        return 6 * self.B * self.C * self.tdtype.itemsize

    def host_sets(self, n):
        from deltakd_b200 import synth
        sets = []
        for i in range(n):
            z, zk, zt, y = synth.make_logits(self.B, self.C, 1234 + self.rank + 17 * i)
            sets.append(tuple(_pin(t.to(self.tdtype)) for t in (z, zk, zt, y)))
        return sets

    def setup(self):
        from deltakd_b200 import DistillationLoss, call_base_loss, synth
        from deltakd_b200.synth import default_args
        self.args = default_args()
        self.teacher = synth.FeatureReplayModel(384)
        self.crit = DistillationLoss(call_base_loss(self.args), self.teacher, self.kind, self.alpha, self.tau)
        self.inputs = torch.zeros(self.B, 3, 2, 2, device=self.device)  # images feed only the (replayed) teacher

    def to_device(self, hs):
        z, zk, zt, y = (t.to(self.device, non_blocking=True) for t in hs)
        return z.requires_grad_(True), zk.requires_grad_(True), zt, y

    def h2d_bytes(self):
        return self.bytes_per_set()

    def step(self, ds):
        z, zk, zt, y = ds
        self.teacher.set_outputs(zt, None)
        z.grad = zk.grad = None
        loss = self.crit(self.inputs, (z, zk), None, None, y, self.args)
        loss.backward()
        return loss

    def kernel_only(self, ds):
        """Launch just the dominant kernel through the C ABI (for the roofline timing)."""
        from deltakd_b200 import functional as Fn
        z, zk, zt, y = ds
        with torch.no_grad():
            pass
        return Fn.logit_kd_loss(z, zk, zt, y, kd_kind=self.kind, alpha=self.alpha, tau=self.tau)

    def cpu_step(self, hs):
        from oracle import losses as O
        if not hasattr(self, "_cpu32"):
            self._cpu32 = {}
        key = id(hs)
        if key not in self._cpu32:  # the reference runs fp32 end to end: convert once, outside the timing
            self._cpu32[key] = tuple(t.float() for t in hs)
        z, zk, zt, y = self._cpu32[key]
        z = z.detach().requires_grad_(True); zk = zk.detach().requires_grad_(True)
        l = O.distillation_loss(self.kind, (z, zk), y, zt, None, None, {}, self.args, self.alpha, self.tau)
        l.backward()
        return l


WORKLOADS = {w.name: w for w in (LogitKD,)}
DEFAULT = LogitKD.name


def cpu_baseline(w: Workload, budget_s: float = 12.0, max_steps: int = 20000):
    torch.set_num_threads(os.cpu_count() or 1)
    hs = w.host_sets(2)
    w.cpu_step(hs[0])
    t0 = time.perf_counter(); n = 0
    while True:
        w.cpu_step(hs[n % 2]); n += 1
        dt = time.perf_counter() - t0
        if dt > budget_s or n >= max_steps:
            break
    return {"value": w.B * n / dt, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n} steps of {w.name} (fp32 torch-CPU oracle of the reference loss, fwd+bwd), {dt:.1f} s"}, dt / n


def run_reference(args, w_cls):
    """Reference arm: the reference's CPU implementation of the path (oracle port), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = w_cls(torch.device("cpu"), 0)
    from deltakd_b200.synth import default_args
    w.args = default_args()
    torch.set_num_threads(os.cpu_count() or 1)
    hs = w.host_sets(2)
    for i in range(max(args.warmup, 1)):
        w.cpu_step(hs[i % 2])
    # each step = one batch; bounded so K steps end within minutes
    t0 = time.perf_counter()
    for i in range(args.steps):
        w.cpu_step(hs[i % 2])
    dt = time.perf_counter() - t0
    v = w.B * args.steps / dt
    print(json.dumps({
        "impl": "reference", "metric": "distill-loss fwd+bwd samples/s", "value": v, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w.name, "note": "reference loss path restated for CPU (oracle port), torch CPU fp32"},
        "cpu_baseline": {"value": v, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{args.steps} steps of {w.name}"},
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default=DEFAULT, choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    w_cls = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, w_cls)

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist_on = world > 1
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if dist_on:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    W = max(args.warmup, 3)
    K = args.steps
    pk = peaks()

    w = w_cls(dev, rank)
    w.setup()
    nsets = w.nsets()
    host = w.host_sets(nsets)
    dsets = [w.to_device(h) for h in host]
    torch.cuda.synchronize()

    sampler = ClockSampler(local).start() if rank == 0 else None

    # ---- device-resident throughput: warm up eagerly, then capture K steps into a CUDA graph --------
    for i in range(W):
        w.step(dsets[i % nsets])
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        w.step(dsets[0])
    torch.cuda.current_stream().wait_stream(side)
    with torch.cuda.graph(graph):
        for i in range(K):
            last = w.step(dsets[(W + i) % nsets])
    graph.replay()  # one untimed replay (graph upload)
    torch.cuda.synchronize()
    if dist_on:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    if dist_on:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    loss_val = float(last.item())

    # ---- dominant kernel alone: events around each launch ------------------------------------------
    evs = []
    for i in range(min(K, 200)):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ds = dsets[(i * 7) % nsets]
        a.record(); w.kernel_only(ds); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    kms = sorted(a.elapsed_time(b) for a, b in evs)
    k_ms = sum(kms) / len(kms)

    # ---- end to end through the public API with host buffers ---------------------------------------
    for i in range(W):
        l = w.step(w.to_device(host[i % nsets])); l.item()
    torch.cuda.synchronize()
    if dist_on:
        dist.barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for i in range(K):
        l = w.step(w.to_device(host[(W + i) % nsets]))
        l.item()
    e3.record()
    torch.cuda.synchronize()
    t_wall1 = time.time()
    ms_e2e = e2.elapsed_time(e3)

    if dist_on:
        t = torch.tensor([ms, ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
    if rank != 0:
        if dist_on:
            dist.destroy_process_group()
        return

    clocks = sampler.stop(t_wall0, t_wall1)
    n = world
    value = n * w.B * K / (ms * 1e-3)
    e2e_v = n * w.B * K / (ms_e2e * 1e-3)
    ach = w.algorithmic_bytes() / (k_ms * 1e-3) / 1e9
    out = {
        "metric": "distill-loss fwd+bwd samples/s", "value": value, "unit": "samples/s", "n_gpus": n, "steps": K,
        "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": w.dtype, "data": "synthetic",
        "config": {"workload": w.name, "batch_per_gpu": w.B, "timing": "CUDA-graph replay of K steps, CUDA events",
                   "l2": f"{nsets} rotating input sets ({nsets * w.bytes_per_set() / 2**20:.0f} MiB > 126 MiB L2)",
                   "teacher": "teacher outputs replayed (inputs of the loss path)", "loss": loss_val},
        "e2e": {"value": e2e_v, "unit": "samples/s", "h2d_bytes_per_step": w.h2d_bytes(), "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / K},
        "gpu_launches": K * w.launches_per_step,
        "roofline": {"bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"],
                     "traffic": None, "kernel": w.dominant, "kernel_us": k_ms * 1e3, "peak_source": pk["src"],
                     "algorithmic_bytes": w.algorithmic_bytes()},
        "clocks": clocks,
    }
    if not args.no_cpu_baseline:
        out["cpu_baseline"], _ = cpu_baseline(w)
    print(json.dumps(out))
    if dist_on:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
