#!/usr/bin/env python
"""Compact table of a bench.py JSON line:  python tools/bench_table.py gpurun_out/x_bench.json"""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads([l for l in open(path) if l.startswith("{")][-1])
    except (IndexError, ValueError, OSError) as e:
        print(f"{path}: no JSON line ({e})")
        continue
    rows = [dict(workload=d["config"]["workload"], ms_per_step=d["ms_per_step"], value=d["value"], e2e=d["e2e"], roofline=d["roofline"],
                 parity=d["config"].get("parity"), cpu_baseline=d.get("cpu_baseline"))] + d.get("workloads", [])
    print(f"# {path}: n_gpus={d['n_gpus']} clocks={d.get('clocks')}")
    for r in rows:
        if "error" in r:
            print(f"{r['workload']:40s} ERROR {r['error'][:200]}")
            continue
        rf = r["roofline"]
        par = r.get("parity") or {}
        cpu = r.get("cpu_baseline") or {}
        extra = f" mufu={rf['mufu_frac']:.3f}" if "mufu_frac" in rf else ""
        if "eigensolve_us" in rf:
            extra += f" eig={rf['eigensolve_us']:.0f}us/{max(rf['eigensolve_sweeps'])}sw"
        print(f"{r['workload']:40s} {r['ms_per_step'] * 1e3:10.1f} us/step  op {rf['kernel_us']:9.1f} us  {rf['bound']:6s} frac {rf['frac']:.3f}{extra}"
              f"  e2e {r["e2e"]["ms_per_step"] * 1e3:9.1f} us ({r["e2e"].get("mode", "serial")}; serial {r["e2e"].get("serial_ms_per_step", float("nan")) * 1e3:.1f})  value {r['value']:.4g}  cpu {cpu.get('value', float('nan')):.4g}"
              f"  parity {par.get('rel_err', float('nan')):.1e}/{par.get('batch')}")
