#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu -i ... --page raw --csv) into one line per kernel launch: duration, DRAM bytes,
DRAM / tensor-pipe / SM throughput %, registers, occupancy.   python tools/ncu_summary.py rep [rep ...]"""
import csv, io, subprocess, sys

COLS = {
    "gpu__time_duration.sum": "dur_us",
    "dram__bytes_read.sum": "dram_rd_MB",
    "dram__bytes_write.sum": "dram_wr_MB",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_inst",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occ_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "lts__t_bytes.sum": "l2_MB",
    "smsp__cycles_active.avg": "cycles",
}

def to_float(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None

for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {rep}")
    for r in rows[2:]:
        name = r[idx["Kernel Name"]][:60]
        parts = [f"{name:60s}"]
        for m, short in COLS.items():
            if m not in idx:
                continue
            v, u = to_float(r[idx[m]]), units[idx[m]]
            if v is None:
                continue
            if short == "dur_us":
                v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
            if short.endswith("_MB"):
                scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
                v *= scale
            parts.append(f"{short}={v:.4g}")
        print("  ".join(parts))
