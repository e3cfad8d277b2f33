#!/usr/bin/env python
"""Turns `ncu --csv` logs of `bench.py --workload W --ncu-op` into profiles/ncu_traffic.json and a readable launch list.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        --profile-from-start off --csv --log-file gpurun_out/ncuop_<W>.csv python bench.py --workload <W> --ncu-op
    python tools/ncu_traffic.py gpurun_out/ncuop_*.csv [--tag r3a]

Per workload: every kernel ONE call of the fused op launches (name, launches, summed device time, DRAM bytes read +
written), their total = `roofline.traffic` of bench.py.  Times are cold-cache and serialised: compare shares.
"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0,
        "ms": 1e3, "msecond": 1e3}


def short(name: str) -> str:
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("dkd::", "").replace("(anonymous namespace)::", "")
    return name if len(name) <= 110 else name[:107] + "..."


def parse(path: str):
    rows = []
    with open(path, newline="") as fh:
        lines = [l for l in fh if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ix = {h: i for i, h in enumerate(hdr)}
    launches = {}
    for r in rd:
        lid = int(r[ix["ID"]])
        d = launches.setdefault(lid, {"name": short(r[ix["Kernel Name"]]), "us": 0.0, "rd": 0.0, "wr": 0.0,
                                      "grid": r[ix["Grid Size"]], "block": r[ix["Block Size"]]})
        v = float(r[ix["Metric Value"]].replace(",", "")) * UNIT.get(r[ix["Metric Unit"]], 1.0)
        m = r[ix["Metric Name"]]
        if m == "gpu__time_duration.sum":
            d["us"] = v
        elif m == "dram__bytes_read.sum":
            d["rd"] = v
        elif m == "dram__bytes_write.sum":
            d["wr"] = v
    return [launches[k] for k in sorted(launches)]


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    tag = next((sys.argv[i + 1] for i, a in enumerate(sys.argv) if a == "--tag"), "r3")
    out_json = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    table = json.load(open(out_json)) if os.path.exists(out_json) else {}
    text = []
    for path in args:
        w = re.sub(r"^ncuop_", "", os.path.splitext(os.path.basename(path))[0])
        try:
            ls = parse(path)
        except (StopIteration, OSError, KeyError, ValueError) as e:
            print(f"skip {path}: {e}", file=sys.stderr)
            continue
        if not ls:
            continue
        agg = {}
        for l in ls:
            a = agg.setdefault(l["name"], {"launches": 0, "us": 0.0, "dram_bytes": 0.0, "grid": l["grid"], "block": l["block"]})
            a["launches"] += 1
            a["us"] += l["us"]
            a["dram_bytes"] += l["rd"] + l["wr"]
        tot_us = sum(a["us"] for a in agg.values())
        tot_b = sum(a["dram_bytes"] for a in agg.values())
        kern = sorted(({"name": k, **v, "share": v["us"] / tot_us} for k, v in agg.items()), key=lambda x: -x["us"])
        table[w] = {"bytes": tot_b, "us_under_ncu": tot_us, "launches": len(ls), "kernels": kern[:12],
                    "source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum over the {len(ls)} launches of one fused-op call "
                              f"(profiles/{tag}_op_launch_lists.txt)"}
        text.append(f"## {w}: {len(ls)} launches, {tot_us:.1f} us under ncu (cold, serialised), DRAM {tot_b / 1e6:.1f} MB")
        for k in kern:
            text.append(f"  {k['share'] * 100:5.1f} %  {k['us']:9.1f} us  x{k['launches']:<3d} {k['dram_bytes'] / 1e6:9.1f} MB  "
                        f"grid {k['grid']:>12s} block {k['block']:>12s}  {k['name']}")
        text.append("")
    json.dump(table, open(out_json, "w"), indent=1, sort_keys=True)
    with open(os.path.join(ROOT, "profiles", f"{tag}_op_launch_lists.txt"), "w") as fh:
        fh.write("# one call of each workload's fused op (bench.py --workload W --ncu-op) under\n"
                 "# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none\n"
                 "# share of the call | device time | launches | DRAM bytes (read + written) | kernel\n\n" + "\n".join(text))
    print(f"{len(text)} lines -> profiles/{tag}_op_launch_lists.txt; {len(table)} workloads in profiles/ncu_traffic.json")


if __name__ == "__main__":
    main()
