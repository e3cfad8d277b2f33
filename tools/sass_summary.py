#!/usr/bin/env python
"""Per-kernel SASS op-count summary of libdeltakd_sm100.so (what proves a Blackwell-native kernel, B200_PROFILING.md):
UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor loads / stores, UBLKCP = cp.async.bulk (1-D bulk copies),
LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, HMMA = legacy mma.sync (must be 0).

    python tools/sass_summary.py [lib.so] > profiles/<tag>_sass_ops.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "deltakd_b200", "libdeltakd_sm100.so")
OPS = ["UTCHMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "SYNCS", "UCGABAR", "HMMA", "LDGSTS", "MUFU", "FFMA2", "DFMA"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
counts = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for o in OPS:            # prefix match: UCGABAR_ARV / UCGABAR_WAIT -> UCGABAR, MUFU.EX2 -> MUFU
            if op.startswith(o):
                counts[cur][o] += 1
                break
dem = subprocess.run(["cu++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
names = dict(zip(counts, dem)) if len(dem) == len(counts) else {k: k for k in counts}


def short(n):
    n = re.sub(r"\(anonymous namespace\)::", "", n).replace("dkd::", "")
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", n)
    return n if len(n) <= 150 else n[:147] + "..."


tot = collections.Counter()
print(f"# {os.path.relpath(lib, ROOT)}: {len(counts)} kernels; columns: " + " ".join(OPS) + " | total instructions | kernel")
for k, c in sorted(counts.items(), key=lambda kv: -kv[1]["UTCHMMA"] * 100000 - kv[1]["_total"]):
    for o in OPS:
        tot[o] += c[o]
    print(" ".join(f"{c[o]:5d}" for o in OPS) + f" | {c['_total']:6d} | {short(names[k])}")
print("# totals: " + ", ".join(f"{o} {tot[o]}" for o in OPS))
