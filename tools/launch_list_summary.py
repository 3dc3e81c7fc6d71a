#!/usr/bin/env python
"""Summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel and grid size (the full-size launches of the
device-resident steps and the slice launches of the host-buffer steps of bench.py are told apart by their grids).
    python tools/launch_list_summary.py profiles/r02_launches_<sha>.csv [states_per_step]  -> markdown"""
import collections
import csv
import re
import sys

path = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
H = {h: i for i, h in enumerate(rows[0])}
agg = collections.OrderedDict()
for r in rows[1:]:
    name = re.sub(r"^void ", "", r[H["Kernel Name"]]).split("(")[0]
    name = re.sub(r"dwbc::|<unnamed>::", "", name)
    if name.startswith("at::") or "elementwise" in name:
        name = "torch copy / fill kernels"
    grid = int(r[H["Grid Size"]].strip("()").split(",")[0])
    v = float(r[H["Metric Value"]].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r[H["Metric Unit"]], 1e-6)
    key = (name, grid)
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += v
full = {k: v for k, v in agg.items() if k[1] >= B // 4 and "probe" not in k[0] and "dfma" not in k[0]}
tot_full = sum(v[1] for k, v in full.items() if k[0].startswith(("lean_cycle", "hqp", "lean_classify")) or k[1] >= B)
print(f"launch list `{path.split('/')[-1]}` (ncu, cold caches, serialised launches; not a bench number)\n")
print("| kernel | grid | launches | total ms | ms / launch | share of the full-size launches |")
print("|---|---|---|---|---|---|")
for (name, grid), (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    share = f"{100 * ms / tot_full:.1f} %" if (name, grid) in full and tot_full > 0 else ""
    print(f"| `{name[:90]}` | {grid} | {cnt} | {ms:.3f} | {ms / cnt:.3f} | {share} |")
fam = collections.OrderedDict()
for (name, grid), (cnt, ms) in full.items():
    f = "hqp_kernel" if name.startswith("hqp") else "lean_cycle_kernel (+ classify)" if name.startswith("lean") else name
    fam[f] = fam.get(f, 0.0) + ms
print("\nshare of the step by kernel family (launches of at least a quarter of the batch): " + ", ".join(f"{k} {100 * v / tot_full:.1f} %" for k, v in fam.items()))
