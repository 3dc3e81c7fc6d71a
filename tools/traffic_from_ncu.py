#!/usr/bin/env python
"""DRAM bytes per launch of the kernels of one device-resident bench step from an `ncu --set full` capture (raw page as csv):
    ncu -i gpurun_out/r02_full.ncu-rep --page raw --csv > raw.csv;  python tools/traffic_from_ncu.py raw.csv 65536 "<source note>" > profiles/r02_traffic.json
Launches of the same kernel family (lean_classify + the per-class lean_cycle launches) are summed under bench.py's kernel name."""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
B = int(sys.argv[2])
hdr, units = rows[0], rows[1]
H = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {}
for r in rows[2:]:
    name = r[H["Kernel Name"]]
    key = "hqp_kernel" if "hqp_kernel" in name else "lean_cycle_kernel" if "lean_" in name else name.split("(")[0]
    rd = float(r[H["dram__bytes_read.sum"]].replace(",", "")) * scale[units[H["dram__bytes_read.sum"]]]
    wr = float(r[H["dram__bytes_write.sum"]].replace(",", "")) * scale[units[H["dram__bytes_write.sum"]]]
    ns = float(r[H["gpu__time_duration.sum"]].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6}[units[H["gpu__time_duration.sum"]]]
    k = out.setdefault(key, {"dram_read_bytes": 0.0, "dram_write_bytes": 0.0, "launches": 0, "duration_ms_under_ncu": 0.0, "grids": []})
    k["dram_read_bytes"] += rd
    k["dram_write_bytes"] += wr
    k["launches"] += 1
    k["duration_ms_under_ncu"] += ns * 1e-6
    k["grids"].append(r[H["Grid Size"]])
print(json.dumps({"states_per_launch": B, "config": "2", "outputs": "torques", "contact_mix": "ds", "source": sys.argv[3], "kernels": out}, indent=1))
