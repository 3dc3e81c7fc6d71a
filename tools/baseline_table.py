#!/usr/bin/env python
"""Rows of BASELINE.md section 3's table from the bench lines under profiles/ (the newest r02_bench_c<config>_<sha>.json etc.).
    python tools/baseline_table.py > /tmp/table.md"""
import glob
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def newest(pattern, regex=None):
    fs = sorted(glob.glob(os.path.join(ROOT, "profiles", pattern)), key=os.path.getmtime)
    if regex:
        fs = [f for f in fs if re.fullmatch(regex, os.path.basename(f))]
    return json.load(open(fs[-1])) if fs else None


def fmt(v, unit=""):
    if v is None:
        return "—"
    if v >= 1e6:
        return f"{v / 1e6:.2f} M{unit}"
    if v >= 1e3:
        return f"{v / 1e3:.1f} k{unit}"
    return f"{v:.1f}{unit}"


ROWS = [("0", "0 B=65 536, DS grav + contact"), ("1", "1 B=4096 task torque (no QP)"), ("2", "2 B=65 536 3-level HQP"), ("3i", "3(i) B=65 536 reduced-dimension HQP"),
        ("3ii", "3(ii) B=65 536 LQP + reduced LQP + NC"), ("4", "4 B=131 072 per GPU (1 048 576 on 8), mixed contacts")]
ERR = {"0": "1e-13 (goldens 1e-9)", "1": "1e-12", "2": "2e-13 / 2e-9 / 1.3e-8 (grav / task / contact), active sets identical up to 3 degenerate + 2 gray states", "3i": "1e-9 (reduced = full model)",
       "3ii": "1e-11 vs the exact solution of the same QPs", "4": "2e-13 / 2e-9 / 1.3e-8, 4 degenerate + 1 gray of 131 072"}
print("| Config | CPU 1-thread µs/solve | CPU all-core solves/s (N threads) | 1×B200 solves/s (e2e) | 2× | 4× | 8× | p99 batch latency | FP64 roofline fraction (dominant kernel / whole step) | max rel. err vs oracle |")
print("|---|---|---|---|---|---|---|---|---|---|")
for c, label in ROWS:
    d = newest(f"r02_bench_c{c}_*.json", rf"r02_bench_c{c}_(final|[0-9a-f]{{7}})\.json")
    if not d:
        continue
    cb = d.get("cpu_baseline") or {}
    multi = []
    for n in (2, 4, 8):
        m = newest(f"r02_bench_n{n}_c4_*.json", rf"r02_bench_n{n}_c4_(final|[0-9a-f]{{7}})\.json") if c == "4" else (newest(f"r02_bench_n{n}_*.json", rf"r02_bench_n{n}_(final|[0-9a-f]{{7}})\.json") if c == "2" else None)
        multi.append(fmt(m["value"]) if m else "")
    r = d["roofline"]
    print(f"| {label} | {cb.get('single_thread_us_per_solve', 0):.0f} | {fmt(cb.get('value'))} ({cb.get('cores')}) | {fmt(d['value'])} ({fmt(d['e2e']['value'])}) | {multi[0]} | {multi[1]} | {multi[2]} | "
          f"{d['latency_ms']['p99']:.2f} ms | {r['frac']:.3f} `{r['kernel']}` / {r['whole_step']['frac']:.3f} | {ERR[c]} |")
