#!/usr/bin/env python
"""Summarises an `ncu --set full` report for profiles/: one row per captured kernel launch.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv && python tools/ncu_summary.py /tmp/raw.csv > profiles/rNN_ncu_xxx.md
"""
import csv
import sys

COLS = [
    ("time ms", "gpu__time_duration.sum"),
    ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("fp64 pipe %", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("smem wavefronts %", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
    ("smem bank conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    ("dram read GB", "dram__bytes_read.sum"),
    ("dram write GB", "dram__bytes_write.sum"),
    ("regs", "launch__registers_per_thread"),
    ("smem KB/CTA", "launch__shared_mem_per_block_dynamic"),
    ("CTAs/SM (smem)", "launch__occupancy_limit_shared_mem"),
    ("warp inst", "smsp__inst_executed.sum"),
    ("dfma/cycle", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed"),
    ("dadd/cycle", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed"),
    ("dmul/cycle", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed"),
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    print("| kernel | " + " | ".join(c for c, _ in COLS) + " | top stalls (warps per issue) |")
    print("|---|" + "---|" * (len(COLS) + 1))
    for r in data:
        name = r[idx["Kernel Name"]].split("(")[0]
        cells = []
        for _, key in COLS:
            if key not in idx or r[idx[key]] in ("", "n/a"):
                cells.append("-")
                continue
            v = float(r[idx[key]].replace(",", ""))
            u = units[idx[key]]
            if u == "byte":
                v, u = v / 1e9, "Gbyte"
            if u == "Mbyte":
                v, u = v / 1e3, "Gbyte"
            cells.append(f"{v:.3g}" if abs(v) < 1e6 else f"{v:.3e}")
        st = sorted(((float(r[idx[h]].replace(",", "")), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")])
                     for h in stall if r[idx[h]] not in ("", "n/a")), reverse=True)[:4]
        print(f"| {name} | " + " | ".join(cells) + " | " + ", ".join(f"{h} {v:.1f}" for v, h in st) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
