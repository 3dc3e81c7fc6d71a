#!/usr/bin/env python
"""Stall samples and executed instructions of one kernel of an ncu report, attributed to source lines.
The SASS rows of `ncu --page source --csv` are matched in order with `nvdisasm -g -c` of the same build (needs -lineinfo).
    python tools/ncu_hot_lines.py gpurun_out/prof.ncu-rep task_space_kernel libdwbc_b200/libdwbc_b200.so [top]"""
import collections
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile

rep, kern, so = sys.argv[1], sys.argv[2], os.path.abspath(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
by_inst = len(sys.argv) > 5 and sys.argv[5] == "inst"  # sort by executed instructions instead of stall samples
mangled = os.environ.get("FUNC_FILTER", "")  # extra substring of the mangled name (template instantiations share the kernel name)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = {h: i for i, h in enumerate(rows[hdr])}
sass = []
for r in rows[hdr + 1:]:
    if len(r) < len(H) or not r[0].startswith("0x"):
        break  # first launch only
    sass.append((r[H["Source"]].strip(), int(r[H["# Samples"]] or 0), int(r[H["Instructions Executed"]] or 0)))
d = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", so], cwd=d, stdout=subprocess.DEVNULL)
lines = None
for cubin in glob.glob(d + "/*.cubin"):
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    cur, func, acc = None, None, []
    for line in dis.splitlines():
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\.text\.(\S+):", line)
        if m:
            func = m.group(1)
        if func and kern in func and mangled in func and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            acc.append(cur)
    if acc:
        lines = acc
if lines is None or abs(len(lines) - len(sass)) > 8:
    sys.exit(f"instruction count mismatch: report {len(sass)} vs disassembly {0 if lines is None else len(lines)} (different build?)")
samp, inst = collections.Counter(), collections.Counter()
for (src, s, i), ln in zip(sass, lines):
    samp[ln] += s
    inst[ln] += i
ts, ti = sum(samp.values()), sum(inst.values())
print(f"{kern}: {len(sass)} SASS instructions, {ts} stall samples, {ti:.3e} warp instructions executed")
print("| source line | samples % | warp inst % |")
print("|---|---|---|")
for ln, _ in (inst if by_inst else samp).most_common(top):
    print(f"| {ln[0]}:{ln[1]} | {100.0 * samp[ln] / ts:.1f} | {100.0 * inst[ln] / ti:.1f} |")
