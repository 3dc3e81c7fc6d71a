#!/usr/bin/env python
"""SASS instruction count of one kernel attributed to source files / line buckets (needs -lineinfo).
    python tools/sass_by_source.py libdwbc_b200/libdwbc_b200.so hqp_kernel [file.cuh:lo-hi=label ...]"""
import collections
import glob
import os
import re
import subprocess
import sys
import tempfile

so, kern = os.path.abspath(sys.argv[1]), sys.argv[2]
buckets = []
for a in sys.argv[3:]:
    rng, label = a.split("=")
    f, lh = rng.split(":")
    lo, hi = lh.split("-")
    buckets.append((f, int(lo), int(hi), label))
d = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", so], cwd=d, stdout=subprocess.DEVNULL)
for cubin in glob.glob(d + "/*.cubin"):
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    cur, func = None, None
    cnt = collections.Counter()
    for line in dis.splitlines():
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\.text\.(\S+):", line)
        if m:
            func = m.group(1)
        if func and kern in func and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            cnt[cur] += 1
    if not cnt:
        continue
    byfile = collections.Counter()
    byb = collections.Counter()
    for (fn, l), c in cnt.items():
        byfile[fn] += c
        for f, lo, hi, label in buckets:
            if fn == f and lo <= l <= hi:
                byb[label] += c
    print("total", sum(cnt.values()), dict(byfile.most_common()))
    if byb:
        print(dict(byb.most_common()))
    top = sorted(cnt.items(), key=lambda kv: -kv[1])[:25]
    print("top lines:", [(f"{k[0]}:{k[1]}", v) for k, v in top])
