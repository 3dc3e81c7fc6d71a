#!/usr/bin/env python
"""Builds a variant of the CUDA library with extra -D flags into tools/_build/<name>.so (experiments: run with DWBC_LIB=<path>).
    python tools/build_variant.py name -DLEAN_GANG=2 ..."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g

name, flags = sys.argv[1], sys.argv[2:]
csrc = os.path.join(ROOT, "libdwbc_b200", "csrc")
out = os.path.join(ROOT, "tools", "_build", name + ".so")
os.makedirs(os.path.dirname(out), exist_ok=True)
subprocess.check_call([os.environ.get("NVCC", "nvcc")] + g.NVCC_FLAGS + flags + ["-o", out, os.path.join(csrc, "api.cu"), os.path.join(csrc, "probe.cu"), os.path.join(csrc, "model.cpp")])
print(out)
