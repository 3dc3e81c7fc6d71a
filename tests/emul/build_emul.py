"""Builds tests/emul/build/libdwbc_emul.so: the product's kernel sources compiled for the host through
tests/emul/cuda_emul.h.  TEST INFRASTRUCTURE ONLY -- lets the GPU-less container check kernel logic against
the oracle.  The product never loads this library."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "libdwbc_b200", "csrc")
TEAM = int(os.environ.get("DWBC_EMUL_TEAM", "128"))  # CUDA threads per state in the emulation (same as the product)
OUT = os.path.join(HERE, "build", f"libdwbc_emul_t{TEAM}.so")


def build(force=False):
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "cuda_emul.h"), os.path.join(ROOT, "include", "dwbc_b200.h"), os.path.abspath(__file__)]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(s) <= os.path.getmtime(OUT) for s in srcs):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["g++", "-O2", "-g", "-std=c++20", "-fPIC", "-shared", "-DDWBC_CPU_EMUL", f"-DDWBC_TEAM={TEAM}", "-I", HERE, "-I", CSRC, "-x", "c++",
           os.path.join(CSRC, "api.cu"), os.path.join(CSRC, "probe.cu"), os.path.join(CSRC, "model.cpp"), "-o", OUT, "-lpthread", "-Wno-unused-variable", "-Wno-unused-but-set-variable"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
