"""bench.py's host logic on the CPU emulation of the kernels: every --config builds, steps, reports its solved fraction, and the
CPU arm of the same configuration runs on the same states (the numbers themselves are measured on the GPU box only)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from tests.emul.emul_backend import EmulBackend

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


@pytest.mark.parametrize("config", ["0", "1", "2", "3i", "3ii", "4"])
def test_bench_product_steps_on_emulation(config):
    args = bench.parse(["--config", config, "--batch", "2"])
    P = bench.Product(args, 2, 0, [777], backend=EmulBackend())
    P.step(0)
    assert P.solved_fraction() == 1.0, (config, P.solved_fraction())
    fl = bench.flops_per_kernel(args, P.qp_iterations_mean())
    assert sum(fl.values()) > 1e5
    # the CPU arm of the same configuration on the same states
    r, orc = bench.oracle_setup(args)
    q, qd, fs, flg = P.sets[0]
    assert bench.cpu_run(args, r, orc, q, qd, fs, flg, 2, 1) > 0
    # both arms print the same `config` object
    assert bench.config_dict(args, 1) == bench.config_dict(bench.parse(["--config", config, "--batch", "2", "--impl", "reference"]), 1)


def test_bench_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "2", "--batch", "64", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["details"]["same_config"] and line["config"]["states_per_gpu"] == 64
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
