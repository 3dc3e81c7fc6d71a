"""Device time of the variant kernels (a9 reduced, a10 LQP) at B = 65536 on cuda:0, measured with the library's own
CUDA-event brackets (dwbc_batch_set_profiling).  Prints one JSON line; bench.py stays on the headline config."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libdwbc_b200 import RobotDataBatch  # noqa: E402
from tests.common import URDF, configure_product, lqp_fstars, synth_states  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    reps = 5
    dev = torch.device("cuda", 0)
    T = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev).contiguous()
    q, qd, rng = synth_states(B, 20243)
    fs = lqp_fstars(rng, B, False)
    hard = rng.choice(B, B // 20, replace=False)
    fs[0][hard, 1] = 12.0
    flags = np.ones((B, 2), dtype=np.uint8)
    rd = configure_product(RobotDataBatch(URDF, B, 0), "lqp")
    rd.update_kinematics(T(q), T(qd))
    rd.set_contact(T(flags, torch.uint8))
    rd.calc_contact_constraint()
    for l, f in enumerate(fs):
        rd.set_task_space(l, T(f))
    rd.calc_control_torque_lqp()
    torch.cuda.synchronize()
    rd.set_profiling(True)
    rd.kernel_times(reset=True)
    for _ in range(reps):
        y, tau, st = rd.calc_control_torque_lqp()
    torch.cuda.synchronize()
    kt = rd.kernel_times(reset=True)
    ms = kt["lqp_kernel"][0] / max(kt["lqp_kernel"][1], 1)
    it = rd.lqp_iterations().float()
    print(json.dumps({"kernel": "lqp_kernel", "B": B, "ms_per_launch": ms, "solves_per_s": B / ms * 1e3, "solved_fraction": (st == 4).float().mean().item(),
                      "mean_active_set_iterations": it.mean().item(), "max_active_set_iterations": it.max().item()}))


if __name__ == "__main__":
    main()
