"""Times dwbc_batch_cycle_host (pinned host buffers, BASELINE config 2, B = 65536) for several slice counts.
Usage (GPU box): python tools/time_host_slices.py > gpurun_out/host_slices.json"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import ALL_STAGES, make_inputs, parse  # noqa: E402
from libdwbc_b200 import RobotDataBatch  # noqa: E402
from tests.common import URDF, configure_product, posture_selector  # noqa: E402

B = 65536
dev = torch.device("cuda", 0)
sets = [make_inputs(B, 20242 + j, parse([])) for j in range(2)]
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
hsets = [(pin(q), pin(qd), pin(np.concatenate([f.reshape(-1) for f in fs])), pin(fl)) for q, qd, fs, fl in sets]
hout = tuple(torch.empty((B, 33), dtype=torch.float64).pin_memory() for _ in range(3)) + (torch.empty((B,), dtype=torch.int32).pin_memory(),)
rd = configure_product(RobotDataBatch(URDF, B, 0), 2)
rd.set_task_space(2, torch.zeros((B, 21), dtype=torch.float64, device=dev), torch.as_tensor(posture_selector(), dtype=torch.float64, device=dev))
res = {}
for K in (1, 2, 4, 8, 16, 32, 64):
    rd.set_host_slices(K)
    for j in range(2):
        rd.cycle_host(ALL_STAGES, *hsets[j % 2], out=hout)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 6
    for j in range(n):
        rd.cycle_host(ALL_STAGES, *hsets[j % 2], out=hout)
    torch.cuda.synchronize()
    res[K] = 1e3 * (time.perf_counter() - t0) / n
print(json.dumps({"B": B, "ms_per_cycle_by_slices": res}))
