#!/usr/bin/env python
"""Per-phase latency of lean_cycle_kernel (cycles per state, mean over the batch) from the clock64 stamps of
dwbc_batch_set_lean_timing.   python tools/lean_phases.py [B] [--mixed]  ->  JSON on stdout"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from libdwbc_b200 import OUT_TORQUES, RobotDataBatch
from tests.common import URDF, configure_product, posture_selector, synth_fstar, synth_states

NAMES = ["sweeps", "A", "A^-1", "Jc", "X", "Lambda_c", "Jbar+P", "V", "J_task", "T", "W+VV^T", "(W+VV^T)^-1", "W^+,NwJw,tau_grav", "Y", "M_l", "M_l^-1", "JL",
         "Ntorque", "QP tests", "end"]
B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 65536
dev = torch.device("cuda", 0)
q, qd, rng = synth_states(B, 20242)
fs = synth_fstar(rng, B)
flags = np.ones((B, 2), dtype=np.uint8)
if "--mixed" in sys.argv:
    u = rng.uniform(size=B)
    flags[(u >= 0.5) & (u < 0.75)] = [1, 0]
    flags[u >= 0.75] = [0, 1]
t = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
rd = configure_product(RobotDataBatch(URDF, B, 0), 2)
rd.set_outputs(OUT_TORQUES)
stamps = torch.zeros((B, 32), dtype=torch.int64, device=dev)
Jp = t(posture_selector())
dq, dqd, dfl, dfs = t(q), t(qd), t(flags, torch.uint8), [t(f) for f in fs]
for it in range(3):
    if it == 2:
        rd._ck(rd.L.dwbc_batch_set_lean_timing(rd._h, stamps.data_ptr()))
    rd.update_kinematics(dq, dqd)
    rd.set_contact(dfl)
    rd.calc_contact_constraint()
    for l in range(3):
        rd.set_task_space(l, dfs[l], Jp if l == 2 else None)
    rd.solve(0xF)
torch.cuda.synchronize()
st = stamps.cpu().numpy()
d = np.diff(st[:, :21], axis=1).astype(np.float64)
d[st[:, 1:21] == 0] = np.nan  # phases a state did not go through
out = {"B": B, "mixed": "--mixed" in sys.argv, "total_cycles_mean": float(np.nanmean(st[:, 20] - st[:, 0])),
       "phases": {NAMES[k]: float(np.nanmean(d[:, k])) for k in range(20) if not np.all(np.isnan(d[:, k]))}}
# hqp_kernel: cycles accumulated per phase (slots 21..26), slot 29 = QPs that were not decided by the x = 0 test
H = st[:, 21:30].astype(np.float64)
act = H.sum(axis=1) > 0
if act.any():
    HN = {0: "other (base, torque updates, clears)", 1: "x = 0 feasibility tests", 2: "contact columns (once per state)", 3: "contact-only solve", 4: "full assembly", 5: "solve_level (stages B, C)"}
    tot = H[act][:, :6].sum(axis=1)
    out["hqp"] = {"active_states": int(act.sum()), "total_cycles_mean": float(tot.mean()), "total_cycles_p50": float(np.median(tot)), "total_cycles_p99": float(np.percentile(tot, 99)),
                  "phases_mean": {HN[k]: float(H[act][:, k].mean()) for k in range(6)}, "qps_not_quick_mean": float(H[act][:, 8].mean()),
                  "states_with_full_solve": int((H[:, 5] > 0).sum()), "full_solve_states_total_cycles_mean": float(tot[H[act][:, 5] > 0].mean()) if (H[:, 5] > 0).any() else None}
print(json.dumps(out, indent=1))
