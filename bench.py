#!/usr/bin/env python
"""Benchmark of the batched whole-body-control cycle (BASELINE.json metric: TOCABI HQP-WBC solves/sec, batch 65536; p99 latency).

    python bench.py --gpus N --steps K --warmup W                      # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W     # the reference's CPU path (oracle port) on the host cores
    python bench.py --config {0,1,2,3i,3ii,4}                          # the BASELINE.json configurations (default 2, the headline)
    python bench.py --scaling strong --gpus N                          # one global batch of 65 536 states cut over N GPUs
    python bench.py --single-process --gpus N                          # one host process, dwbc_multi_cycle_host over N GPUs (host buffers)

A "step" is one pass of the hot path over one batch of synthetic TOCABI states:
  config 0   UpdateKinematics -> SetContact -> CalcContactConstraint -> CalcGravCompensation            (tests/dwbc_test.cpp:48-120)
  config 1   + CalcTaskControlTorque(hqp=false), COM position / pelvis rotation, 4096 states
  config 2   + 3-level HQP (COM / pelvis+upper-body rotation / 21-joint posture), torque + friction rows, redistribution QP   [headline]
  config 3i  reduced-dimension pipeline (Ahn 2024): ReducedDynamicsCalculate ... ReducedCalcTaskControlTorque(hqp) + redistribution
  config 3ii LQP hierarchy (Ahn 2025): ConfigureLQP + CalcControlTorqueLQP, then the reduced chain ConfigureLQP_R / _R_NC on the
             same states, tasks of tests/sp_test/jacc_compare.cpp:319-334
  config 4   config 2 with DS / left-SS / right-SS = 0.5 / 0.25 / 0.25, 131 072 states per GPU (1 048 576 over 8)
`value` is timed with the inputs resident in HBM; `e2e` goes through the host-buffer entry point (dwbc_batch_cycle_host, pinned
buffers; H2D of q, qdot, f*, flags and D2H of the torque vectors + status inside the timed region).  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

METRIC = "TOCABI HQP-WBC solves/sec"
UNIT = "solves/s"
ALL_STAGES = 0xF
N_INPUT_SETS = 4  # rotated every step: 4 x 57 MB of inputs > 126 MB of L2 (headline size)

# stages: SOLVE_GRAV 1 | SOLVE_TASK 2 | SOLVE_HQP 4 | SOLVE_REDIST 8
CONFIGS = {
    "0": dict(batch=65536, dofs=(), stages=1, mix="ds", kind=0, tlim=False,
              what="double-support gravity compensation + contact constraint (tests/dwbc_test.cpp:48-120)"),
    "1": dict(batch=4096, dofs=(3, 3), stages=1 | 2, mix="ds", kind=1, tlim=False,
              what="contact-consistent COM position + pelvis-orientation task torque, CalcTaskControlTorque(hqp=false)"),
    "2": dict(batch=65536, dofs=(3, 6, 21), stages=ALL_STAGES, mix="ds", kind=2, tlim=True,
              what="3-level HQP (COM / pelvis+upper-body rotation / 21-joint posture), torque limit 300 + friction/ZMP rows, contact redistribution QP"),
    "3i": dict(batch=65536, dofs=(3, 3, 3, 6), stages=1 | 4 | 8, mix="ds", kind="reduced", tlim=True,
               what="reduced-dimension pipeline (Ahn 2024) on the 24/18-dof model: COM / pelvis rotation / upper-body rotation / left-hand 6D, HQP + redistribution"),
    "3ii": dict(batch=65536, dofs=(6, 6), stages=0, mix="ds", kind="lqp", tlim=True,
                what="LQP hierarchy (Ahn 2025): ConfigureLQP + CalcControlTorqueLQP, then ConfigureLQP_R / _R_NC on the reduced model; COM position + pelvis rotation / left hand 6D"),
    "4": dict(batch=131072, dofs=(3, 6, 21), stages=ALL_STAGES, mix="mixed", kind=2, tlim=True,
              what="3-level HQP of config 2 with a per-state contact mix"),
}
MIXES = {"ds": "double support", "mixed": "DS/left-SS/right-SS 0.5/0.25/0.25", "ss": "left single support",
         "sorted": "DS/left-SS/right-SS 0.5/0.25/0.25 grouped by pattern"}


def parse(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="2", choices=sorted(CONFIGS), help="BASELINE.json configuration (default 2: the headline)")
    ap.add_argument("--batch", type=int, default=0, help="states per GPU (weak) or in total (strong); 0 = the configuration's size")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="weak: --batch states per GPU; strong: --batch states in total, cut over the GPUs")
    ap.add_argument("--single-process", action="store_true", help="one host process drives all GPUs through dwbc_multi_cycle_host (host buffers: e2e numbers only)")
    ap.add_argument("--contact-mix", default="", choices=["", "ds", "mixed", "ss", "sorted"],
                    help="override the configuration's contact pattern: ds, mixed (0.5/0.25/0.25), ss (left single support), sorted (mixed, grouped by pattern)")
    ap.add_argument("--outputs", default="torques", choices=["torques", "all"],
                    help="torques: DWBC_OUT_TORQUES, the one-kernel cycle (what a controller consumes); all: DWBC_OUT_ALL, every public RobotData field materialised")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling appendix of a multi-GPU weak run")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    a = ap.parse_args(argv)
    a.cfg = CONFIGS[a.config]
    if not a.batch:
        a.batch = a.cfg["batch"]
    if not a.contact_mix:
        a.contact_mix = a.cfg["mix"]
    return a


def workload_name(args):
    return f"config {args.config}: TOCABI 33-DoF, {args.cfg['what']}, {MIXES[args.contact_mix]}"


def make_inputs(B, seed, args):
    from tests.common import lqp_fstars, synth_fstar, synth_states

    q, qd, rng = synth_states(B, seed)
    dofs = args.cfg["dofs"]
    if args.config == "3ii":
        fs = lqp_fstars(rng, B, False)
    elif args.config == "3i":
        fs = [0.5 * rng.standard_normal((B, 3)), rng.standard_normal((B, 3)), rng.standard_normal((B, 3)), rng.standard_normal((B, 6))]
    elif dofs:
        fs = synth_fstar(rng, B, dofs=dofs, scales=(0.5, 1.0, 2.0)[:len(dofs)], hard_frac=float(os.environ.get("DWBC_BENCH_HARD_FRAC", "0.05")) if args.cfg["kind"] == 2 else 0.0)  # env: diagnostics only
    else:
        fs = []
    flags = np.ones((B, 2), dtype=np.uint8)
    mix = args.contact_mix
    if mix in ("mixed", "sorted"):
        u = rng.uniform(size=B)
        if mix == "sorted":
            u = np.sort(u)
        flags[(u >= 0.5) & (u < 0.75)] = [1, 0]
        flags[u >= 0.75] = [0, 1]
    elif mix == "ss":
        flags[:] = [1, 0]
    return q, qd, fs, flags


def config_dict(args, world):
    """The `config` object BOTH arms print: a function of the command line only, so the two lines of one command describe one workload."""
    B = args.batch if args.scaling == "weak" else max(1, args.batch // world)
    in_mb = B * (8 * 40 + 8 * 39 + 8 * sum(args.cfg["dofs"]) + 2) / 1e6
    return {"workload": workload_name(args), "config": args.config, "states_per_gpu": B, "global_batch": B * world,
            "scaling": args.scaling, "contact_mix": args.contact_mix, "task_dofs": list(args.cfg["dofs"]), "stages": args.cfg["stages"],
            "outputs": args.outputs if args.cfg["kind"] in (0, 1, 2) else "variant",
            "l2": f"{N_INPUT_SETS} input sets rotated per step ({N_INPUT_SETS} x {in_mb:.0f} MB; L2 is 126 MB)",
            "parallelism": f"independent states sharded over {world} GPU(s), no collective on the data path"}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi sampled every 200 ms during the timed region (the recipe's clocks line, B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            self.f = os.fdopen(fd, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            try:
                self.proc.kill()
            except Exception:
                pass
        try:
            self.f.close()
            rows = [l.strip().split(",") for l in open(self.path) if l.strip()]
            os.unlink(self.path)
        except Exception:
            return out
        sm, mx, pw = [], [], []
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            r = [x.strip() for x in r]
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                pw.append(float(r[3]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if r[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(np.max(mx)), power_w=float(np.median(pw)), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------ CPU arm (oracle port of the reference path)
def add_feet(r, product):
    from tests.common import FOOT_POINT, LEFT_FOOT, RIGHT_FOOT

    for foot in (LEFT_FOOT, RIGHT_FOOT):
        if product:
            from libdwbc_b200 import CONTACT_6D

            r.add_contact_constraint(foot, CONTACT_6D, FOOT_POINT, (0, 0, 1), 0.15, 0.075)
        else:
            r.add_contact(foot, FOOT_POINT, (0, 0, 1), 0.15, 0.075)


def oracle_setup(args):
    from oracle import oracle as orc
    from tests.common import URDF, configure_oracle

    orc.use_native()  # -O3 -march=native built on THIS host (falls back to the portable build)
    r = orc.OracleRobot(URDF)
    kind = args.cfg["kind"]
    if kind in (0, 1, 2):
        configure_oracle(r, kind)
    else:
        add_feet(r, False)
        if kind == "reduced":
            r.add_task(0, orc.TASK_LINK_POSITION, "COM")
            r.add_task(1, orc.TASK_LINK_ROTATION, "pelvis_link")
            r.add_task(2, orc.TASK_LINK_ROTATION, "upperbody_link")
            r.add_task(3, orc.TASK_LINK_6D, "L_Wrist2_Link")
        else:
            r.add_task(0, orc.TASK_LINK_POSITION, "COM")
            r.add_task(0, orc.TASK_LINK_ROTATION, "pelvis_link")
            r.add_task(1, orc.TASK_LINK_6D, "L_Wrist2_Link")
    return r, orc


def cpu_threads(args, orc):
    return orc.num_threads() if args.cfg["kind"] in (0, 1, 2) else 1


def cpu_run(args, r, orc, q, qd, fs, flags, S, nthreads):
    """The reference path (CPU oracle port) over the first S states; returns seconds."""
    kind = args.cfg["kind"]
    t0 = time.perf_counter()
    if kind in (0, 1, 2):  # the batched entry point: one state per call in reference call order on `nthreads` host threads
        r.run_batch(q[:S], qd[:S], [f[:S] for f in fs], flags[:S], stages=args.cfg["stages"], nthreads=nthreads)
        return time.perf_counter() - t0
    for b in range(S):  # the variants: single-state calls in reference call order from one thread
        r.update_kinematics(q[b], qd[b])
        r.set_contact(*[int(x) for x in flags[b]])
        for l, f in enumerate(fs):
            r.set_task(l, f[b])
        if kind == "reduced":  # tests/sp_test/redu_dyn_test.cpp:262-267
            r.reduced_dynamics()
            r.reduced_contact_constraint()
            r.reduced_grav()
            r.reduced_task_space()
            r.reduced_task_torque(hqp=True)
            r.reduced_redistribute()
        else:  # tests/sp_test/jacc_compare.cpp:367-384, 457-511
            r.calc_contact_constraint()
            r.update_task_space()
            r.configure_lqp()
            r.calc_lqp()
            r.reduced_dynamics()
            r.reduced_contact_constraint()
            r.reduced_task_space()
            r.configure_lqp(reduced=True)
            r.calc_lqp()
            r.configure_lqp_r_nc()
            r.calc_lqp_r_nc()
    return time.perf_counter() - t0


def cpu_baseline(args, q, qd, fs, flags):
    """The CPU restatement of the reference path (oracle/, kind "port": Eigen/RBDL/qpOASES/OSQP are not available offline, so the
    reference itself cannot be built) timed on this box's host cores over a bounded sample of the same workload."""
    r, orc = oracle_setup(args)
    cores = cpu_threads(args, orc)
    B = q.shape[0]
    S0 = min(B, 64 * cores)
    dt0 = cpu_run(args, r, orc, q, qd, fs, flags, S0, cores)  # also warms caches / threads
    S = int(min(B, max(S0, S0 / max(dt0, 1e-9) * args.cpu_seconds)))
    dt = cpu_run(args, r, orc, q, qd, fs, flags, S, cores)
    S1 = min(B, 128)
    dt1 = cpu_run(args, r, orc, q, qd, fs, flags, S1, 1)
    return {"value": S / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {S} of the {B} states of the workload, {cores} host thread(s), one state per call in reference call order, cold-start QPs ({dt:.1f} s); oracle {orc.flavour}",
            "single_thread_us_per_solve": 1e6 * dt1 / S1}


def run_reference(args):
    """The reference arm: every step pushes one GPU's WHOLE batch (the same states_per_gpu the CUDA arm times) through the CPU port
    of the reference path on all host threads.  The single-threaded variants (3i, 3ii) are bounded to ~5 s of CPU work per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    world = max(int(os.environ.get("WORLD_SIZE", "1")), args.gpus)
    B = args.batch if args.scaling == "weak" else max(1, args.batch // world)
    q, qd, fs, flags = make_inputs(B, 20242, args)
    r, orc = oracle_setup(args)
    cores = cpu_threads(args, orc)
    S = B
    if cores == 1:
        dt0 = cpu_run(args, r, orc, q, qd, fs, flags, min(B, 32), 1)
        S = int(min(B, max(32, 5.0 * min(B, 32) / dt0)))
    for _ in range(min(args.warmup, 1)):
        cpu_run(args, r, orc, q, qd, fs, flags, min(S, 64 * cores), cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_run(args, r, orc, q, qd, fs, flags, S, cores)
    dt = time.perf_counter() - t0
    val = S * args.steps / dt
    sample = (f"each step = {'all' if S == B else 'the first ' + str(S) + ' of the'} {B} states of one GPU's batch through the CPU restatement of the "
              f"reference path (oracle port; the reference itself needs Eigen/RBDL/qpOASES/OSQP, absent offline), {cores} host thread(s); oracle {orc.flavour}")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, world), "details": {"same_config": S == B, "states_per_step": S},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------ GPU arm
class Product:
    """The product-side workload of one configuration: set up once, step(j) on device-resident inputs, e2e_step(j) from pinned
    host buffers.  `backend` is only passed by tests/test_bench_host.py (the CPU emulation of the kernels)."""

    def __init__(self, args, B, device, seeds, backend=None):
        from libdwbc_b200 import OUT_ALL, OUT_TORQUES, TASK_LINK_6D, TASK_LINK_POSITION, TASK_LINK_ROTATION, RobotDataBatch
        from tests.common import URDF, configure_product, posture_selector

        self.args, self.B = args, B
        kind = self.kind = args.cfg["kind"]
        rd = RobotDataBatch(URDF, B, device, _backend=backend) if backend is not None else RobotDataBatch(URDF, B, device)
        if kind in (0, 1, 2):
            configure_product(rd, kind)
        else:
            add_feet(rd, True)
            if kind == "reduced":
                rd.add_task_space(0, TASK_LINK_POSITION, "COM")
                rd.add_task_space(1, TASK_LINK_ROTATION, "pelvis_link")
                rd.add_task_space(2, TASK_LINK_ROTATION, "upperbody_link")
                rd.add_task_space(3, TASK_LINK_6D, "L_Wrist2_Link")
            else:
                rd.add_task_space(0, TASK_LINK_POSITION, "COM")
                rd.add_task_space(0, TASK_LINK_ROTATION, "pelvis_link")
                rd.add_task_space(1, TASK_LINK_6D, "L_Wrist2_Link")
        self.rd, self.be = rd, rd.be
        self.main_path = kind in (0, 1, 2)
        if self.main_path:
            rd.set_outputs(OUT_TORQUES if args.outputs == "torques" else OUT_ALL)
        self.stages = args.cfg["stages"]
        D = self.be.from_numpy
        self.Jp = D(posture_selector()) if kind == 2 else None
        self.sets = [make_inputs(B, s, args) for s in seeds]
        self.dsets = [(D(q), D(qd), [D(f) for f in fs], D(fl)) for q, qd, fs, fl in self.sets]
        E = self.be.empty
        self.out = (E((B, 33)), E((B, 33)), E((B, 33)), E((B,), "i4"))
        self.lqp_out = self.lqp_r_out = self.lqp_nc_out = None

    def step(self, j, dset=None):
        rd, kind = self.rd, self.kind
        q, qd, fs, fl = dset if dset is not None else self.dsets[j % len(self.dsets)]
        rd.update_kinematics(q, qd)
        rd.set_contact(fl)
        if kind == "reduced":
            rd.reduced_dynamics_calculate((1, 1))
            rd.reduced_calc_contact_constraint()
            for l, f in enumerate(fs):
                rd.set_task_space(l, f)
            rd.reduced_solve(self.stages, out=self.out)
            return
        rd.calc_contact_constraint()
        for l, f in enumerate(fs):
            rd.set_task_space(l, f, self.Jp if (kind == 2 and l == 2) else None)
        if kind == "lqp":
            self.lqp_out = rd.calc_control_torque_lqp(self.lqp_out)
            rd.reduced_dynamics_calculate((1, 1))
            rd.reduced_calc_contact_constraint()
            rd.reduced_solve(1 | 2, out=self.out)
            self.lqp_r_out = rd.calc_control_torque_lqp_r(self.lqp_r_out)
            self.lqp_nc_out = rd.calc_control_torque_lqp_r_nc(self.lqp_nc_out)
            return
        rd.solve(self.stages, out=self.out)

    def solved_fraction(self):
        N = self.be.to_numpy
        if self.kind == "lqp":  # all levels solved in the full, the reduced and the non-contact hierarchy
            return float(np.mean((N(self.lqp_out[2]) == 4) & (N(self.lqp_r_out[2]) == 3) & (N(self.lqp_nc_out[2]) == 2)))
        want = 1 | (2 if self.stages & 6 else 0) | (4 if self.stages & 8 else 0)
        return float(np.mean((N(self.out[3]) & want) == want))

    def qp_iterations_mean(self):
        if self.kind == "lqp":
            return float(self.be.to_numpy(self.rd.lqp_iterations()).mean())
        if self.main_path and self.stages & 12:
            return float(self.be.to_numpy(self.rd.qp_iterations()).mean())
        return 0.0

    # ---- end to end from pinned host buffers ----
    def prepare_e2e(self):
        import torch

        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        B = self.B
        self.hsets = []
        for q, qd, fs, fl in self.sets[:2]:
            fcat = pin(np.concatenate([f.reshape(-1) for f in fs])) if fs else None
            self.hsets.append((pin(q), pin(qd), [pin(f) for f in fs], pin(fl), fcat))
        self.hout = (torch.empty((B, 33), dtype=torch.float64).pin_memory(), torch.empty((B, 33), dtype=torch.float64).pin_memory(),
                     torch.empty((B, 33), dtype=torch.float64).pin_memory(), torch.empty((B,), dtype=torch.int32).pin_memory())
        if not self.main_path:  # the variants have no host-buffer entry point: explicit staging (H2D + the variant's calls + D2H, all timed)
            q, qd, fs, fl = self.dsets[0]
            self.stage = (q.clone(), qd.clone(), [f.clone() for f in fs], fl.clone())

    def e2e_step(self, j):
        import torch

        q, qd, fs, fl, fcat = self.hsets[j % len(self.hsets)]
        if self.main_path:
            self.rd.cycle_host(self.stages, q, qd, fcat, fl, out=self.hout)
            return
        sq, sqd, sfs, sfl = self.stage
        sq.copy_(q, non_blocking=True)
        sqd.copy_(qd, non_blocking=True)
        sfl.copy_(fl, non_blocking=True)
        for d, h in zip(sfs, fs):
            d.copy_(h, non_blocking=True)
        self.step(0, self.stage)
        if self.kind == "lqp":
            self.hout[0].copy_(self.lqp_out[1], non_blocking=True)
            self.hout[1].copy_(self.lqp_nc_out[1], non_blocking=True)
            self.hout[3].copy_(self.lqp_nc_out[2], non_blocking=True)
        else:
            for h, d in zip(self.hout, self.out):
                h.copy_(d, non_blocking=True)
        torch.cuda.synchronize()

    def io_bytes(self):
        n_in = 8 * 40 + 8 * 39 + 8 * sum(self.args.cfg["dofs"]) + 2
        n_out = (2 * 33 * 8 + 4) if self.kind == "lqp" else (3 * 33 * 8 + 4)
        return n_in, n_out


def flops_per_kernel(args, iters_mean):
    """Algorithmic flops of one solve by kernel (the reference's dense as-written counts, libdwbc_b200/flops_model.py), contact
    dimension averaged over the mix."""
    from libdwbc_b200.flops_model import lqp_flops, per_kernel_flops, reduced_flops

    dofs, st, kind = args.cfg["dofs"], args.cfg["stages"], args.cfg["kind"]
    if kind == "reduced":
        return reduced_flops(dofs, c=12, qp_iters_total=None)
    if kind == "lqp":
        base = per_kernel_flops((), c=12, hqp=False, redist=False)
        red = reduced_flops(dofs, c=12, hqp=False, redist=False)
        return {"rbd_kernel": base["rbd_kernel"], "ainv_kernel": base["ainv_kernel"], "contact_jacobian_kernel": base["contact_jacobian_kernel"],
                "contact_dynamics_kernel": base["contact_dynamics_kernel"], "reduced_kernels": red["reduced_kernels"],
                "lqp_kernel": lqp_flops(dofs) + lqp_flops(dofs[:1], n=24, m=18) + lqp_flops((6,), c=0, n=21, m=21)}
    kw = dict(hqp=bool(st & 4), redist=bool(st & 8), tlim=args.cfg["tlim"], qp_iters_total=iters_mean if st & 12 else None)
    fl_ds, fl_ss = per_kernel_flops(dofs, c=12, **kw), per_kernel_flops(dofs, c=6, **kw)
    wds = {"ds": 1.0, "ss": 0.0}.get(args.contact_mix, 0.5)
    return {k: wds * fl_ds[k] + (1 - wds) * fl_ss[k] for k in fl_ds}


DENSE = ("rbd_kernel", "ainv_kernel", "contact_jacobian_kernel", "contact_dynamics_kernel", "task_space_kernel")


def run_b200(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.single_process:
        return run_single_process(args)
    if world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    if not torch.cuda.is_available():
        print("bench.py: no CUDA device (the product has no CPU fallback); use --impl reference for the CPU arm", file=sys.stderr)
        return 2
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from libdwbc_b200._lib import load_library

    L = load_library()
    B = args.batch if args.scaling == "weak" else max(1, args.batch // world)
    # independent states per rank (no collective on the data path): rank r owns seeds 20242 + 1000 r + set
    P = Product(args, B, local, [20242 + 1000 * rank + j for j in range(N_INPUT_SETS)])
    rd = P.rd

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(prod, steps, warm):
        for j in range(warm):
            prod.step(j)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        barrier()
        evs[0].record()
        for k in range(steps):
            prod.step(warm + k)
            evs[k + 1].record()
        barrier()
        return evs[0].elapsed_time(evs[-1]), np.array([evs[k].elapsed_time(evs[k + 1]) for k in range(steps)])

    for j in range(args.warmup):
        P.step(j)
    barrier()
    iters_mean = P.qp_iterations_mean()
    ok_frac = P.solved_fraction()
    path = "lean_cycle_kernel + hqp_kernel over the QP-active states" if rd.last_path() == 1 else "generic kernels"

    # ---- timed region: device-resident inputs ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    rd.kernel_times(reset=True)
    rd.set_profiling(True)
    n0 = L.dwbc_launch_count()
    ms_total, per_step = timed(P, args.steps, 0)
    launches = L.dwbc_launch_count() - n0
    rd.set_profiling(False)
    ktimes = rd.kernel_times(reset=True)

    # ---- e2e: host buffers ----
    P.prepare_e2e()
    e2e_steps = max(3, min(args.steps, 10))
    for j in range(2):
        P.e2e_step(j)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    e2e_lat = []
    for j in range(e2e_steps):
        t1 = time.perf_counter()
        P.e2e_step(j)
        e2e_lat.append(1e3 * (time.perf_counter() - t1))
    e1.record()
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0))
    clocks = sampler.stop() if rank == 0 else None

    # ---- strong-scaling appendix of a multi-GPU weak run: the headline batch (65 536 states in total) cut over the GPUs ----
    strong = None
    if world > 1 and args.scaling == "weak" and not args.no_strong and args.config in ("2", "4"):
        Bs = max(1, 65536 // world)
        Ps = Product(args, Bs, local, [30242 + 1000 * rank + j for j in range(N_INPUT_SETS)])
        s_total, s_per = timed(Ps, max(args.steps, 20), 3)
        strong = (Bs, s_total, s_per)
        del Ps

    # ---- max over ranks ----
    vals = [ms_total, e2e_ms, float(np.percentile(per_step, 50)), float(np.percentile(per_step, 99))]
    if strong:
        vals += [strong[1], float(np.percentile(strong[2], 50)), float(np.percentile(strong[2], 99))]
    red = torch.tensor(vals, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
    red = [float(x) for x in red.tolist()]
    ms_total, e2e_ms, p50, p99 = red[:4]
    lsum = torch.tensor([float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(lsum, op=dist.ReduceOp.SUM)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    value = world * B * args.steps / (ms_total * 1e-3)
    e2e_val = world * B * e2e_steps / (e2e_ms * 1e-3)
    h2d_in, d2h_out = P.io_bytes()

    # ---- roofline of the dominant kernel (FP64 pipe) ----
    flk = flops_per_kernel(args, iters_mean)
    step_flops = sum(flk.values())
    if ktimes.get("lean_cycle_kernel", (0, 0))[0] > 0:  # the one-kernel cycle does the work of the five dense kernels
        flk["lean_cycle_kernel"] = sum(flk.get(k, 0.0) for k in DENSE)
    ksum = max(sum(v[0] for v in ktimes.values()), 1e-12)
    kern = {k: {"ms_per_step": ms / max(args.steps, 1), "launches_per_step": cnt / max(args.steps, 1), "share": ms / ksum} for k, (ms, cnt) in ktimes.items() if ms > 0}
    dom = max(kern, key=lambda k: kern[k]["ms_per_step"])
    tf = None
    try:
        import ctypes as C

        t = C.c_double()
        if L.dwbc_probe_fp64(local, 0.5, C.byref(t)) == 0:
            tf = t.value
    except Exception:
        pass
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    fp64_peak = tf if tf else 37.0
    for k in kern:  # every kernel's own fraction, so that a change of the dominant kernel between runs stays readable
        if flk.get(k, 0.0) > 0:
            kern[k]["achieved_tflops"] = B * flk[k] / (kern[k]["ms_per_step"] * 1e-3) / 1e12
            kern[k]["frac"] = kern[k]["achieved_tflops"] / fp64_peak
    achieved = kern[dom].get("achieved_tflops")
    # DRAM traffic per step of the dominant kernel from the committed `ncu --set full` capture of the same workload, else null
    traffic, traffic_src, step_traffic = None, None, None
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", name)))
            same = (tj.get("states_per_launch") == B and str(tj.get("config", "2")) == args.config and tj.get("outputs", "all") == args.outputs
                    and tj.get("contact_mix", "ds") == args.contact_mix)
            if same and dom in tj["kernels"]:
                traffic = tj["kernels"][dom]["dram_read_bytes"] + tj["kernels"][dom]["dram_write_bytes"]
                step_traffic = sum(v["dram_read_bytes"] + v["dram_write_bytes"] for v in tj["kernels"].values())
                traffic_src = f"profiles/{name}: " + tj["source"]
                break
        except Exception:
            pass
    step_s = ms_total / args.steps * 1e-3
    roofline = {"bound": "fp64", "kernel": dom, "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": (achieved / fp64_peak) if achieved else None,
                "traffic": traffic, "traffic_source": traffic_src, "traffic_whole_step": step_traffic,
                "peak_source": "DFMA-chain probe in this run (dwbc_probe_fp64); MEASURED_PEAKS.json has no FP64 figure" if tf else "nominal",
                "algorithmic_flops_per_solve": step_flops, "kernel_flops_per_solve": flk.get(dom),
                "whole_step": {"achieved": B * step_flops / step_s / 1e12, "frac": B * step_flops / step_s / 1e12 / fp64_peak},
                "hbm": {"io_bytes_per_solve": h2d_in + d2h_out, "achieved_gbs": B * (h2d_in + d2h_out) / step_s / 1e9, "peak_gbs": hbm_peak, "peak_source": hbm_src},
                "kernels": kern, "mean_qp_iterations_per_solve": iters_mean}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, world),
            "details": {"solved_fraction": ok_frac, "path": path if P.main_path else "variant kernels (reduced_kernels / lqp_kernel)",
                        "workspace_bytes_per_state": rd.workspace_bytes() / B},
            "latency_ms": {"p50": p50, "p99": p99, "note": "one batched cycle of states_per_gpu states per GPU, device-resident inputs, max over ranks"},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": B * h2d_in, "d2h_bytes_per_step": B * d2h_out, "steps": e2e_steps,
                    "ms_per_step": e2e_ms / e2e_steps, "latency_ms": {"p50": float(np.percentile(e2e_lat, 50)), "max": float(np.max(e2e_lat))},
                    "api": "dwbc_batch_cycle_host (pinned host buffers)" if P.main_path else "pinned H2D + the variant's entry points + D2H"},
            "gpu_launches": int(lsum.item()), "clocks": clocks, "roofline": roofline}
    if strong:
        steps_s = len(strong[2])
        line["strong_scaling"] = {"global_batch": strong[0] * world, "states_per_gpu": strong[0], "value": world * strong[0] * steps_s / (red[4] * 1e-3), "unit": UNIT,
                                  "ms_per_step": red[4] / steps_s, "latency_ms": {"p50": red[5], "p99": red[6]},
                                  "note": "the headline batch of 65 536 states cut over the GPUs of this run (device-resident inputs, max over ranks)"}
    if world == 1 and not args.no_cpu_baseline:
        q, qd, fs, fl = P.sets[0]
        line["cpu_baseline"] = cpu_baseline(args, q, qd, fs, fl)
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_single_process(args):
    """One host process, one host batch: dwbc_multi_cycle_host cuts it over --gpus GPUs (one host thread + stream + handle per GPU
    inside the library) and gathers the results in place.  Host buffers only, so every number here is end to end."""
    from libdwbc_b200 import CONTACT_6D, OUT_TORQUES, TASK_LINK_POSITION, TASK_LINK_ROTATION, RobotDataMultiBatch
    from tests.common import FOOT_POINT, LEFT_FOOT, RIGHT_FOOT, URDF, posture_selector

    if args.cfg["kind"] != 2:
        print("bench.py --single-process: configurations 2 and 4", file=sys.stderr)
        return 2
    G = args.gpus
    B = args.batch * G if args.scaling == "weak" else args.batch
    mh = RobotDataMultiBatch(URDF, B, G)
    mh.add_contact_constraint(LEFT_FOOT, CONTACT_6D, FOOT_POINT, (0, 0, 1), 0.15, 0.075)
    mh.add_contact_constraint(RIGHT_FOOT, CONTACT_6D, FOOT_POINT, (0, 0, 1), 0.15, 0.075)
    mh.add_task_space(0, TASK_LINK_POSITION, "COM")
    mh.add_task_space(1, TASK_LINK_ROTATION, "pelvis_link")
    mh.add_task_space(1, TASK_LINK_ROTATION, "upperbody_link")
    mh.add_custom_task_space(2, 21, posture_selector())
    mh.set_torque_limit(300.0)
    mh.set_outputs(OUT_TORQUES)
    sets = []
    for j in range(2):
        q, qd, fs, fl = make_inputs(B, 20242 + j, args)
        hq, hqd, hfl, hf = mh.pinned((B, 40)), mh.pinned((B, 39)), mh.pinned((B, 2), np.uint8), mh.pinned((B * 30,))
        hq[:], hqd[:], hfl[:] = q, qd, fl
        hf[:] = np.concatenate([f.reshape(-1) for f in fs])
        sets.append((hq, hqd, hf, hfl))
    out = (mh.pinned((B, 33)), mh.pinned((B, 33)), mh.pinned((B, 33)), mh.pinned((B,), np.int32))
    sampler = ClockSampler(0)
    for j in range(max(args.warmup, 2)):
        mh.cycle_host(ALL_STAGES, *sets[j % 2], out=out)
    sampler.start()
    lat = []
    t0 = time.perf_counter()
    for j in range(args.steps):
        t1 = time.perf_counter()
        mh.cycle_host(ALL_STAGES, *sets[j % 2], out=out)
        lat.append(1e3 * (time.perf_counter() - t1))
    dt = time.perf_counter() - t0
    clocks = sampler.stop()
    val = B * args.steps / dt
    ok = float(((out[3] & 7) == 7).mean())
    line = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": G, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, G),
            "details": {"host": f"ONE host process: dwbc_multi_cycle_host, one host thread + stream + handle per GPU ({G}), contiguous slices, results gathered in place",
                        "solved_fraction": ok, "shard_ms_last": mh.last_shard_ms(), "shards": mh.shards()},
            "latency_ms": {"p50": float(np.percentile(lat, 50)), "p99": float(np.percentile(lat, 99)), "note": "host wall time of the gathered call (H2D + cycle + D2H on every GPU)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": B * (8 * 40 + 8 * 39 + 8 * 30 + 2), "d2h_bytes_per_step": B * (3 * 33 * 8 + 4), "steps": args.steps,
                    "ms_per_step": 1e3 * dt / args.steps, "api": "dwbc_multi_cycle_host (pinned host buffers from dwbc_host_alloc)"},
            "gpu_launches": None, "clocks": clocks, "roofline": None}
    emit(line)
    return 0


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the real stdout; everything libraries print meanwhile (e.g. NCCL's version banner) was sent
    to stderr by main()."""
    out = json.dumps(line) + "\n"
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, out.encode())
    else:
        sys.stdout.write(out)
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # fd 1 -> stderr for the duration of the run
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
