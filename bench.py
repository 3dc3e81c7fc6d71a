#!/usr/bin/env python
"""Benchmark of the batched whole-body-control cycle (BASELINE.json metric: TOCABI HQP-WBC solves/sec, batch 65536).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path (oracle port), host cores

A "step" is one pass of the hot path over one batch of B synthetic TOCABI states per GPU:
UpdateKinematics -> SetContact -> CalcContactConstraint -> SetTaskSpace x3 -> CalcGravCompensation ->
CalcTaskControlTorque(hqp) -> CalcContactRedistribute (BASELINE.json configs[2]).
`value` is timed with the inputs resident in HBM; `e2e` goes through dwbc_batch_cycle_host with pinned HOST buffers
(H2D of q, qdot, f*, flags and D2H of three torque vectors + status inside the timed region).
One JSON line is printed by rank 0.
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
TASK_DOFS = (3, 6, 21)
ALL_STAGES = 0xF
N_INPUT_SETS = 4  # rotated every step: 4 x 57 MB of inputs > 126 MB of L2


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=65536, help="states per GPU (weak scaling)")
    ap.add_argument("--contact-mix", default="ds", choices=["ds", "mixed", "ss", "sorted"],
                    help="ds: all double support (config 2); mixed: DS/left/right 0.5/0.25/0.25 (config 4); ss: all left single support; sorted: the mixed batch grouped by contact pattern (diagnostics)")
    ap.add_argument("--outputs", default="torques", choices=["torques", "all"],
                    help="torques: DWBC_OUT_TORQUES, the one-kernel cycle (what a controller consumes); all: DWBC_OUT_ALL, every public RobotData field materialised (generic kernels)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    return ap.parse_args()


def workload_name(args):
    mix = {"ds": "double support", "mixed": "DS/left-SS/right-SS 0.5/0.25/0.25", "ss": "left single support",
           "sorted": "DS/left-SS/right-SS 0.5/0.25/0.25 grouped by pattern"}[args.contact_mix]
    return f"TOCABI 33-DoF, batch {args.batch} states per GPU, 3-level HQP (COM / pelvis+upper-body rotation / 21-joint posture), torque limit 300 + friction/ZMP rows, contact redistribution QP, {mix}"


def make_inputs(B, seed, mix):
    from tests.common import synth_fstar, synth_states

    q, qd, rng = synth_states(B, seed)
    fs = synth_fstar(rng, B, dofs=TASK_DOFS)
    flags = np.ones((B, 2), dtype=np.uint8)
    if mix in ("mixed", "sorted"):
        u = rng.uniform(size=B)
        if mix == "sorted":
            u = np.sort(u)
        flags[(u >= 0.5) & (u < 0.75)] = [1, 0]
        flags[u >= 0.75] = [0, 1]
    elif mix == "ss":
        flags[:] = [1, 0]
    return q, qd, fs, flags


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


# ------------------------------------------------------------------------------------------------ CPU arm
def oracle_robot():
    from oracle import oracle as orc
    from tests.common import URDF, configure_oracle

    return configure_oracle(orc.OracleRobot(URDF), 2), orc


def cpu_time_sample(ref, q, qd, fs, flags, S, nthreads):
    t0 = time.perf_counter()
    o = ref.run_batch(q[:S], qd[:S], [f[:S] for f in fs], flags[:S], stages=ALL_STAGES, nthreads=nthreads)
    return time.perf_counter() - t0, o


def cpu_baseline(args, q, qd, fs, flags):
    """The CPU restatement of the reference path (oracle/, kind "port": Eigen/RBDL/qpOASES are not available offline, so
    the reference itself cannot be built) timed on this box's host cores over a bounded sample of the same workload."""
    ref, orc = oracle_robot()
    cores = orc.num_threads()
    B = q.shape[0]
    S0 = min(B, 64 * cores)
    dt0, _ = cpu_time_sample(ref, q, qd, fs, flags, S0, cores)  # also warms caches / threads
    rate = S0 / max(dt0, 1e-9)
    S = int(min(B, max(S0, rate * args.cpu_seconds)))
    dt, _ = cpu_time_sample(ref, q, qd, fs, flags, S, cores)
    S1 = min(B, 256)
    dt1, _ = cpu_time_sample(ref, q, qd, fs, flags, S1, 1)
    return {"value": S / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {S} states of the workload, all {cores} host threads, one state per call in reference call order, cold-start QPs ({dt:.1f} s)",
            "single_thread_us_per_solve": 1e6 * dt1 / S1}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    q, qd, fs, flags = make_inputs(min(args.batch, 65536), 20242, args.contact_mix)
    ref, orc = oracle_robot()
    cores = orc.num_threads()
    B = q.shape[0]
    S0 = min(B, 32 * cores)
    dt0, _ = cpu_time_sample(ref, q, qd, fs, flags, S0, cores)
    S = int(min(B, max(S0, (S0 / dt0) * 2.0)))  # ~2 s of CPU work per step
    for _ in range(args.warmup):
        cpu_time_sample(ref, q, qd, fs, flags, S, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.run_batch(q[:S], qd[:S], [f[:S] for f in fs], flags[:S], stages=ALL_STAGES, nthreads=cores)
    dt = time.perf_counter() - t0
    val = S * args.steps / dt
    sample = f"each step = first {S} states of the workload through the CPU restatement of the reference path (oracle port), {cores} host threads"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args), "states_per_step": S},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------ GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
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

    from libdwbc_b200 import RobotDataBatch
    from libdwbc_b200._lib import load_library
    from libdwbc_b200.flops_model import io_bytes_per_solve, per_kernel_flops
    from tests.common import URDF, configure_product, posture_selector

    L = load_library()
    B = args.batch
    # independent states per rank (no collective on the data path): rank r owns seeds 20242 + 1000 r + set
    sets = [make_inputs(B, 20242 + 1000 * rank + j, args.contact_mix) for j in range(N_INPUT_SETS)]
    rd = configure_product(RobotDataBatch(URDF, B, local), 2)
    from libdwbc_b200 import OUT_ALL, OUT_TORQUES
    rd.set_outputs(OUT_TORQUES if args.outputs == "torques" else OUT_ALL)
    Jp = torch.as_tensor(posture_selector(), dtype=torch.float64, device=dev)
    t64 = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    dsets = [(t64(q), t64(qd), [t64(f) for f in fs], torch.as_tensor(fl, device=dev)) for q, qd, fs, fl in sets]
    out = (torch.empty((B, 33), dtype=torch.float64, device=dev), torch.empty((B, 33), dtype=torch.float64, device=dev),
           torch.empty((B, 33), dtype=torch.float64, device=dev), torch.empty((B,), dtype=torch.int32, device=dev))

    def step(j):
        q, qd, fs, fl = dsets[j % N_INPUT_SETS]
        rd.update_kinematics(q, qd)
        rd.set_contact(fl)
        rd.calc_contact_constraint()
        for l in range(3):
            rd.set_task_space(l, fs[l], Jp if l == 2 else None)
        rd.solve(ALL_STAGES, out=out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for j in range(args.warmup):
        step(j)
    barrier()
    iters_mean = float(rd.qp_iterations().double().mean().item())
    ok_frac = float(((out[3] & 7) == 7).double().mean().item())

    # ---- timed region: device-resident inputs ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    rd.kernel_times(reset=True)
    rd.set_profiling(True)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    n0 = L.dwbc_launch_count()
    barrier()
    evs[0].record()
    for k in range(args.steps):
        step(args.warmup + k)
        evs[k + 1].record()
    barrier()
    launches = L.dwbc_launch_count() - n0
    rd.set_profiling(False)
    ms_total = evs[0].elapsed_time(evs[-1])
    per_step = np.array([evs[k].elapsed_time(evs[k + 1]) for k in range(args.steps)])
    ktimes = rd.kernel_times(reset=True)

    # ---- e2e: host buffers through dwbc_batch_cycle_host ----
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    hsets = []
    for q, qd, fs, fl in sets[:2]:
        hsets.append((pin(q), pin(qd), pin(np.concatenate([f.reshape(-1) for f in fs])), pin(fl)))
    hout = (torch.empty((B, 33), dtype=torch.float64).pin_memory(), torch.empty((B, 33), dtype=torch.float64).pin_memory(),
            torch.empty((B, 33), dtype=torch.float64).pin_memory(), torch.empty((B,), dtype=torch.int32).pin_memory())

    def e2e_step(j):
        q, qd, f, fl = hsets[j % len(hsets)]
        rd.cycle_host(ALL_STAGES, q, qd, f, fl, out=hout)

    e2e_steps = max(3, min(args.steps, 10))
    for j in range(2):
        e2e_step(j)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    for j in range(e2e_steps):
        e2e_step(j)
    e1.record()
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0))
    clocks = sampler.stop() if rank == 0 else None

    # ---- max over ranks ----
    red = torch.tensor([ms_total, e2e_ms, float(np.percentile(per_step, 50)), float(np.percentile(per_step, 99))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, p50, p99 = [float(x) for x in red.tolist()]
    lsum = torch.tensor([float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(lsum, op=dist.ReduceOp.SUM)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    value = world * B * args.steps / (ms_total * 1e-3)
    e2e_val = world * B * e2e_steps / (e2e_ms * 1e-3)
    h2d_in, d2h_out = io_bytes_per_solve(TASK_DOFS)

    # ---- roofline of the dominant kernel (FP64 pipe) ----
    c_mean = 12.0 if args.contact_mix == "ds" else 9.0
    fl_ds = per_kernel_flops(TASK_DOFS, c=12, qp_iters_total=iters_mean)
    fl_ss = per_kernel_flops(TASK_DOFS, c=6, qp_iters_total=iters_mean)
    wds = {"ds": 1.0, "ss": 0.0}.get(args.contact_mix, 0.5)
    flk = {k: wds * fl_ds[k] + (1 - wds) * fl_ss[k] for k in fl_ds}
    # the one-kernel cycle does the work of the five dense kernels
    flk["lean_cycle_kernel"] = sum(flk[k] for k in ("rbd_kernel", "ainv_kernel", "contact_jacobian_kernel", "contact_dynamics_kernel", "task_space_kernel"))
    step_flops = sum(v for k, v in flk.items() if k != "lean_cycle_kernel")
    kern = {k: {"ms_per_launch": ms / max(args.steps, 1), "share": ms / max(sum(v[0] for v in ktimes.values()), 1e-12)}
            for k, (ms, cnt) in ktimes.items() if ms > 0}
    dom = max(kern, key=lambda k: kern[k]["ms_per_launch"])
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
    dom_s = kern[dom]["ms_per_launch"] * 1e-3
    achieved = B * flk[dom] / dom_s / 1e12
    for k in kern:  # every kernel's own fraction, so that a change of the dominant kernel between runs stays readable
        if k in flk and kern[k]["ms_per_launch"] > 0:
            kern[k]["achieved_tflops"] = B * flk[k] / (kern[k]["ms_per_launch"] * 1e-3) / 1e12
            kern[k]["frac"] = kern[k]["achieved_tflops"] / fp64_peak
    # DRAM traffic of the dominant kernel per launch from the committed `ncu --set full` capture (same B), else null
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        if tj.get("states_per_launch") == B and dom in tj["kernels"]:
            traffic = tj["kernels"][dom]["dram_read_bytes"] + tj["kernels"][dom]["dram_write_bytes"]
            traffic_src = "profiles/r01_traffic.json: " + tj["source"]
    except Exception:
        pass
    roofline = {"bound": "fp64", "kernel": dom, "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
                "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": "DFMA-chain probe in this run (dwbc_probe_fp64); MEASURED_PEAKS.json has no FP64 figure" if tf else "nominal",
                "algorithmic_flops_per_solve": step_flops, "kernel_flops_per_solve": flk[dom],
                "whole_step": {"achieved": world * B * step_flops / (ms_total / args.steps * 1e-3) / 1e12 / world, "frac": B * step_flops / (ms_total / args.steps * 1e-3) / 1e12 / fp64_peak},
                "hbm": {"io_bytes_per_solve": h2d_in + d2h_out, "achieved_gbs": B * (h2d_in + d2h_out) / (ms_total / args.steps * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                        "peak_source": hbm_src},
                "kernels": kern, "mean_qp_iterations_per_solve": iters_mean}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args), "states_per_gpu": B, "global_batch": world * B,
                       "l2": f"{N_INPUT_SETS} input sets rotated per step (> 126 MB L2); per-step workspace traffic is GBs",
                       "parallelism": f"independent states sharded over {world} GPU(s), no collective on the data path",
                       "solved_fraction": ok_frac, "outputs": args.outputs, "path": "lean_cycle_kernel + hqp_kernel (compacted)" if rd.last_path() == 1 else "generic kernels"},
            "latency_ms": {"p50": p50, "p99": p99, "note": "one batched cycle of B states per GPU, device-resident inputs"},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": B * h2d_in, "d2h_bytes_per_step": B * d2h_out, "steps": e2e_steps,
                    "ms_per_step": e2e_ms / e2e_steps, "api": "dwbc_batch_cycle_host (pinned host buffers)"},
            "gpu_launches": int(lsum.item()), "clocks": clocks, "roofline": roofline}
    if world == 1 and not args.no_cpu_baseline:
        q, qd, fs, fl = sets[0]
        line["cpu_baseline"] = cpu_baseline(args, q, qd, fs, fl)
    emit(line)
    if world > 1:
        dist.destroy_process_group()
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
