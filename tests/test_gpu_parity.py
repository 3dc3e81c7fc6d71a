"""Parity tests proper (`-m gpu`): the CUDA path, called through the C ABI (libdwbc_b200.so via
libdwbc_b200.RobotDataBatch), against (a) the reference's golden vectors (reference tests/cases/{1,2},
tests/dwbc_test.cpp:29-361) and (b) the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): Lambda matrices and closed-form torques 1e-9 relative; HQP solutions
and torques 1e-6 with identical active sets.
"""
import os

import numpy as np
import pytest

from tests.common import *

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ALL = 0xF


def _dev():
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: the gpu-marked tests need one (there is no CPU fallback)")
    return torch.device("cuda", 0)


def T(a, dt=None):
    dt = dt or torch.float64
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=_dev()).contiguous()


def run_product(config, q, qd, fstars, flags, stages, ncontacts=2):
    from libdwbc_b200 import RobotDataBatch

    B = q.shape[0]
    rd = configure_product(RobotDataBatch(URDF, B, 0), config, ncontacts)
    rd.update_kinematics(T(q), T(qd))
    rd.set_contact(T(flags, torch.uint8))
    rd.calc_contact_constraint()
    for l, f in enumerate(fstars):
        rd.set_task_space(l, T(f), T(posture_selector()) if (config == 2 and l == 2) else None)
    out = rd.solve(stages)
    torch.cuda.synchronize()
    return rd, tuple(o.cpu().numpy() for o in out)


def G(rd, name, level=0):
    return rd.get(name, level).cpu().numpy()


def test_library_is_cuda():
    """The loaded library is the nvcc-built one and it launches kernels (no emulation, no fallback)."""
    from libdwbc_b200._lib import library_path, load_library

    L = load_library()
    assert os.path.basename(library_path()) == "libdwbc_b200.so"
    n0 = L.dwbc_launch_count()
    q = np.array([Q_CASE[1]], dtype=np.float64)
    run_product("golden", q, np.zeros((1, 39)), [np.zeros((1, 6)), np.zeros((1, 3))], np.array([[1, 1, 0, 0]], np.uint8), ALL, 4)
    assert L.dwbc_launch_count() >= n0 + 6


@pytest.mark.parametrize("case", [1, 2])
def test_golden_cases(case):
    from oracle.urdf_model import read_golden

    q = np.array([Q_CASE[case]], dtype=np.float64)
    flags = np.array([[1, 1, 0, 0]], dtype=np.uint8)
    fs = [np.array([FSTAR_CASE[case][0]]), np.array([FSTAR_CASE[case][1]])]
    rd, (tg, tt, tc, st) = run_product("golden", q, np.zeros((1, 39)), fs, flags, ALL, 4)
    assert st[0] == 7
    g = lambda n: read_golden(os.path.join(GOLDEN, "cases", str(case), n))
    for name, gname in GOLDEN_NAMES:
        a = G(rd, name)[0]
        gg = g(gname)
        if a.ndim == 1:
            gg = gg[:, 0]
        else:
            a = a[:gg.shape[0], :gg.shape[1]]
        assert relerr(a, gg) < 1e-9, name
    if case == 1:
        V2, V2g = G(rd, "V2")[0], g("V2")
        assert np.abs(V2.T @ V2 - V2g.T @ V2g).max() < 1e-9
    assert relerr(tt[0], g("torque_task_")[:, 0]) < 1e-9
    assert relerr(tc[0], g("torque_contact_")[:, 0]) < (1e-9 if case == 1 else 1e-4)
    if case == 1:
        assert rd.active_set(0)[0] == [66, 67, 68, 69, 70, 71, 72, 73, 74, 75, 78]
        assert rd.active_set(1)[0] == [67, 68, 71, 72, 75, 78]
    else:
        assert rd.active_set(1)[0] == [68, 70]


def test_case3_frame_invariance():
    """reference tests/dwbc_test.cpp:262-361 (CASE 3): the whole robot yawed by 90 degrees, f* rotated with it,
    reproduces the same joint torques (checked on case 1, whose QP solution is unique)."""
    from oracle.urdf_model import read_golden

    q = np.array([Q_CASE[1]], dtype=np.float64)
    q[0, 5] = np.sin(np.pi / 4)
    q[0, 39] = np.cos(np.pi / 4)
    flags = np.array([[1, 1, 0, 0]], dtype=np.uint8)
    Rz = np.array([[0.0, -1, 0], [1, 0, 0], [0, 0, 1]])
    f0 = np.array(FSTAR_CASE[1][0])
    f1 = np.array(FSTAR_CASE[1][1])
    fs = [np.concatenate([Rz @ f0[:3], Rz @ f0[3:]])[None], (Rz @ f1)[None]]
    rd, (tg, tt, tc, st) = run_product("golden", q, np.zeros((1, 39)), fs, flags, ALL, 4)
    g = lambda n: read_golden(os.path.join(GOLDEN, "cases", "1", n))[:, 0]
    assert st[0] == 7
    assert relerr(tg[0], g("torque_grav_")) < 1e-9
    assert relerr(tt[0], g("torque_task_")) < 1e-8
    assert relerr(tc[0], g("torque_contact_")) < 1e-8


def test_config1_closed_form_vs_oracle():
    from libdwbc_b200 import SOLVE_GRAV, SOLVE_TASK
    from oracle import oracle as orc

    B = 64
    q, qd, rng = synth_states(B, 20241)
    fs = synth_fstar(rng, B, dofs=(3, 3), scales=(0.5, 1.0), hard_frac=0)
    flags = np.ones((B, 2), dtype=np.uint8)
    rd, (tg, tt, tc, st) = run_product(1, q, qd, fs, flags, SOLVE_GRAV | SOLVE_TASK)
    ref = configure_oracle(orc.OracleRobot(URDF), 1)
    mats = {k: G(rd, k) for k in ("A", "A_inv", "B", "G", "J_C", "Lambda_contact", "J_C_INV_T", "N_C", "W", "W_inv", "NwJw", "com_pos",
                                  "link_pos", "link_rot", "link_vel")}
    lv = {(k, l): G(rd, k, l) for k in ("Lambda_task", "J_kt", "J_task") for l in (0, 1)}
    nul = G(rd, "Null_task", 0)
    assert np.all(st & 3 == 3)
    for b in range(B):
        ref.update_kinematics(q[b], qd[b]); ref.set_contact(1, 1); ref.calc_contact_constraint()
        ref.set_task(0, fs[0][b]); ref.set_task(1, fs[1][b]); ref.calc_grav(); ref.calc_task_torque(False)
        assert relerr(tg[b], ref.get("torque_grav")[:, 0]) < 1e-9
        assert relerr(tt[b], ref.get("torque_task")[:, 0]) < 1e-9
        for k in ("A", "A_inv", "J_C", "Lambda_contact", "J_C_INV_T", "N_C", "W", "W_inv", "NwJw"):
            o = ref.get(k)
            assert relerr(mats[k][b][:o.shape[0], :o.shape[1]], o) < 1e-9, k
        for k in ("B", "G", "com_pos"):
            assert relerr(mats[k][b], ref.get(k)[:, 0]) < 1e-9, k
        for k in ("link_pos", "link_rot", "link_vel"):
            assert relerr(mats[k][b], ref.get(k)) < 1e-10, k
        for l in (0, 1):
            assert relerr(lv[("Lambda_task", l)][b], ref.get(f"Lambda_task{l}")) < 1e-9
            assert relerr(lv[("J_kt", l)][b], ref.get(f"J_kt{l}")) < 1e-9
            assert relerr(lv[("J_task", l)][b], ref.get(f"J_task{l}")) < 1e-12
        assert relerr(nul[b], ref.get("Null_task0")) < 1e-9


def test_gpu_walking_state_of_data_confirmation():
    """reference tests/sp_test/data_confirmation.cpp:41-84 (captured walking state, hierarchy 3 / 3 / 3 / 12, four registered contacts)"""
    from tests.common import check_walking_state, walking_inputs

    q, qd, fs, flags = walking_inputs(1)
    rd, out = run_product("walking", q, qd, fs, flags, ALL)
    check_walking_state(rd, out, np.asarray)


def test_gpu_perturbed_golden_set():
    """The reference's repeat test (tests/test.cpp:156-162): the golden configuration perturbed by 0.01 U per coordinate, base
    rotation reset to identity, through the golden hierarchy (pelvis 6D / upper-body rotation, torque limit 300, HQP +
    redistribution) -- every state of a 4096-state set against the oracle."""
    from oracle import oracle as orc

    B = 4096
    rng = np.random.default_rng(156162)
    q = np.array([Q_CASE[1]] * B, dtype=np.float64) + 0.01 * rng.uniform(-1, 1, (B, 40))
    q[:, 3:6] = 0.0
    q[:, 39] = 1.0
    qd = np.zeros((B, 39))
    fs = [np.array([FSTAR_CASE[1][0]] * B), np.array([FSTAR_CASE[1][1]] * B)]
    flags = np.array([[1, 1, 0, 0]] * B, dtype=np.uint8)
    rd, (tg, tt, tc, st) = run_product("golden", q, qd, fs, flags, ALL, 4)
    o = configure_oracle(orc.OracleRobot(URDF), "golden", 4).run_batch(q, qd, fs, flags, sets=True)
    assert np.array_equal(st & 7, o["status"] & 7) and ((st & 7) == 7).mean() > 0.99
    good = (st & 7) == 7
    scale = lambda a: np.maximum(1.0, np.abs(a).max(axis=1))
    eg = (np.abs(tg - o["torque_grav"]).max(axis=1) / scale(o["torque_grav"]))[good].max()
    et = (np.abs(tt - o["torque_task"]).max(axis=1) / scale(o["torque_task"]))[good].max()
    ec = (np.abs(tc - o["torque_contact"]).max(axis=1) / scale(o["torque_contact"]))[good].max()
    print(f"perturbed golden set: {int(good.sum())} of {B} solved, errors {eg:.1e} {et:.1e} {ec:.1e}")
    assert eg < 1e-9 and et < 1e-6 and ec < 1e-6, (eg, et, ec)
    act = active_masks(rd, 2)[:, :2]
    diff = (act ^ o["active"][:, :2])[good]
    n_hard = int(((diff & ~o["gray"][:, :2][good]) != 0).any(axis=(1, 2)).sum())
    assert n_hard <= 2, n_hard


def _mixed_flags(B, rng):
    """DS : left-SS : right-SS = 0.5 : 0.25 : 0.25."""
    u = rng.uniform(size=B)
    flags = np.ones((B, 2), dtype=np.uint8)
    flags[(u >= 0.5) & (u < 0.75)] = [1, 0]
    flags[u >= 0.75] = [0, 1]
    return flags


@pytest.mark.parametrize("mixed", [False, True])
def test_config2_hqp_vs_oracle(mixed):
    """3-level HQP + redistribution on 1024 seeded states: torques 1e-6, status and active sets identical."""
    from oracle import oracle as orc

    B = 1024
    q, qd, rng = synth_states(B, 20242 + (2 if mixed else 0))
    fs = synth_fstar(rng, B)
    flags = _mixed_flags(B, rng) if mixed else np.ones((B, 2), dtype=np.uint8)
    rd, (tg, tt, tc, st) = run_product(2, q, qd, fs, flags, ALL)
    ref = configure_oracle(orc.OracleRobot(URDF), 2)
    o = ref.run_batch(q, qd, fs, flags)
    assert np.array_equal(st & 7, o["status"] & 7)
    good = (o["status"] & 7) == 7
    assert good.mean() > 0.9
    scale = lambda a: np.maximum(1.0, np.abs(a).max(axis=1))
    assert (np.abs(tg - o["torque_grav"]).max(axis=1) / scale(o["torque_grav"]))[good].max() < 1e-9
    assert (np.abs(tt - o["torque_task"]).max(axis=1) / scale(o["torque_task"]))[good].max() < 1e-6
    assert (np.abs(tc - o["torque_contact"]).max(axis=1) / scale(o["torque_contact"]))[good].max() < 1e-6
    # active sets: re-run a subset state by state through the oracle's single-state API
    act = [rd.active_set(l) for l in range(3)] + [rd.active_set(-1)]
    idx = [b for b in range(B) if good[b]][:96]
    # include every state of the subset that actually has an active QP
    hard = [b for b in range(B) if good[b] and any(len(act[l][b]) for l in range(4))][:64]
    nact = 0
    for b in sorted(set(idx + hard)):
        ref.update_kinematics(q[b], qd[b]); ref.set_contact(*flags[b]); ref.calc_contact_constraint()
        for l in range(3):
            ref.set_task(l, fs[l][b], posture_selector() if l == 2 else None)
        ref.calc_grav(); ref.calc_task_torque(True); ref.calc_redistribute(True)
        for l in (0, 1, 2, -1):
            a_ref = ref.active_set(l)
            nact += len(a_ref)
            assert same_active_set(act[l if l >= 0 else 3][b], a_ref, gray_rows(ref, l)), (b, l)
    assert nact > 50  # the sample does exercise the active-set logic


def gray_rows(ref, level):
    """Rows whose slack at the oracle's solution lies within a decade of the reporting threshold 1e-9 (1 + |b|) (the rule
    both sides use to call a row active).  Such weakly active rows -- zero multiplier, slack inherited from the round-off
    of the level above -- are the only ones allowed to differ."""
    if level < 0:
        A = ref.get("redisA")
        if A.size == 0 or ref.get("cf_redis_qp").size == 0:  # single support: no internal-force space, no QP
            return set()
        ub, x = ref.get("redisub")[:, 0], ref.get("cf_redis_qp")[:, 0]
    else:
        A, ub = ref.get(f"qpA{level}"), ref.get(f"qpub{level}")[:, 0]
        x = np.concatenate([ref.get(f"f_star_qp{level}")[:, 0], ref.get(f"contact_qp{level}")[:, 0]])[:A.shape[1]]
    if A.size == 0:
        return set()
    s = (A @ x - ub) / (1.0 + np.abs(ub))
    return set(np.where((s > -1e-8) & (s < -1e-10))[0].tolist())


def same_active_set(a, b, gray):
    return set(a) - gray == set(b) - gray


def test_full_size_properties():
    """BASELINE size (B = 65536, config 2): size-independent properties instead of an oracle pass.
    (1) batch-permutation invariance: solving a permuted batch gives the permuted result bit for bit;
    (2) the returned torques satisfy every QP row: |tau_grav + tau_task + tau_contact| <= tlim (+1e-6);
    (3) a seeded sample of 256 states equals the oracle to 1e-6."""
    from oracle import oracle as orc

    B = 65536
    q, qd, rng = synth_states(B, 20242)
    fs = synth_fstar(rng, B)
    flags = np.ones((B, 2), dtype=np.uint8)
    rd, (tg, tt, tc, st) = run_product(2, q, qd, fs, flags, ALL)
    perm = np.random.default_rng(1).permutation(B)
    rd2, (tg2, tt2, tc2, st2) = run_product(2, q[perm], qd[perm], [f[perm] for f in fs], flags[perm], ALL)
    assert np.array_equal(st[perm], st2)
    assert np.array_equal(tg[perm], tg2) and np.array_equal(tt[perm], tt2) and np.array_equal(tc[perm], tc2)
    good = (st & 7) == 7
    assert good.mean() > 0.9
    tot = tg + tt + tc
    assert np.abs(tot[good]).max() <= 300.0 + 1e-6
    sample = np.random.default_rng(2).choice(B, 256, replace=False)
    ref = configure_oracle(orc.OracleRobot(URDF), 2)
    o = ref.run_batch(q[sample], qd[sample], [f[sample] for f in fs], flags[sample])
    assert np.array_equal(st[sample] & 7, o["status"] & 7)
    ok = (o["status"] & 7) == 7
    for a, k in ((tg, "torque_grav"), (tt, "torque_task"), (tc, "torque_contact")):
        err = np.abs(a[sample] - o[k]).max(axis=1) / np.maximum(1.0, np.abs(o[k]).max(axis=1))
        assert err[ok].max() < 1e-6, k


def test_cycle_host_matches_device_calls():
    """dwbc_batch_cycle_host (host buffers in, host buffers out) == the staged device-pointer calls."""
    from libdwbc_b200 import RobotDataBatch

    B = 257  # ragged: not a multiple of the warps per block
    q, qd, rng = synth_states(B, 7)
    fs = synth_fstar(rng, B)
    flags = np.ones((B, 2), dtype=np.uint8)
    rd, (tg, tt, tc, st) = run_product(2, q, qd, fs, flags, ALL)
    from libdwbc_b200 import OUT_ALL
    rd2 = configure_product(RobotDataBatch(URDF, B, 0), 2)
    rd2.set_outputs(OUT_ALL)  # same kernels as the staged calls above (the default of cycle_host is the one-kernel cycle, tested below)
    rd2.set_task_space(2, T(np.zeros((B, 21))), T(posture_selector()))
    tg2, tt2, tc2, st2 = rd2.cycle_host(ALL, q, qd, fs, flags)
    assert np.array_equal(tg, tg2) and np.array_equal(tt, tt2) and np.array_equal(tc, tc2) and np.array_equal(st, st2)


def test_gpu_sliced_cycle_host():
    """The pipelined (sliced, two-stream) dwbc_batch_cycle_host == the unsliced call == the staged device-pointer calls, bit for bit,
    for ragged slices, mixed support and pinned host buffers; automatic slicing at a batch large enough to trigger it."""
    from libdwbc_b200 import RobotDataBatch
    from tests.test_emul_parity import run_sliced_cycle_host

    make = lambda B: RobotDataBatch(URDF, B, 0)
    (q, qd, fs, flags), out = run_sliced_cycle_host(make, T, 1031, ((1, False), (2, False), (7, True), (64, False)))
    rd, staged = run_product(2, q, qd, fs, flags, ALL)
    for a, b in zip(out, staged):
        assert np.array_equal(a, b)
    # automatic slicing (B >= 8192 -> 2 slices) with pinned buffers, against the unsliced call
    B = 8192 + 40
    q, qd, rng = synth_states(B, 12)
    fs = synth_fstar(rng, B)
    flags = np.ones((B, 2), dtype=np.uint8)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    args = (pin(q), pin(qd), pin(np.concatenate([f.reshape(-1) for f in fs])), pin(flags))
    res = []
    for slices in (0, 1):
        rd = configure_product(make(B), 2)
        rd.set_task_space(2, T(np.zeros((B, 21))), T(posture_selector()))
        rd.set_host_slices(slices)
        res.append([np.array(a) for a in rd.cycle_host(ALL, *args)])
    for a, b in zip(*res):
        assert np.array_equal(a, b)


def test_error_behaviour():
    """Call-order and argument errors are reported, not silently ignored (reference prints and returns:
    src/dwbc.cpp:282-296, 407-411, 664-671)."""
    from libdwbc_b200 import CONTACT_6D, DwbcError, RobotDataBatch

    rd = RobotDataBatch(URDF, 4, 0)
    with pytest.raises(DwbcError):
        rd.calc_contact_constraint()
    with pytest.raises(DwbcError):
        rd.update_kinematics(T(np.zeros((4, 39))), T(np.zeros((4, 39))))
    with pytest.raises(DwbcError):
        rd.add_task_space(0, 0, "no_such_link")
    rd.add_contact_constraint(6, CONTACT_6D, FOOT_POINT)
    with pytest.raises(DwbcError):
        rd.add_contact_constraint(6, CONTACT_6D, FOOT_POINT)  # "Contact Constraint Already Exist"


def test_no_contact_and_empty_flags():
    """A state with no active contact has no contact-consistent dynamics: status 0, zero torques (edge case)."""
    B = 8
    q, qd, rng = synth_states(B, 3)
    fs = synth_fstar(rng, B)
    flags = np.ones((B, 2), dtype=np.uint8)
    flags[3] = [0, 0]
    rd, (tg, tt, tc, st) = run_product(2, q, qd, fs, flags, ALL)
    assert st[3] & 7 == 0
    assert np.all(tt[3] == 0) and np.all(tc[3] == 0)
    assert np.all((st[[0, 1, 2, 4, 5, 6, 7]] & 1) == 1)


def test_analytic_redistribution_contact_force_zmp():
    """CalcTaskControlTorque(hqp=false) + CalcContactRedistribute(hqp=false) (reference src/dwbc.cpp:1570-1619,
    src/wbd.cpp:273-404: the path the reference's Python bridge runs), getContactForce and getZMP (src/dwbc.cpp:891-939)
    against the oracle, 1e-9 relative."""
    from libdwbc_b200 import SOLVE_GRAV, SOLVE_REDIST_ANALYTIC, SOLVE_TASK
    from oracle import oracle as orc

    B = 512
    q, qd, rng = synth_states(B, 20251)
    q[:, 3:6] *= 3.0  # base yaw of either sign: both branches of Eigen's eulerAngles
    q[:, 39] = np.sqrt(1.0 - np.sum(q[:, 3:6] ** 2, axis=1))
    fs = synth_fstar(rng, B, dofs=config_dofs(1), scales=(0.5, 1.0))
    flags = np.ones((B, 2), dtype=np.uint8)
    flags[5] = [1, 0]
    rd, (tg, tt, tc, st) = run_product(1, q, qd, fs, flags, SOLVE_GRAV | SOLVE_TASK | SOLVE_REDIST_ANALYTIC)
    ref = configure_oracle(orc.OracleRobot(URDF), 1)
    o = ref.run_batch(q, qd, fs, flags, stages=1 | 2 | 16)
    assert np.array_equal(st & 7, o["status"] & 7)
    for name, a in (("torque_grav", tg), ("torque_task", tt), ("torque_contact", tc)):
        err = np.linalg.norm(a - o[name], axis=1) / np.maximum(np.linalg.norm(o[name], axis=1), 1e-3)
        assert err.max() < 1e-9, (name, err.max(), int(err.argmax()))
    cf = rd.get_contact_force()
    assert relerr(cf.cpu().numpy(), o["contact_force"]) < 1e-9
    zmp = rd.get_zmp(cf)
    assert relerr(zmp.cpu().numpy(), o["zmp"]) < 1e-9


@pytest.mark.parametrize("flags", [(1, 1), (1, 0)])
def test_reduced_dimension_variant_vs_oracle(flags):
    """a9 (reference src/dwbc.cpp:2752-3446): reduced dynamics, contact constraint, gravity compensation, task space and
    closed-form task torque against the oracle's literal restatement (1e-9) and against the full-model torques
    (the reference's own check, tests/sp_test/redu_dyn_test.cpp:323-360)."""
    from tests.test_emul_reduced import check_reduced_against_oracle, reduced_product

    B = 8
    q, qd, rng = synth_states(B, 20243)
    fs = [0.5 * rng.standard_normal((B, 3)), rng.standard_normal((B, 3)), rng.standard_normal((B, 3)), rng.standard_normal((B, 6))]
    rd = reduced_product(None, q, qd, flags, fs, B)
    rd.update_kinematics(T(q), T(qd))
    rd.set_contact(T(np.tile(np.array(flags, dtype=np.uint8), (B, 1)), torch.uint8))
    rd.reduced_dynamics_calculate(flags)
    rd.reduced_calc_contact_constraint()
    keep = [T(f) for f in fs]
    for l, f in enumerate(keep):
        rd.set_task_space(l, f)
    check_reduced_against_oracle(rd, lambda x: x.cpu().numpy(), q, qd, flags, fs)


def test_reduced_full_size_equals_full_model():
    """B = 65536 (config 3(i)): the reduced pipeline reproduces the full-model torque_grav_ / torque_task_ of every state."""
    from libdwbc_b200 import SOLVE_GRAV, SOLVE_TASK
    from tests.test_emul_reduced import reduced_product

    B = 65536
    q, qd, rng = synth_states(B, 20243)
    fs = [0.5 * rng.standard_normal((B, 3)), rng.standard_normal((B, 3)), rng.standard_normal((B, 3)), rng.standard_normal((B, 6))]
    flags = T(np.ones((B, 2), dtype=np.uint8), torch.uint8)
    qt, qdt, fst = T(q), T(qd), [T(f) for f in fs]
    full = reduced_product(None, q, qd, (1, 1), fs, B)
    full.update_kinematics(qt, qdt)
    full.set_contact(flags)
    full.calc_contact_constraint()
    for l, f in enumerate(fst):
        full.set_task_space(l, f)
    tg, tt, _, st = full.solve(SOLVE_GRAV | SOLVE_TASK)
    red = reduced_product(None, q, qd, (1, 1), fs, B)
    red.update_kinematics(qt, qdt)
    red.set_contact(flags)
    red.reduced_dynamics_calculate((1, 1))
    red.reduced_calc_contact_constraint()
    for l, f in enumerate(fst):
        red.set_task_space(l, f)
    rg, rt, _, rs = red.reduced_solve(SOLVE_GRAV | SOLVE_TASK)
    torch.cuda.synchronize()
    assert bool((st == 3).all()) and bool((rs == 3).all())
    eg = ((rg - tg).norm(dim=1) / tg.norm(dim=1)).max().item()
    et = ((rt - tt).norm(dim=1) / tt.norm(dim=1).clamp_min(1.0)).max().item()
    assert eg < 1e-8 and et < 1e-7, (eg, et)


def _lqp_product(q, qd, flags, fs):
    from libdwbc_b200 import RobotDataBatch

    B = q.shape[0]
    rd = configure_product(RobotDataBatch(URDF, B, 0), "lqp")
    rd.update_kinematics(T(q), T(qd))
    rd.set_contact(T(flags, torch.uint8))
    rd.calc_contact_constraint()
    for l, f in enumerate(fs):
        rd.set_task_space(l, T(f))
    out = rd.calc_control_torque_lqp()
    torch.cuda.synchronize()
    return rd, out


class _LqpView:
    """numpy view of the device results with the accessors tests/test_emul_lqp.check_lqp uses"""

    def __init__(self, rd):
        self.rd = rd

    def get_lqp(self, name, level=0):
        return self.rd.get_lqp(name, level).cpu().numpy()


@pytest.mark.parametrize("hard", [False, True])
def test_gpu_lqp_vs_oracle(hard):
    """a10 (config 3(ii) tasks, reference tests/sp_test/jacc_compare.cpp:319-334): every level's y, the level-1 slack and the
    torque equal the oracle's exact solve to 1e-6, double and single support.  Parity w.r.t. the reference is unpinned (OSQP)."""
    from tests.test_emul_lqp import check_lqp, lqp_oracle_states

    B = 24
    q, qd, rng = synth_states(B, 20243)
    fs = lqp_fstars(rng, B, hard)
    flags = np.ones((B, 2), dtype=np.uint8)
    flags[2::4] = [1, 0]
    flags[3::4] = [0, 1]
    rd, out = _lqp_product(q, qd, flags, fs)
    refs = lqp_oracle_states(q, qd, flags, fs)
    check_lqp(_LqpView(rd), tuple(o.cpu().numpy() for o in out), refs, flags)


def test_gpu_lqp_full_size_properties():
    """B = 65536 (config 3(ii)): size-independent properties of the LQP answer.
    (1) level 0 (floating-base rows of the equation of motion) holds for the final answer of every solved state;
    (2) hierarchy: the contact-row residual Jc qdd of the final answer equals the one level 1 left (later levels work in its
        null space);
    (3) torque = A[6:,:] qdd + Jc^T[6:,:] Fc + B[6:] and |torque| <= 200 (+1e-6), |qdd_joint| <= 5 + v1 (+1e-6);
    (4) a seeded sample of 64 states equals the oracle to 1e-6."""
    from tests.test_emul_lqp import check_lqp, lqp_oracle_states

    B = 65536
    n = 39
    q, qd, rng = synth_states(B, 20243)
    fs = lqp_fstars(rng, B, False)
    hard = rng.choice(B, B // 20, replace=False)
    fs[0][hard, 1] = 12.0
    flags = np.ones((B, 2), dtype=np.uint8)
    rd, (y, tau, st) = _lqp_product(q, qd, flags, fs)
    ok = st == 4
    assert ok.float().mean().item() > 0.98
    A, Bn, Jc = rd.get("A"), rd.get("B"), rd.get("J_C")
    qdd, Fc = y[:, :n], y[:, n:]
    eom = torch.einsum("bij,bj->bi", A, qdd) + torch.einsum("bji,bj->bi", Jc, Fc) + Bn
    scale = Bn.norm(dim=1).clamp_min(1.0)
    assert (eom[:, :6].norm(dim=1) / scale)[ok].max().item() < 1e-8
    assert ((eom[:, 6:] - tau).norm(dim=1) / scale)[ok].max().item() < 1e-10
    assert tau[ok].abs().max().item() <= 200.0 + 1e-6
    y1 = rd.get_lqp("y", 1)
    r1 = torch.einsum("bij,bj->bi", Jc, y1[:, :n])
    rf = torch.einsum("bij,bj->bi", Jc, qdd)
    assert ((rf - r1).norm(dim=1))[ok].max().item() < 1e-7
    v1 = rd.get_lqp("v1")
    m = 33
    va = torch.maximum(v1[:, 20:20 + m], v1[:, 20 + m:20 + 2 * m])
    assert ((qdd[:, 6:].abs() - 5.0 - va)[ok]).max().item() < 1e-6
    sample = np.random.default_rng(3).choice(B, 64, replace=False)
    refs = lqp_oracle_states(q[sample], qd[sample], flags[sample], [f[sample] for f in fs])

    class V:
        def get_lqp(self, name, level=0):
            return rd.get_lqp(name, level)[torch.as_tensor(sample, device=y.device)].cpu().numpy()

    idx = torch.as_tensor(sample, device=y.device)
    check_lqp(V(), (y[idx].cpu().numpy(), tau[idx].cpu().numpy(), st[idx].cpu().numpy()), refs, flags[sample])


def test_gpu_fstar_and_wire_format():
    """SURVEY.md section 8(f) rows 1 and 3 on the device: f* from quintic / slerp trajectories for every LINK_MODE family vs the
    numpy restatement (oracle/fstar.py), MuJoCo qpos/qvel layout, float32 torque command."""
    from libdwbc_b200 import RobotDataBatch
    from tests.test_fstar import run_fstar_case

    def D(a):
        a = np.ascontiguousarray(a)
        return T(a, torch.uint8 if a.dtype == np.uint8 else torch.float64)

    run_fstar_case(lambda B: RobotDataBatch(URDF, B, 0), D, lambda x: x.cpu().numpy())


@pytest.mark.parametrize("dof", [18, 25, 45])
def test_gpu_generic_dof(dof):
    """the reference's DoF-study models (18 / 25 / 45 joints) on the device vs the oracle."""
    from libdwbc_b200 import RobotDataBatch
    from tests.test_generic_dof import run_generic

    def D(a):
        a = np.ascontiguousarray(a)
        return T(a, torch.uint8 if a.dtype == np.uint8 else torch.float64)

    run_generic(dof, lambda urdf, B: RobotDataBatch(urdf, B, 0), D, lambda x: x.cpu().numpy(), B=8)


@pytest.mark.parametrize("name", ["wo_head", "fixed_head", "ub"])
def test_gpu_model_variant(name):
    """the reference's model variants (tests/dyros_tocabi_{wo_head,fixed_head,ub}.urdf) on the device
    vs the oracle: the whole path on the 32-joint models, UpdateKinematics on the leg-less upper-body model."""
    from libdwbc_b200 import RobotDataBatch
    from tests.test_generic_dof import run_upper_body, run_variant

    def D(a):
        a = np.ascontiguousarray(a)
        return T(a, torch.uint8 if a.dtype == np.uint8 else torch.float64)

    make = lambda urdf, B: RobotDataBatch(urdf, B, 0)
    H = lambda x: x.cpu().numpy()
    if name == "ub":
        run_upper_body(make, D, H, B=8)
    else:
        run_variant(name, make, D, H, B=8)


@pytest.mark.parametrize("limited", [False, True])
def test_gpu_reduced_hqp_vs_oracle(limited):
    """a9 QP branches (CalcSingleTaskTorqueWithQP_R, CalcContactRedistributeR; reference src/dwbc.cpp:3448-3599, 4777-4925) on the
    device vs the oracle, whose restatement is pinned by "reduced HQP == full-model HQP" (tests/test_oracle_reduced.py)."""
    from libdwbc_b200 import RobotDataBatch
    from tests.test_emul_reduced import TLIM_R, run_reduced_hqp

    def D(a):
        a = np.ascontiguousarray(a)
        return T(a, torch.uint8 if a.dtype == np.uint8 else torch.float64)

    run_reduced_hqp(lambda B: RobotDataBatch(URDF, B, 0), D, lambda x: x.cpu().numpy(), B=16, tlimR=TLIM_R if limited else None)


def test_gpu_reduced_hqp_full_size_equals_full_model():
    """B = 65536: without a torque limit the reduced HQP + redistribution reproduces the full-model HQP of every state
    (COM position / pelvis rotation hierarchy; 5 % of the states saturate the friction rows)."""
    from libdwbc_b200 import (CONTACT_6D, SOLVE_GRAV, SOLVE_HQP, SOLVE_REDIST, TASK_LINK_POSITION, TASK_LINK_ROTATION, RobotDataBatch)

    B = 65536
    q, qd, rng = synth_states(B, 20247)
    fs = [0.5 * rng.standard_normal((B, 3)), rng.standard_normal((B, 3))]
    fs[0][rng.choice(B, B // 20, replace=False), 1] = 4.0
    qt, qdt, fst, flags = T(q), T(qd), [T(f) for f in fs], T(np.ones((B, 2), dtype=np.uint8), torch.uint8)
    res = []
    for reduced in (False, True):
        rd = RobotDataBatch(URDF, B, 0)
        rd.add_contact_constraint(LEFT_FOOT, CONTACT_6D, FOOT_POINT, (0, 0, 1), 0.15, 0.075)
        rd.add_contact_constraint(RIGHT_FOOT, CONTACT_6D, FOOT_POINT, (0, 0, 1), 0.15, 0.075)
        rd.add_task_space(0, TASK_LINK_POSITION, "COM")
        rd.add_task_space(1, TASK_LINK_ROTATION, "pelvis_link")
        rd.update_kinematics(qt, qdt)
        rd.set_contact(flags)
        if reduced:
            rd.reduced_dynamics_calculate((1, 1))
            rd.reduced_calc_contact_constraint()
        else:
            rd.calc_contact_constraint()
        for l, f in enumerate(fst):
            rd.set_task_space(l, f)
        res.append((rd.reduced_solve if reduced else rd.solve)(SOLVE_GRAV | SOLVE_HQP | SOLVE_REDIST))
    torch.cuda.synchronize()
    (tg, tt, tc, st), (rg, rt, rc_, rs) = res
    both = ((st & 7) == 7) & ((rs & 7) == 7)
    assert both.float().mean().item() > 0.98
    et = ((rt - tt).norm(dim=1) / tt.norm(dim=1).clamp_min(1.0))[both].max().item()
    ec = ((rc_[:, :12] - tc[:, :12]).norm(dim=1) / tc.norm(dim=1).clamp_min(1.0))[both].max().item()
    assert et < 1e-6 and ec < 1e-5, (et, ec)


def test_gpu_reduced_lqp_vs_oracle():
    """a10 on the reduced model (ConfigureLQP_R + CalcControlTorqueLQP_R, reference src/dwbc.cpp:4504-4632) vs the oracle."""
    from libdwbc_b200 import RobotDataBatch
    from tests.test_emul_lqp import run_reduced_lqp

    def D(a):
        a = np.ascontiguousarray(a)
        return T(a, torch.uint8 if a.dtype == np.uint8 else torch.float64)

    run_reduced_lqp(lambda B: RobotDataBatch(URDF, B, 0), D, lambda x: x.cpu().numpy(), B=12)


def test_gpu_reduced_hqp_with_non_contact_levels():
    """a9: reduced HQP on the full hierarchy of reference tests/sp_test/redu_dyn_test.cpp:97-101 incl. the base-wrench QP of the
    non-contact levels (CalcSingleTaskTorqueWithQP_R_NC, src/dwbc.cpp:3601-3756) vs the oracle; at least one state has it active."""
    from libdwbc_b200 import RobotDataBatch
    from tests.test_emul_reduced import run_reduced_hqp_nc

    def D(a):
        a = np.ascontiguousarray(a)
        return T(a, torch.uint8 if a.dtype == np.uint8 else torch.float64)

    n_active = run_reduced_hqp_nc(lambda B: RobotDataBatch(URDF, B, 0), D, lambda x: x.cpu().numpy(), scales=(0.0, 0.0, 6.5, 6.8, 7.0, 7.2, 7.5, 1.0))
    assert n_active >= 1


def test_gpu_tight_torque_limit():
    """torque-limit rows active in most states (limit 45 Nm): GPU HQP + redistribution vs the oracle"""
    from libdwbc_b200 import RobotDataBatch
    from tests.test_emul_parity import run_tight_torque_limit

    def D(a):
        a = np.ascontiguousarray(a)
        return T(a, torch.uint8 if a.dtype == np.uint8 else torch.float64)

    run_tight_torque_limit(lambda B: RobotDataBatch(URDF, B, 0), D, lambda x: x.cpu().numpy(), 96)


def test_gpu_rl_bridge_batch():
    """RlWBCBridgeBatch / RlWBCBridge on the device (reference src/pybind/rl_wbc_bridge.cpp call sequence): MuJoCo-ordered input,
    float32 torque command, per-environment contact flags; checked against the oracle running the calls the reference bridge makes."""
    from libdwbc_b200 import RlWBCBridge, RlWBCBridgeBatch
    from tests.test_bridge import _oracle_cmd

    B = 16
    q, qd, rng = synth_states(B, 20252)
    qpos = np.concatenate([q[:, 0:3], q[:, 39:40], q[:, 3:6], q[:, 6:39]], axis=1)
    f0, f1 = 0.3 * rng.standard_normal((B, 6)), 0.3 * rng.standard_normal((B, 3))
    br = RlWBCBridgeBatch(B, device=0, urdf_path=URDF)
    br.UpdateKinematics(T(qpos), T(qd), None)
    br.SetContact(True, True)
    br.SetTaskSpace(0, T(f0))
    br.SetTaskSpace(1, T(f1))
    br.CalcTorque()
    cmd = br.getTorqueCommand()
    assert cmd.dtype == torch.float32 and tuple(cmd.shape) == (B, 33) and cmd.is_cuda
    cmd = cmd.cpu().numpy()
    for b in range(0, B, 3):
        want = _oracle_cmd(q[b], qd[b], (1, 1), f0[b], f1[b])
        assert np.abs(cmd[b] - want).max() <= 1e-5 * max(1.0, np.abs(want).max()), b
    one = RlWBCBridge(0, device=0, urdf_path=URDF)
    one.UpdateKinematics(qpos[1], qd[1], np.zeros(39))
    one.SetContact(True, True)
    one.SetTaskSpace(0, f0[1])
    one.SetTaskSpace(1, f1[1])
    one.CalcTorque()
    out = one.getTorqueCommand()
    assert len(out) == 33 and np.array_equal(np.float32(out), cmd[1])


# ---------------------------------------------------------------------------------------------------------------------
# DWBC_OUT_TORQUES: the one-kernel cycle (lean_cycle_kernel) + the active-set kernel over the compacted QP-active states
# ---------------------------------------------------------------------------------------------------------------------
def run_lean(config, q, qd, fstars, flags, stages, ncontacts=2, outputs=None):
    from libdwbc_b200 import OUT_TORQUES, RobotDataBatch

    B = q.shape[0]
    rd = configure_product(RobotDataBatch(URDF, B, 0), config, ncontacts)
    rd.set_outputs(OUT_TORQUES if outputs is None else outputs)
    rd.update_kinematics(T(q), T(qd))
    rd.set_contact(T(flags, torch.uint8))
    rd.calc_contact_constraint()
    for l, f in enumerate(fstars):
        rd.set_task_space(l, T(f), T(posture_selector()) if (config == 2 and l == 2) else None)
    out = rd.solve(stages)
    torch.cuda.synchronize()
    return rd, tuple(o.cpu().numpy() for o in out)


def active_masks(rd, nlevels):
    """[B, nlevels + 1, 4] uint32 bit masks of the product's reported active rows (last = redistribution)"""
    B = rd.batch
    out = np.zeros((B, nlevels + 1, 4), dtype=np.uint32)
    for k, l in enumerate(list(range(nlevels)) + [-1]):
        buf = torch.zeros((B, 4), dtype=torch.int32, device=_dev())
        rd._ck(rd.L.dwbc_batch_get_active_set(rd._h, l, rd.be.ptr(buf), rd.be.stream()))
        torch.cuda.synchronize()
        out[:, k, :] = buf.cpu().numpy().view(np.uint32)
    return out


@pytest.mark.parametrize("case", [1, 2])
def test_lean_golden(case):
    from oracle.urdf_model import read_golden

    q = np.array([Q_CASE[case]], dtype=np.float64)
    flags = np.array([[1, 1, 0, 0]], dtype=np.uint8)
    fs = [np.array([FSTAR_CASE[case][0]]), np.array([FSTAR_CASE[case][1]])]
    rd, (tg, tt, tc, st) = run_lean("golden", q, np.zeros((1, 39)), fs, flags, ALL, 4)
    assert rd.last_path() == 1 and st[0] == 7
    g = lambda n: read_golden(os.path.join(GOLDEN, "cases", str(case), n))
    assert relerr(tg[0], g("torque_grav_")[:, 0]) < 1e-9
    assert relerr(tt[0], g("torque_task_")[:, 0]) < 1e-9
    assert relerr(tc[0], g("torque_contact_")[:, 0]) < (1e-9 if case == 1 else 1e-4)
    if case == 1:
        assert rd.active_set(0)[0] == [66, 67, 68, 69, 70, 71, 72, 73, 74, 75, 78]
        assert rd.active_set(1)[0] == [67, 68, 71, 72, 75, 78]


def _full_parity(B, seed, mixed, outputs):
    """Every state of a full-size batch against the oracle: status identical, torque_grav_ 1e-9, torque_task_ / torque_contact_ 1e-6,
    active sets of all four QPs identical outside the oracle's weakly-active (gray) rows; the gray exemptions are counted."""
    from oracle import oracle as orc

    q, qd, rng = synth_states(B, seed)
    fs = synth_fstar(rng, B)
    flags = _mixed_flags(B, rng) if mixed else np.ones((B, 2), dtype=np.uint8)
    rd, (tg, tt, tc, st) = run_lean(2, q, qd, fs, flags, ALL, outputs=outputs)
    o = configure_oracle(orc.OracleRobot(URDF), 2).run_batch(q, qd, fs, flags, sets=True)
    bad = np.where((st & 7) != (o["status"] & 7))[0]
    if len(bad):  # leave the evidence where the builder can pick it up (gpurun_out/ travels back from the GPU box)
        import json
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"status_mismatch_B{B}_{'mixed' if mixed else 'ds'}_out{outputs}.json"), "w") as f:
            json.dump({"seed": seed, "B": B, "mixed": mixed, "index": bad.tolist(), "gpu": st[bad].tolist(), "oracle": o["status"][bad].tolist(),
                       "iters": rd.qp_iterations().cpu().numpy()[bad].tolist(), "n_gpu_not7": int(((st & 7) != 7).sum()),
                       "n_oracle_not7": int(((o["status"] & 7) != 7).sum())}, f)
    assert np.array_equal(st & 7, o["status"] & 7), (len(bad), bad[:10], st[bad][:10], o["status"][bad][:10])
    good = (o["status"] & 7) == 7
    assert good.mean() > 0.99
    scale = lambda a: np.maximum(1.0, np.abs(a).max(axis=1))
    eg = (np.abs(tg - o["torque_grav"]).max(axis=1) / scale(o["torque_grav"]))[good].max()
    et = (np.abs(tt - o["torque_task"]).max(axis=1) / scale(o["torque_task"]))[good].max()
    ec = (np.abs(tc - o["torque_contact"]).max(axis=1) / scale(o["torque_contact"]))[good].max()
    assert eg < 1e-9 and et < 1e-6 and ec < 1e-6, (eg, et, ec)
    act = active_masks(rd, 3)
    diff = (act ^ o["active"])[good]
    exempt = diff & o["gray"][good]
    hard = diff & ~o["gray"][good]
    n_hard = int((hard != 0).any(axis=(1, 2)).sum())
    n_exempt = int((exempt != 0).any(axis=(1, 2)).sum())
    n_active = int((o["active"][good] != 0).any(axis=(1, 2)).sum())
    print(f"B={B} mixed={mixed}: {int(good.sum())} solved, {n_active} with an active QP row, torque errors {eg:.1e} {et:.1e} {ec:.1e}, "
          f"{n_exempt} states used a gray-row exemption, {n_hard} differ outside gray rows, slow flag on {int(((st & 16) != 0).sum())}")
    if n_hard or n_exempt >= 1e-3 * B:
        import json
        gi = np.where(good)[0]
        hs = gi[np.where((hard != 0).any(axis=(1, 2)))[0]][:64]
        es = gi[np.where((exempt != 0).any(axis=(1, 2)))[0]][:64]
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"active_mismatch_B{B}_{'mixed' if mixed else 'ds'}_out{outputs}.json"), "w") as f:
            json.dump({"seed": seed, "B": B, "mixed": mixed, "n_hard": n_hard, "n_exempt": n_exempt,
                       "hard": [{"i": int(i), "gpu": act[i].tolist(), "oracle": o["active"][i].tolist(), "gray": o["gray"][i].tolist()} for i in hs],
                       "exempt": [{"i": int(i), "gpu": act[i].tolist(), "oracle": o["active"][i].tolist(), "gray": o["gray"][i].tolist()} for i in es]}, f)
    # Identical active sets, with two documented kinds of exemption, both counted and bounded:
    #  * gray rows: the oracle's own slack is within a decade of the 1e-9 (1 + |b|) reporting threshold (weakly active, zero multiplier);
    #  * degenerate vertices: more rows meet at the optimum than there are variables (e.g. 7 contact rows, 6 contact variables); which of
    #    the redundant rows shows a slack below 1e-9 is decided by rounding on both sides.  Seen on 3 of 65 536 (DS) and 4 of 131 072
    #    (mixed) states; the solutions themselves agree (torques above).
    assert n_hard <= 1e-4 * B, n_hard
    assert n_exempt < 1e-3 * B
    assert n_active > 0.2 * B  # the batch does exercise the active-set logic
    return rd


def test_lean_full_size_every_state_ds():
    """BASELINE.json configs[2] at its full size (65 536 states, double support), every state against the oracle."""
    from libdwbc_b200 import OUT_TORQUES

    rd = _full_parity(65536, 20242, False, OUT_TORQUES)
    assert rd.last_path() == 1


def test_lean_full_size_every_state_mixed():
    """One GPU's slice of BASELINE.json configs[4] (131 072 states, DS / left-SS / right-SS 0.5 / 0.25 / 0.25), every state."""
    from libdwbc_b200 import OUT_TORQUES

    rd = _full_parity(131072, 20244, True, OUT_TORQUES)
    assert rd.last_path() == 1


def test_generic_full_size_every_state_ds():
    """The same check for the generic kernels (DWBC_OUT_ALL), 65 536 states."""
    from libdwbc_b200 import OUT_ALL

    rd = _full_parity(65536, 20242, False, OUT_ALL)
    assert rd.last_path() == 0


def test_lean_equals_generic_bitwise_status_and_close_torques():
    """Output groups both ways on the same batch: identical status and active sets, torques to 1e-9."""
    from libdwbc_b200 import OUT_ALL

    B = 4096
    q, qd, rng = synth_states(B, 20247)
    fs = synth_fstar(rng, B)
    flags = _mixed_flags(B, rng)
    rl, (tg, tt, tc, st) = run_lean(2, q, qd, fs, flags, ALL)
    rg, (tg0, tt0, tc0, st0) = run_lean(2, q, qd, fs, flags, ALL, outputs=OUT_ALL)
    assert rl.last_path() == 1 and rg.last_path() == 0
    assert np.array_equal(st & 7, st0 & 7)
    for a, b in ((tg, tg0), (tt, tt0), (tc, tc0)):
        assert (np.abs(a - b).max(axis=1) / np.maximum(1.0, np.abs(b).max(axis=1))).max() < 1e-8
    assert (active_masks(rl, 3) != active_masks(rg, 3)).any(axis=(1, 2)).mean() < 1e-3
    with pytest.raises(Exception):
        rl.get("A_inv")


def test_lean_cycle_host_and_config1():
    from libdwbc_b200 import RobotDataBatch

    B = 1000
    q, qd, rng = synth_states(B, 20248)
    fs = synth_fstar(rng, B)
    flags = _mixed_flags(B, rng)
    _, ref = run_lean(2, q, qd, fs, flags, ALL)
    rd = configure_product(RobotDataBatch(URDF, B, 0), 2)
    rd.set_task_space(2, T(np.zeros((B, 21))), T(posture_selector()))
    rd.set_host_slices(3)
    out = rd.cycle_host(ALL, q, qd, fs, flags)
    assert rd.last_path() == 1
    for a, b in zip(out, ref):
        assert np.array_equal(a, b)
    # closed-form hierarchy of config 1 (hqp = false)
    fs1 = synth_fstar(rng, B, dofs=(3, 3), scales=(0.5, 1.0), hard_frac=0)
    rl, (tg, tt, tc, st) = run_lean(1, q, qd, fs1, flags, 3)
    rg, (tg0, tt0, tc0, st0) = run_product(1, q, qd, fs1, flags, 3)
    assert rl.last_path() == 1 and np.array_equal(st & 7, st0 & 7)
    assert relerr(tg, tg0) < 1e-10 and relerr(tt, tt0) < 1e-9


def test_cpp_hosts_on_the_gpu():
    """The two C++ hosts of the C ABI against the REAL library on the GPU: include/dwbc_b200.hpp (DWBC::RobotDataBatch, tests/shim_check.cpp)
    and include/dwbc_compat/dwbc.h (the header-compatible single-state DWBC::RobotData running the reference's own CASE 1 / CASE 2 sequence
    against the golden binaries, tests/robotdata_check.cpp)."""
    from libdwbc_b200._lib import library_path
    from oracle.urdf_model import read_golden
    from tests.test_abi import build_and_run_cpp

    _dev()
    lib = library_path()
    r = build_and_run_cpp(lib, "robotdata_check.cpp", ["include/dwbc_compat"], [URDF, os.path.join(GOLDEN, "cases")])
    assert r.returncode == 0 and r.stdout.strip().endswith("ALL OK"), r.stdout + r.stderr
    r = build_and_run_cpp(lib, "shim_check.cpp", ["include"], [URDF])
    assert r.returncode == 0, r.stdout + r.stderr
    out = r.stdout.splitlines()
    vals = np.array([[float(x) for x in l.split()] for l in out[:33]])
    g = lambda n: read_golden(os.path.join(GOLDEN, "cases", "1", n))[:, 0]
    assert relerr(vals[:, 0], g("torque_grav_")) < 1e-9 and relerr(vals[:, 1], g("torque_task_")) < 1e-9 and relerr(vals[:, 2], g("torque_contact_")) < 1e-9
    assert out[33] == "status 7" and out[34].startswith("error ok: Link Name is Wrong")


def test_pybind_module_on_the_gpu():
    """`import dwbc` (the compiled pybind11 module with the reference's RlWBCBridge signature) against libdwbc_b200.so on the GPU."""
    from libdwbc_b200._lib import library_path
    from tests.test_bridge import run_pybind_module

    _dev()
    run_pybind_module(library_path(), os.path.dirname(library_path()))
