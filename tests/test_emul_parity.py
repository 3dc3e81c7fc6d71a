"""Kernel logic vs oracle WITHOUT a GPU: the product's kernel sources compiled for the host through
tests/emul/cuda_emul.h (test infrastructure) and driven through the same C ABI.  These tests catch algorithmic
regressions in the container; the authoritative parity tests are the `-m gpu` ones (tests/test_gpu_parity.py)."""
import os

import numpy as np
import pytest

from libdwbc_b200 import SOLVE_GRAV, SOLVE_HQP, SOLVE_REDIST, SOLVE_TASK, RobotDataBatch
from oracle import oracle as orc
from oracle.urdf_model import read_golden
from tests.common import *
from tests.emul.emul_backend import EmulBackend


def run_product(config, q, qd, fstars, flags, stages, ncontacts=2):
    B = q.shape[0]
    rd = configure_product(RobotDataBatch(URDF, B, _backend=EmulBackend()), config, ncontacts)
    rd.update_kinematics(np.ascontiguousarray(q), np.ascontiguousarray(qd))
    rd.set_contact(np.ascontiguousarray(flags, dtype=np.uint8))
    rd.calc_contact_constraint()
    for l, f in enumerate(fstars):
        rd.set_task_space(l, np.ascontiguousarray(f), np.ascontiguousarray(posture_selector()) if (config == 2 and l == 2) else None)
    out = rd.solve(stages)
    return rd, out


@pytest.mark.parametrize("case", [1, 2])
def test_emul_golden(case):
    q = np.array([Q_CASE[case]], dtype=np.float64)
    flags = np.array([[1, 1, 0, 0]], dtype=np.uint8)
    fs = [np.array([FSTAR_CASE[case][0]]), np.array([FSTAR_CASE[case][1]])]
    rd, (tg, tt, tc, st) = run_product("golden", q, np.zeros((1, 39)), fs, flags, SOLVE_GRAV | SOLVE_HQP | SOLVE_REDIST, 4)
    assert st[0] == 7
    g = lambda n: read_golden(os.path.join(GOLDEN, "cases", str(case), n))
    for name, gname in GOLDEN_NAMES:
        a = rd.get(name)[0]
        gg = g(gname)
        if a.ndim == 1:
            gg = gg[:, 0]
        a = a[:gg.shape[0]] if a.ndim == 2 else a
        a = a[:, :gg.shape[1]] if a.ndim == 2 else a
        assert relerr(a, gg) < 1e-9, name
    assert relerr(tt[0], g("torque_task_")[:, 0]) < 1e-9
    assert relerr(tc[0], g("torque_contact_")[:, 0]) < (1e-9 if case == 1 else 1e-4)
    if case == 1:
        assert rd.active_set(0)[0] == [66, 67, 68, 69, 70, 71, 72, 73, 74, 75, 78]
        assert rd.active_set(1)[0] == [67, 68, 71, 72, 75, 78]
    else:
        assert rd.active_set(1)[0] == [68, 70]


def test_emul_config2_vs_oracle():
    """3-level HQP, double and single support, including the forced-active state."""
    B = 6
    q, qd, rng = synth_states(B, 20242)
    fs = synth_fstar(rng, B)
    fs[0][0, 1] = 4.0
    flags = np.ones((B, 2), dtype=np.uint8)
    flags[4] = [1, 0]
    flags[5] = [0, 1]
    rd, (tg, tt, tc, st) = run_product(2, q, qd, fs, flags, SOLVE_GRAV | SOLVE_HQP | SOLVE_REDIST)
    ref = configure_oracle(orc.OracleRobot(URDF), 2)
    o = ref.run_batch(q, qd, fs, flags)
    assert np.array_equal(st & 7, o["status"] & 7)
    for b in range(B):
        if (o["status"][b] & 7) != 7:
            continue
        assert relerr(tg[b], o["torque_grav"][b]) < 1e-9
        assert relerr(tt[b], o["torque_task"][b]) < 1e-6
        assert np.abs(tc[b] - o["torque_contact"][b]).max() < 1e-6 * max(1.0, np.abs(o["torque_contact"][b]).max())


def test_emul_config1_closed_form():
    B = 3
    q, qd, rng = synth_states(B, 20241)
    fs = synth_fstar(rng, B, dofs=(3, 3), scales=(0.5, 1.0), hard_frac=0)
    flags = np.ones((B, 2), dtype=np.uint8)
    rd, (tg, tt, tc, st) = run_product(1, q, qd, fs, flags, SOLVE_GRAV | SOLVE_TASK)
    ref = configure_oracle(orc.OracleRobot(URDF), 1)
    for b in range(B):
        ref.update_kinematics(q[b], qd[b]); ref.set_contact(1, 1); ref.calc_contact_constraint()
        ref.set_task(0, fs[0][b]); ref.set_task(1, fs[1][b]); ref.calc_grav(); ref.calc_task_torque(False)
        assert relerr(tt[b], ref.get("torque_task")[:, 0]) < 1e-9
        for l in (0, 1):
            assert relerr(rd.get("Lambda_task", l)[b], ref.get(f"Lambda_task{l}")) < 1e-9
            assert relerr(rd.get("J_kt", l)[b], ref.get(f"J_kt{l}")) < 1e-9
            assert relerr(rd.get("J_task", l)[b], ref.get(f"J_task{l}")) < 1e-12
        assert relerr(rd.get("Null_task", 0)[b], ref.get("Null_task0")) < 1e-9
        assert relerr(rd.get("B")[b], ref.get("B")[:, 0]) < 1e-10
        assert relerr(rd.get("com_pos")[b], ref.get("com_pos")[:, 0]) < 1e-12
        assert relerr(rd.get("link_pos")[b], ref.get("link_pos")) < 1e-12
        assert relerr(rd.get("link_rot")[b], ref.get("link_rot")) < 1e-12
        assert relerr(rd.get("link_vel")[b], ref.get("link_vel")) < 1e-11


def test_emul_analytic_redistribution_contact_force_zmp():
    """CalcTaskControlTorque(hqp=false) + CalcContactRedistribute(hqp=false) (the path the reference's Python bridge runs,
    src/pybind/rl_wbc_bridge.cpp:123-129), getContactForce and getZMP against the oracle."""
    from libdwbc_b200 import SOLVE_REDIST_ANALYTIC
    B = 12
    q, qd, rng = synth_states(B, 20251)
    q[:, 3:6] *= 3.0  # larger base tilt / yaw of either sign: exercises the Eigen eulerAngles branch
    q[:, 39] = np.sqrt(1.0 - np.sum(q[:, 3:6] ** 2, axis=1))
    fs = synth_fstar(rng, B, dofs=config_dofs(1), scales=(0.5, 1.0))
    flags = np.ones((B, 2), dtype=np.uint8)
    flags[5] = [1, 0]
    rd, (tg, tt, tc, st) = run_product(1, q, qd, fs, flags, SOLVE_GRAV | SOLVE_TASK | SOLVE_REDIST_ANALYTIC)
    ref = configure_oracle(orc.OracleRobot(URDF), 1)
    o = ref.run_batch(q, qd, fs, flags, stages=1 | 2 | 16)
    assert np.array_equal(st & 7, o["status"] & 7)
    assert (o["status"][np.arange(B) != 5] & 4).all() and not (o["status"][5] & 4)
    assert np.abs(o["torque_contact"]).max() > 1e-3
    for name, a in (("torque_grav", tg), ("torque_task", tt), ("torque_contact", tc)):
        assert relerr(a, o[name]) < 1e-9, name
    cf = rd.get_contact_force()
    assert relerr(cf, o["contact_force"]) < 1e-9
    zmp = rd.get_zmp(np.ascontiguousarray(cf))
    assert relerr(zmp, o["zmp"]) < 1e-9
    # explicit command torque
    tau = np.ascontiguousarray(tg + tt)
    cf2 = rd.get_contact_force(tau)
    ref.update_kinematics(q[0], qd[0]); ref.set_contact(1, 1); ref.calc_contact_constraint(); ref.calc_grav()
    assert relerr(cf2[0], ref.contact_force(tau[0])) < 1e-9


def run_tight_torque_limit(make_rd, D, H, B, tl=45.0):
    """Config-2 hierarchy with a torque limit low enough that the limit rows are active in most states (the goldens and the
    workload rarely activate them): product vs oracle, and the limit holds for the summed torque."""
    from libdwbc_b200 import CONTACT_6D, TASK_LINK_POSITION, TASK_LINK_ROTATION

    q, qd, rng = synth_states(B, 20255)
    fs = synth_fstar(rng, B)
    flags = np.ones((B, 2), dtype=np.uint8)
    rd = make_rd(B)
    rd.add_contact_constraint(LEFT_FOOT, CONTACT_6D, FOOT_POINT, (0, 0, 1), 0.15, 0.075)
    rd.add_contact_constraint(RIGHT_FOOT, CONTACT_6D, FOOT_POINT, (0, 0, 1), 0.15, 0.075)
    rd.add_task_space(0, TASK_LINK_POSITION, "COM")
    rd.add_task_space(1, TASK_LINK_ROTATION, "pelvis_link")
    rd.add_task_space(1, TASK_LINK_ROTATION, "upperbody_link")
    rd.add_custom_task_space(2, 21)
    rd.set_torque_limit(tl)
    rd.update_kinematics(D(q), D(qd))
    rd.set_contact(D(flags))
    rd.calc_contact_constraint()
    for l, f in enumerate(fs):
        rd.set_task_space(l, D(f), D(posture_selector()) if l == 2 else None)
    tg, tt, tc, st = [H(x) for x in rd.solve(SOLVE_GRAV | SOLVE_HQP | SOLVE_REDIST)]
    ref = configure_oracle(orc.OracleRobot(URDF), 2)
    ref.set_torque_limit(tl)
    o = ref.run_batch(q, qd, fs, flags)
    assert np.array_equal(st & 7, o["status"] & 7)
    ok = (o["status"] & 7) == 7
    assert ok.sum() >= B // 2
    n_at_limit = 0
    for b in np.where(ok)[0]:
        for a, k in ((tt, "torque_task"), (tc, "torque_contact")):
            assert np.abs(a[b] - o[k][b]).max() < 1e-6 * max(1.0, np.abs(o[k][b]).max()), (b, k)
        tot = np.abs(tg[b] + tt[b] + tc[b]).max()
        assert tot <= tl + 1e-6
        n_at_limit += int(tot > tl - 1e-6)
    assert n_at_limit >= ok.sum() // 3


def test_emul_tight_torque_limit():
    run_tight_torque_limit(lambda B: RobotDataBatch(URDF, B, _backend=EmulBackend()), np.ascontiguousarray, lambda a: a, 3)


def run_sliced_cycle_host(make_rd, D, B, variants, outputs=1):
    """dwbc_batch_cycle_host cut into slices (copies of one slice under the kernels of another) returns bit for bit what the
    unsliced call returns: ragged slice sizes, mixed support, shared and per-state custom Jacobian."""
    q, qd, rng = synth_states(B, 11)
    fs = synth_fstar(rng, B)
    flags = np.ones((B, 2), dtype=np.uint8)
    flags[::3, 1] = 0
    flags[1::6, 0] = 0
    Jp = posture_selector()
    Jps = np.ascontiguousarray(np.broadcast_to(Jp, (B,) + Jp.shape))  # the same selector handed over per state
    ALL = SOLVE_GRAV | SOLVE_HQP | SOLVE_REDIST
    outs = []
    for slices, per_state in variants:
        rd = configure_product(make_rd(B), 2)
        rd.set_outputs(outputs)  # 1 = OUT_ALL: the generic kernels, comparable bit for bit with the staged calls; 0 = the one-kernel cycle
        rd.set_task_space(2, D(np.zeros((B, 21))), D(Jps if per_state else Jp))
        rd.set_host_slices(slices)
        outs.append([np.array(a) for a in rd.cycle_host(ALL, q, qd, fs, flags)])
        assert rd.batch == B and rd.get("A" if outputs == 1 else "torque_grav").shape[0] == B  # the handle describes the whole batch again afterwards
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            assert np.array_equal(a, b)
    assert ((outs[0][3] & 1) == 1).all()
    return (q, qd, fs, flags), outs[0]


def test_emul_walking_state_of_data_confirmation():
    """reference tests/sp_test/data_confirmation.cpp:41-84: a captured walking state, four registered contacts (feet active), four
    hierarchy levels (3 / 3 / 3 / 12), HQP + redistribution, no torque limit."""
    from tests.common import check_walking_state, walking_inputs

    q, qd, fs, flags = walking_inputs(1)
    rd, out = run_product("walking", q, qd, fs, flags, SOLVE_GRAV | SOLVE_HQP | SOLVE_REDIST)
    check_walking_state(rd, out, np.asarray)


def test_emul_sliced_cycle_host():
    run_sliced_cycle_host(lambda B: RobotDataBatch(URDF, B, _backend=EmulBackend()), np.ascontiguousarray, 3, ((1, False), (3, True)))
    run_sliced_cycle_host(lambda B: RobotDataBatch(URDF, B, _backend=EmulBackend()), np.ascontiguousarray, 3, ((1, False), (2, True)), outputs=0)


def test_emul_split_call_order_equals_fused():
    """The reference's split order -- CalcGravCompensation, CalcTaskControlTorque(true), CalcContactRedistribute(true) as three
    calls (tests/dwbc_test.cpp:100-130) -- must give the torques of the fused solve: the redistribution reads
    torque_grav_ + torque_task_ + torque_contact_ left by the earlier calls (reference src/dwbc.cpp:1372-1375)."""
    B = 2
    q, qd, rng = synth_states(B, 20242)
    fs = synth_fstar(rng, B)
    fs[0][0, 1] = 4.0
    flags = np.ones((B, 2), dtype=np.uint8)
    rd, (tg, tt, tc, st) = run_product(2, q, qd, fs, flags, SOLVE_GRAV | SOLVE_HQP | SOLVE_REDIST)
    tg, tt, tc, st = tg.copy(), tt.copy(), tc.copy(), st.copy()
    rd2, _ = run_product(2, q, qd, fs, flags, SOLVE_GRAV)
    _, tt2, tc_mid, _ = rd2.solve(SOLVE_GRAV | SOLVE_TASK | SOLVE_HQP)
    tt2, tc_mid = tt2.copy(), tc_mid.copy()
    tg3, tt3, tc3, st3 = rd2.solve(SOLVE_REDIST)
    assert np.array_equal(st3 & 7, st & 7)
    assert np.array_equal(tg3, tg) and np.array_equal(tt3, tt) and np.array_equal(tt2, tt)
    assert np.array_equal(tc3, tc)
    assert np.abs(tt).max() > 0 and np.abs(tc).max() > 0  # non-trivial torques: a wiped torque_task_ / torque_contact_ would show
    # a lone gravity call afterwards leaves the task / contact torques in place
    tg4, tt4, tc4, _ = rd2.solve(SOLVE_GRAV)
    assert np.array_equal(tt4, tt) and np.array_equal(tc4, tc)
