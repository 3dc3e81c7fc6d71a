"""a9 reduced-dimension variant: the product's kernels (host emulation build, tests/emul) against the oracle's literal
restatement of reference src/dwbc.cpp:2752-3446, and against the full-model torques (the reference's own check,
tests/sp_test/redu_dyn_test.cpp:323-360).  The authoritative run is the `-m gpu` twin in tests/test_gpu_parity.py."""
import numpy as np
import pytest

from libdwbc_b200 import CONTACT_6D, SOLVE_GRAV, SOLVE_TASK, TASK_LINK_6D, TASK_LINK_POSITION, TASK_LINK_ROTATION, DwbcError, RobotDataBatch
from oracle import oracle as orc
from tests.common import *
from tests.emul.emul_backend import EmulBackend
from tests.test_oracle_reduced import full_and_reduced


def reduced_product(backend, q, qd, flags, fstars, B):
    rd = RobotDataBatch(URDF, B, _backend=backend) if backend is not None else RobotDataBatch(URDF, B, 0)
    rd.add_contact_constraint(LEFT_FOOT, CONTACT_6D, FOOT_POINT, (0, 0, 1), 0.15, 0.075)
    rd.add_contact_constraint(RIGHT_FOOT, CONTACT_6D, FOOT_POINT, (0, 0, 1), 0.15, 0.075)
    rd.add_task_space(0, TASK_LINK_POSITION, "COM")
    rd.add_task_space(1, TASK_LINK_ROTATION, "pelvis_link")
    rd.add_task_space(2, TASK_LINK_ROTATION, "upperbody_link")
    rd.add_task_space(3, TASK_LINK_6D, "L_Wrist2_Link")
    return rd


REDUCED_NAMES = ["A_R_inv", "A_R", "J_I_nc", "J_I_nc_inv_T", "G_R", "Lambda_CR", "J_CR_INV_T", "A_R_inv_N_CR", "W_R", "W_R_inv", "NwJw_R", "P_CR",
                 "torque_grav_R", "J_base_R_kt", "torque_task_R"]


def check_reduced_against_oracle(rd, to_np, q, qd, flags, fs, tol=1e-9):
    B = q.shape[0]
    tg, tt, tc, st = [to_np(x) for x in rd.reduced_solve(SOLVE_GRAV | SOLVE_TASK)]
    c = 6 * int(sum(flags))
    for b in range(B):
        full, red = full_and_reduced(q[b], qd[b], flags, [f[b] for f in fs])
        assert st[b] == 3
        for name in REDUCED_NAMES:
            a = to_np(rd.get_reduced(name))[b]
            o = red.get(name)
            if o.shape[1] == 1:
                o = o[:, 0]
            if name in ("Lambda_CR", "J_CR_INV_T", "P_CR"):
                a = a[:c]
                a = a[:, :c] if name == "Lambda_CR" else a
            if name == "NwJw_R":
                a, o = a[:, :c - 6], o[:, :c - 6]
            assert relerr(a, o) < tol, (b, name, relerr(a, o))
        for l in range(4):
            assert relerr(to_np(rd.get_reduced("Lambda_task", l))[b], red.get(f"Lambda_task{l}")) < tol, l
        for l in range(2):
            assert relerr(to_np(rd.get_reduced("J_task_R", l))[b], red.get(f"R_J_task{l}")) < tol
            assert relerr(to_np(rd.get_reduced("J_kt_R", l))[b], red.get(f"R_J_kt{l}")) < tol
            assert relerr(to_np(rd.get_reduced("Null_task_R", l))[b], red.get(f"R_Null_task{l}")) < tol
        assert relerr(tg[b], red.get("torque_grav")[:, 0]) < tol
        assert relerr(tt[b], red.get("torque_task")[:, 0]) < tol
        # the reference's own check: reduced == full-model torques
        assert relerr(tg[b], full.get("torque_grav")[:, 0]) < 1e-8
        assert relerr(tt[b], full.get("torque_task")[:, 0]) < 1e-8
        assert np.all(tc[b] == 0)


@pytest.mark.parametrize("flags", [(1, 1), (1, 0)])
def test_emul_reduced_vs_oracle(flags):
    B = 3
    q, qd, rng = synth_states(B, 20243)
    fs = [0.5 * rng.standard_normal((B, 3)), rng.standard_normal((B, 3)), rng.standard_normal((B, 3)), rng.standard_normal((B, 6))]
    rd = reduced_product(EmulBackend(), q, qd, flags, fs, B)
    rd.update_kinematics(q, qd)
    fl = np.tile(np.array(flags, dtype=np.uint8), (B, 1))
    fl[2] = (1, 1) if flags == (1, 0) else (0, 1)  # a state that does not show the pattern
    rd.set_contact(fl)
    rd.reduced_dynamics_calculate(flags)
    assert (rd.co_dof, rd.nc_dof, rd.reduced_system_dof_, rd.reduced_model_dof_) == ((12, 21, 24, 18) if flags == (1, 1) else (6, 27, 18, 12))
    rd.reduced_calc_contact_constraint()
    for l, f in enumerate(fs):
        rd.set_task_space(l, np.ascontiguousarray(f))
    tg, tt, tc, st = rd.reduced_solve(SOLVE_GRAV | SOLVE_TASK)
    assert st[2] == 0 and np.all(tt[2] == 0)
    check_reduced_against_oracle(rd, lambda x: np.asarray(x)[:2], q[:2], qd[:2], flags, [f[:2] for f in fs])


def test_emul_reduced_rejections():
    rd = reduced_product(EmulBackend(), None, None, None, None, 1)
    q, qd, _ = synth_states(1, 1)
    with pytest.raises(DwbcError):
        rd.reduced_dynamics_calculate((1, 1))  # before update_kinematics / set_contact
    rd.update_kinematics(q, qd)
    rd.set_contact(np.array([[0, 1]], dtype=np.uint8))
    with pytest.raises(DwbcError, match="prefix"):
        rd.reduced_dynamics_calculate((0, 1))  # right-foot single support: the reference's block indexing breaks
    rd.set_contact(np.array([[1, 1]], dtype=np.uint8))
    rd.reduced_dynamics_calculate((1, 1))
    with pytest.raises(DwbcError):
        rd.reduced_solve(SOLVE_GRAV)  # before reduced_calc_contact_constraint
    rd.reduced_calc_contact_constraint()
    tg, _, _, st = rd.reduced_solve(SOLVE_GRAV)
    assert st[0] == 1 and np.abs(tg[0]).max() > 1
    # config-2 hierarchy: level 1 mixes pelvis (contact chain) and upper body (non-contact) -> UNDEFINED TASK TYPE
    rd2 = configure_product(RobotDataBatch(URDF, 1, _backend=EmulBackend()), 2)
    rd2.update_kinematics(q, qd)
    rd2.set_contact(np.array([[1, 1]], dtype=np.uint8))
    rd2.reduced_dynamics_calculate((1, 1))
    rd2.reduced_calc_contact_constraint()
    for l, t in enumerate((3, 6, 21)):
        rd2.set_task_space(l, np.zeros((1, t)), np.ascontiguousarray(posture_selector()) if l == 2 else None)
    with pytest.raises(DwbcError, match="UNDEFINED TASK TYPE"):
        rd2.reduced_solve(SOLVE_TASK)


def run_reduced_hqp(make_rd, D, H, B=4, tlimR=None):
    """Reduced HQP + reduced contact redistribution (CalcSingleTaskTorqueWithQP_R, CalcContactRedistributeR) on a COM /
    contact-chain hierarchy vs the oracle; every other state saturates the friction rows."""
    from libdwbc_b200 import CONTACT_6D, SOLVE_GRAV, SOLVE_HQP, SOLVE_REDIST, TASK_LINK_POSITION, TASK_LINK_ROTATION

    q, qd, rng = synth_states(B, 20246)
    fs = [0.5 * rng.standard_normal((B, 3)), rng.standard_normal((B, 3))]
    fs[0][::2, 1] = 4.0
    rd = make_rd(B)
    rd.add_contact_constraint(LEFT_FOOT, CONTACT_6D, FOOT_POINT, (0, 0, 1), 0.15, 0.075)
    rd.add_contact_constraint(RIGHT_FOOT, CONTACT_6D, FOOT_POINT, (0, 0, 1), 0.15, 0.075)
    rd.add_task_space(0, TASK_LINK_POSITION, "COM")
    rd.add_task_space(1, TASK_LINK_ROTATION, "pelvis_link")
    rd.set_torque_limit(300.0)  # full-model limit: must be ignored by the reduced QPs
    if tlimR is not None:
        rd.set_reduced_torque_limit(tlimR)
    rd.update_kinematics(D(q), D(qd))
    rd.set_contact(D(np.ones((B, 2), dtype=np.uint8)))
    rd.reduced_dynamics_calculate((1, 1))
    rd.reduced_calc_contact_constraint()
    for l, f in enumerate(fs):
        rd.set_task_space(l, D(f))
    tg, tt, tc, st = [H(x) for x in rd.reduced_solve(SOLVE_GRAV | SOLVE_HQP | SOLVE_REDIST)]
    n_active = 0
    for b in range(B):
        ref = orc.OracleRobot(URDF)
        ref.add_contact(LEFT_FOOT, FOOT_POINT, (0, 0, 1), 0.15, 0.075)
        ref.add_contact(RIGHT_FOOT, FOOT_POINT, (0, 0, 1), 0.15, 0.075)
        ref.add_task(0, orc.TASK_LINK_POSITION, "COM")
        ref.add_task(1, orc.TASK_LINK_ROTATION, "pelvis_link")
        if tlimR is not None:
            ref.set_reduced_torque_limit(tlimR)
        ref.update_kinematics(q[b], qd[b])
        ref.set_contact(1, 1)
        for l, f in enumerate(fs):
            ref.set_task(l, f[b])
        ref.reduced_dynamics()
        assert ref.reduced_contact_constraint() == 1
        ref.reduced_grav()
        assert ref.reduced_task_space() == 1
        ok = ref.reduced_task_torque(hqp=True)
        ok2 = ref.reduced_redistribute()
        assert (st[b] & 7) == (1 | (2 if ok else 0) | (4 if ok2 else 0)), (b, st[b], ok, ok2)
        n_active += int(np.linalg.norm(ref.get("f_star_qp0")) > 1e-6)
        assert relerr(tg[b], ref.get("torque_grav")[:, 0]) < 1e-9
        assert relerr(tt[b], ref.get("torque_task")[:, 0]) < 1e-6, b
        tcr = ref.get("torque_contact")[:, 0]
        assert np.abs(tc[b] - tcr).max() < 1e-6 * max(1.0, np.abs(tcr).max()), b
    assert n_active >= B // 2 - 1


# reduced limit: 300 on the 12 contact-chain joints; the 6 coordinates of the condensed non-contact body carry its weight
# (42.6 kg) and are not motor torques, so they get a wide bound
TLIM_R = np.concatenate([np.full(12, 300.0), np.full(6, 2000.0)])


@pytest.mark.parametrize("limited", [0, 1, 2])
def test_emul_reduced_hqp_vs_oracle(limited):
    """limited = 2: a uniform 300 limit, which the level QPs leave ACTIVE on a condensed-body coordinate; the redistribution QP then
    sees that row with a zero normal and b ~ -1e-8 -- feasible under the relative zero-row rule."""
    tl = (None, TLIM_R, np.full(18, 300.0))[limited]
    run_reduced_hqp(lambda B: RobotDataBatch(URDF, B, _backend=EmulBackend()), np.ascontiguousarray, lambda a: a, B=2, tlimR=tl)  # GPU twin: B = 4


def run_reduced_hqp_nc(make_rd, D, H, scales=(0.0, 6.5, 7.0, 7.5)):
    """Reduced HQP with non-contact levels (CalcSingleTaskTorqueWithQP_R + _R_NC, reference src/dwbc.cpp:3255-3446, 3601-3756) vs the
    oracle on the hierarchy of reference tests/sp_test/redu_dyn_test.cpp:97-101; scales around 7 make the base-wrench QP active."""
    from libdwbc_b200 import CONTACT_6D, SOLVE_GRAV, SOLVE_HQP, SOLVE_REDIST, TASK_LINK_6D, TASK_LINK_POSITION, TASK_LINK_ROTATION

    B = len(scales)
    q, qd, rng = synth_states(B, 20249)
    fs = [0.3 * rng.standard_normal((B, 3)), 0.3 * rng.standard_normal((B, 3)), 0.3 * rng.standard_normal((B, 3)), 0.3 * rng.standard_normal((B, 6))]
    for b, sc in enumerate(scales):
        if sc > 0:
            fs[0][b], fs[1][b] = 0.3, 0.3
            fs[2][b] = [25.0 * sc, 0.0, 0.0]
            fs[3][b] = [0.0, 150.0 * sc, 0.0, 0.0, 0.0, 0.0]
    rd = make_rd(B)
    rd.add_contact_constraint(LEFT_FOOT, CONTACT_6D, FOOT_POINT, (0, 0, 1), 0.15, 0.075)
    rd.add_contact_constraint(RIGHT_FOOT, CONTACT_6D, FOOT_POINT, (0, 0, 1), 0.15, 0.075)
    rd.add_task_space(0, TASK_LINK_POSITION, "COM")
    rd.add_task_space(1, TASK_LINK_ROTATION, "pelvis_link")
    rd.add_task_space(2, TASK_LINK_ROTATION, "upperbody_link")
    rd.add_task_space(3, TASK_LINK_6D, "L_Wrist2_Link")
    rd.update_kinematics(D(q), D(qd))
    rd.set_contact(D(np.ones((B, 2), dtype=np.uint8)))
    rd.reduced_dynamics_calculate((1, 1))
    rd.reduced_calc_contact_constraint()
    for l, f in enumerate(fs):
        rd.set_task_space(l, D(f))
    tg, tt, tc, st = [H(x) for x in rd.reduced_solve(SOLVE_GRAV | SOLVE_HQP | SOLVE_REDIST)]
    n_active = 0
    for b in range(B):
        ref = orc.OracleRobot(URDF)
        ref.add_contact(LEFT_FOOT, FOOT_POINT, (0, 0, 1), 0.15, 0.075)
        ref.add_contact(RIGHT_FOOT, FOOT_POINT, (0, 0, 1), 0.15, 0.075)
        ref.add_task(0, orc.TASK_LINK_POSITION, "COM")
        ref.add_task(1, orc.TASK_LINK_ROTATION, "pelvis_link")
        ref.add_task(2, orc.TASK_LINK_ROTATION, "upperbody_link")
        ref.add_task(3, orc.TASK_LINK_6D, "L_Wrist2_Link")
        ref.update_kinematics(q[b], qd[b])
        ref.set_contact(1, 1)
        for l, f in enumerate(fs):
            ref.set_task(l, f[b])
        ref.reduced_dynamics()
        assert ref.reduced_contact_constraint() == 1
        ref.reduced_grav()
        assert ref.reduced_task_space() == 1
        ok = ref.reduced_task_torque(hqp=True)
        assert bool(st[b] & 2) == bool(ok), (b, st[b], ok)
        if not ok:
            assert np.all(tt[b] == 0.0)
            continue
        n_active += int(np.linalg.norm(ref.get("force_on_nc_R_qp")) > 1e-6)
        want = ref.get("torque_task")[:, 0]
        assert np.abs(tt[b] - want).max() < 1e-6 * max(1.0, np.abs(want).max()), (b, np.abs(tt[b] - want).max())
        # ReducedCalcContactRedistribute on torque_grav_R_ + the final torque_task_R_ (reference src/dwbc.cpp:3758-3770)
        ok2 = ref.reduced_redistribute()
        assert bool(st[b] & 4) == bool(ok2), (b, st[b], ok2)
        if ok2:
            tcr = ref.get("torque_contact")[:, 0]
            assert np.abs(tc[b] - tcr).max() < 1e-6 * max(1.0, np.abs(tcr).max()), b
    return n_active


def test_emul_reduced_hqp_with_non_contact_levels():
    n_active = run_reduced_hqp_nc(lambda B: RobotDataBatch(URDF, B, _backend=EmulBackend()), np.ascontiguousarray, lambda a: a)
    assert n_active >= 1


def test_emul_reduced_analytic_redistribution():
    """analytic branch of CalcContactRedistributeR (reference src/dwbc.cpp:4938-4990) on the reduced workspace vs the oracle"""
    from libdwbc_b200 import SOLVE_REDIST_ANALYTIC

    B = 3
    q, qd, rng = synth_states(B, 20250)
    fs = [0.5 * rng.standard_normal((B, 3)), rng.standard_normal((B, 3)), rng.standard_normal((B, 3)), rng.standard_normal((B, 6))]
    rd = reduced_product(EmulBackend(), q, qd, (1, 1), fs, B)
    rd.update_kinematics(np.ascontiguousarray(q), np.ascontiguousarray(qd))
    rd.set_contact(np.ones((B, 2), dtype=np.uint8))
    rd.reduced_dynamics_calculate((1, 1))
    rd.reduced_calc_contact_constraint()
    for l, f in enumerate(fs):
        rd.set_task_space(l, np.ascontiguousarray(f))
    tg, tt, tc, st = rd.reduced_solve(SOLVE_GRAV | SOLVE_TASK | SOLVE_REDIST_ANALYTIC)
    for b in range(B):
        full, red = full_and_reduced(q[b], qd[b], (1, 1), [f[b] for f in fs])
        assert red.reduced_redistribute(hqp=False) == 1 and (st[b] & 7) == 7
        tcr = red.get("torque_contact")[:, 0]
        assert np.abs(tc[b] - tcr).max() < 1e-8 * max(1.0, np.abs(tcr).max()), b
