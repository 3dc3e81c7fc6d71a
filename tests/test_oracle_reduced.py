"""a9 reduced-dimension variant of the oracle.  The reference holds no goldens for it; its own check is the printed
"reduced torques == full-model torques" similarity (reference tests/sp_test/redu_dyn_test.cpp:323-360, threshold 1e-4).
The oracle restates src/dwbc.cpp:2752-3446 literally and must satisfy that property to 1e-9 on the golden
configurations and on synthetic states; structural facts from SURVEY.md Appendix B are checked as well."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests.common import *


def reduced_tasks(ref):
    """The hierarchy of reference tests/sp_test/redu_dyn_test.cpp:97-101: COM position / pelvis rotation /
    upper-body rotation / left-hand 6D (COM task, contact-chain task, two non-contact tasks)."""
    ref.add_task(0, orc.TASK_LINK_POSITION, "COM")
    ref.add_task(1, orc.TASK_LINK_ROTATION, "pelvis_link")
    ref.add_task(2, orc.TASK_LINK_ROTATION, "upperbody_link")
    ref.add_task(3, orc.TASK_LINK_6D, "L_Wrist2_Link")


def full_and_reduced(q, qd, flags, fstars):
    out = []
    for reduced in (False, True):
        ref = orc.OracleRobot(URDF)
        ref.add_contact(LEFT_FOOT, FOOT_POINT, (0, 0, 1), 0.15, 0.075)
        ref.add_contact(RIGHT_FOOT, FOOT_POINT, (0, 0, 1), 0.15, 0.075)
        reduced_tasks(ref)
        ref.update_kinematics(q, qd)
        ref.set_contact(*flags)
        for l, f in enumerate(fstars):
            ref.set_task(l, f)
        if not reduced:
            assert ref.calc_contact_constraint() == 1
            ref.calc_grav()
            assert ref.calc_task_torque(hqp=False) == 1
        else:
            ref.reduced_dynamics()
            assert ref.reduced_contact_constraint() == 1
            ref.reduced_grav()
            assert ref.reduced_task_space() == 1
            assert ref.reduced_task_torque(hqp=False) == 1
        out.append(ref)
    return out


FS = [np.array([0.3, 0.6, -0.3]), np.array([0.1, -0.1, 0.1]), np.array([-0.1, 0.1, -0.1]), np.array([0.05, 0.1, -0.05, 0.05, -0.05, 0.05])]


@pytest.mark.parametrize("case", [1, 2])
def test_reduced_equals_full_on_golden_configurations(case):
    full, red = full_and_reduced(np.array(Q_CASE[case]), np.zeros(39), (1, 1), FS)
    d = red.reduced_dims()
    assert d == dict(co_dof=12, nc_dof=21, vc_dof=18, reduced_system_dof=24, reduced_model_dof=18)  # reference src/dwbc.cpp:2818-2823
    assert abs(red.get("SI_nc_b")[0, 0] - 42.6195) < 1e-3  # nc-chain mass, reference tests/dwbc_test.cpp:664
    assert np.linalg.matrix_rank(red.get("W_R"), tol=1e-8) == 12
    assert relerr(red.get("A_R") @ red.get("A_R_inv"), np.eye(24)) < 1e-10
    assert relerr(red.get("J_R") @ full.get("A_inv") @ red.get("J_R").T, red.get("A_R_inv")) < 1e-12
    assert relerr(red.get("Lambda_CR"), full.get("Lambda_contact")) < 1e-9
    assert relerr(red.get("torque_grav")[:, 0], full.get("torque_grav")[:, 0]) < 1e-9
    assert relerr(red.get("torque_task")[:, 0], full.get("torque_task")[:, 0]) < 1e-9


def test_reduced_equals_full_on_synthetic_states():
    B = 6
    q, qd, rng = synth_states(B, 20243)
    for b in range(B):
        fs = [0.5 * rng.standard_normal(3), rng.standard_normal(3), rng.standard_normal(3), rng.standard_normal(6)]
        flags = (1, 1) if b != 3 else (1, 0)  # left-foot single support: co 6..11, nc 12..38 (the layout the reference supports)
        full, red = full_and_reduced(q[b], qd[b], flags, fs)
        if b == 3:
            assert red.reduced_dims()["reduced_system_dof"] == 18
        assert relerr(red.get("torque_grav")[:, 0], full.get("torque_grav")[:, 0]) < 1e-9
        assert relerr(red.get("torque_task")[:, 0], full.get("torque_task")[:, 0]) < 1e-8


def test_undefined_task_type_is_rejected():
    """A level mixing a contact-chain link and a non-contact link is the reference's "UNDEFINED TASK TYPE" (src/task.cpp:129-136)."""
    ref = orc.OracleRobot(URDF)
    configure_oracle(ref, 2)
    ref.update_kinematics(np.array(Q_CASE[1]), np.zeros(39))
    ref.set_contact(1, 1)
    for l, t in enumerate((3, 6, 21)):
        ref.set_task(l, np.zeros(t), posture_selector() if l == 2 else None)
    ref.reduced_dynamics()
    ref.reduced_contact_constraint()
    assert ref.reduced_task_space() == 0


def _hqp_pair(q, qd, flags, fstars, tlimR=None):
    """full-model HQP and reduced HQP (CalcSingleTaskTorqueWithQP_R, src/dwbc.cpp:3448-3599) on a COM / contact-chain hierarchy"""
    out = []
    for reduced in (False, True):
        ref = orc.OracleRobot(URDF)
        ref.add_contact(LEFT_FOOT, FOOT_POINT, (0, 0, 1), 0.15, 0.075)
        ref.add_contact(RIGHT_FOOT, FOOT_POINT, (0, 0, 1), 0.15, 0.075)
        ref.add_task(0, orc.TASK_LINK_POSITION, "COM")
        ref.add_task(1, orc.TASK_LINK_ROTATION, "pelvis_link")
        ref.update_kinematics(q, qd)
        ref.set_contact(*flags)
        for l, f in enumerate(fstars):
            ref.set_task(l, f)
        if not reduced:
            assert ref.calc_contact_constraint() == 1
            ref.calc_grav()
            ok = ref.calc_task_torque(hqp=True)
            ok2 = ref.calc_redistribute(hqp=True)
        else:
            if tlimR is not None:
                ref.set_reduced_torque_limit(tlimR)
            ref.reduced_dynamics()
            assert ref.reduced_contact_constraint() == 1
            ref.reduced_grav()
            assert ref.reduced_task_space() == 1
            ok = ref.reduced_task_torque(hqp=True)
            ok2 = ref.reduced_redistribute()
        out.append((ref, ok, ok2))
    return out


def test_reduced_hqp_equals_full_hqp_without_torque_limit():
    """No golden exists for the reduced QP branch.  Property: for COM / contact-chain tasks the reduced model is exact, so without
    a torque limit (the two models' limits are different objects) the reduced HQP reproduces the full-model HQP -- including a
    state whose lateral COM acceleration saturates the friction rows (f* = 4, reference tests/dwbc_test.cpp:56)."""
    B = 5
    q, qd, rng = synth_states(B, 20244)
    n_active = 0
    for b in range(B):
        fs = [0.5 * rng.standard_normal(3), rng.standard_normal(3)]
        if b % 2 == 0:
            fs[0][1] = 4.0
        (full, ok, ok2), (red, rok, rok2) = _hqp_pair(q[b], qd[b], (1, 1), fs)
        assert ok == 1 and rok == 1 and ok2 == 1 and rok2 == 1
        n_active += int(np.linalg.norm(full.get("f_star_qp0")) > 1e-6)
        assert relerr(red.get("f_star_qp0")[:, 0], full.get("f_star_qp0")[:, 0]) < 1e-6 or np.linalg.norm(full.get("f_star_qp0")) < 1e-9
        assert relerr(red.get("torque_task")[:, 0], full.get("torque_task")[:, 0]) < 1e-6
        # contact redistribution: the leg part of the full-model contact torque (upper-body entries of NwJw are ~0)
        tc_full = full.get("torque_contact")[:, 0]
        assert np.abs(red.get("torque_contact")[:12, 0] - tc_full[:12]).max() < 1e-6 * max(1.0, np.abs(tc_full).max())
    assert n_active >= 2


def test_reduced_hqp_with_reduced_torque_limit_respects_it():
    q, qd, rng = synth_states(2, 20245)
    fs = [np.array([0.3, 4.0, 0.2]), np.array([0.1, -0.1, 0.1])]
    (full, ok, ok2), (red, rok, rok2) = _hqp_pair(q[0], qd[0], (1, 1), fs, tlimR=np.full(18, 300.0))
    assert rok == 1
    tr = red.get("torque_grav_R")[:, 0] + red.get("torque_task_R")[:, 0]
    assert np.abs(tr).max() <= 300.0 + 1e-6


def test_reduced_hqp_with_non_contact_levels():
    """CalcSingleTaskTorqueWithQP_R_NC (reference src/dwbc.cpp:3601-3756), no golden: (1) when no row is active every QP returns
    zero and the hqp result equals the closed-form reduced result (which equals the full model, see above); (2) on a
    friction-saturating state the base-wrench QP is active, its answer moves only the contact-chain torques through
    J_base_R_kt, and it satisfies every row of the assembled QP with at least one of them active."""
    q, qd, rng = synth_states(4, 20249)
    n_active = 0
    for b in range(4):
        fs = [0.3 * rng.standard_normal(3), 0.3 * rng.standard_normal(3), 0.3 * rng.standard_normal(3), 0.3 * rng.standard_normal(6)]
        if b >= 1:
            # violent upper-body roll / lateral hand accelerations: the reaction of the non-contact chain on the base breaks the
            # ZMP / friction rows for scales around 7 (the COM-level null space hides most of it; above ~8 the QP is infeasible)
            sc = (0.0, 6.5, 7.0, 7.5)[b]
            fs = [0.3 * np.ones(3), 0.3 * np.ones(3), np.array([25.0 * sc, 0.0, 0.0]), np.array([0.0, 150.0 * sc, 0.0, 0.0, 0.0, 0.0])]
        res = []
        for hqp in (False, True):
            ref = orc.OracleRobot(URDF)
            ref.add_contact(LEFT_FOOT, FOOT_POINT, (0, 0, 1), 0.15, 0.075)
            ref.add_contact(RIGHT_FOOT, FOOT_POINT, (0, 0, 1), 0.15, 0.075)
            reduced_tasks(ref)
            ref.update_kinematics(q[b], qd[b])
            ref.set_contact(1, 1)
            for l, f in enumerate(fs):
                ref.set_task(l, f)
            ref.reduced_dynamics()
            assert ref.reduced_contact_constraint() == 1
            ref.reduced_grav()
            assert ref.reduced_task_space() == 1
            ok = ref.reduced_task_torque(hqp=hqp)
            assert ok == 1 or (hqp and b >= 1)
            res.append((ref, ok))
        if not res[1][1]:
            continue  # infeasible base-wrench QP: the reference returns 0 as well
        closed, hq = res[0][0], res[1][0]
        x6 = hq.get("force_on_nc_R_qp")[:, 0]
        dt = hq.get("torque_task")[:, 0] - closed.get("torque_task")[:, 0]
        if np.linalg.norm(x6) < 1e-9 and all(np.linalg.norm(hq.get(f"f_star_qp{l}")) < 1e-9 for l in range(2)):
            assert np.abs(dt).max() < 1e-8 * max(1.0, np.abs(closed.get("torque_task")).max())
        else:
            n_active += 1
            assert np.abs(dt[12:]).max() < 1e-9  # non-contact joints are untouched by the QPs
            # the answer satisfies every row of the QP the reference assembles, with at least one row active
            x = np.concatenate([x6, hq.get("nc_qp_contact")[:, 0]])
            slack = hq.get("nc_qpub")[:, 0] - hq.get("nc_qpA") @ x
            assert slack.min() > -1e-7 * max(1.0, np.abs(hq.get("nc_qpub")).max()) and (slack < 1e-6).sum() >= 1
    assert n_active >= 1


def test_reduced_analytic_redistribution_equals_full():
    """Analytic branch of CalcContactRedistributeR (reference src/dwbc.cpp:4938-4990), no golden: the reduced model is exact for
    the contact wrench, so the leg part of the redistribution torque equals the full-model CalcContactRedistribute(hqp=false)."""
    q, qd, rng = synth_states(3, 20250)
    for b in range(3):
        fs = [0.5 * rng.standard_normal(3), rng.standard_normal(3), rng.standard_normal(3), rng.standard_normal(6)]
        full, red = full_and_reduced(q[b], qd[b], (1, 1), fs)
        assert full.calc_redistribute(hqp=False) == 1 and red.reduced_redistribute(hqp=False) == 1
        tcf, tcr = full.get("torque_contact")[:, 0], red.get("torque_contact")[:, 0]
        assert np.abs(tcr[:12] - tcf[:12]).max() < 1e-7 * max(1.0, np.abs(tcf).max())
        assert np.all(tcr[12:] == 0.0)
