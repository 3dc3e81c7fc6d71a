"""a10 LQP variant: kernel logic (emulation build, no GPU) vs the oracle's exact level-by-level solve.  The authoritative
comparison is tests/test_gpu_parity.py::test_gpu_lqp; PARITY UNPINNED w.r.t. the reference (OSQP, no goldens)."""
import numpy as np
import pytest

from libdwbc_b200 import RobotDataBatch
from oracle import oracle as orc
from tests.common import *
from tests.emul.emul_backend import EmulBackend


def lqp_oracle_states(q, qd, flags, fstars):
    res = []
    for b in range(q.shape[0]):
        r = configure_oracle(orc.OracleRobot(URDF), "lqp")
        r.update_kinematics(q[b], qd[b])
        r.set_contact(*flags[b])
        ok_cc = r.calc_contact_constraint()
        for l, f in enumerate(fstars):
            r.set_task(l, f[b])
        r.update_task_space()
        nl = r.configure_lqp()
        ok = r.calc_lqp()
        res.append((r, nl, ok if ok_cc == 1 else 0))
    return res


def check_lqp(rd, out, refs, flags, tol=1e-6):
    y, tau, st = out
    n = 39
    for b, (r, nl, ok) in enumerate(refs):
        c = 6 * int(sum(flags[b]))
        assert int(st[b]) == (nl if ok else st[b]) and (ok == 1) == (int(st[b]) == nl), (b, st[b], nl, ok)
        if not ok:
            continue
        for l in range(nl):
            yl = rd.get_lqp("y", l)[b]
            yr = r.get(f"lqp_y{l}")[:, 0]
            assert relerr(yl[:n + c], yr) < tol, (b, l, relerr(yl[:n + c], yr))
            assert np.all(yl[n + c:] == 0)
        v1 = rd.get_lqp("v1")[b]
        vr = r.get("lqp_v1")[:, 0]
        assert np.linalg.norm(v1[:len(vr)] - vr) < tol * max(1.0, np.linalg.norm(vr)), b
        assert relerr(tau[b], r.get("lqp_torque")[:, 0]) < tol, b
        assert relerr(y[b][:n + c], r.get(f"lqp_y{nl - 1}")[:, 0]) < tol


@pytest.mark.parametrize("hard", [False, True])
def test_emul_lqp_vs_oracle(hard):
    B = 3
    q, qd, rng = synth_states(B, 20243)
    fs = lqp_fstars(rng, B, hard)
    flags = np.ones((B, 2), dtype=np.uint8)
    flags[1] = [1, 0]
    flags[2] = [0, 1]
    rd = configure_product(RobotDataBatch(URDF, B, _backend=EmulBackend()), "lqp")
    rd.update_kinematics(np.ascontiguousarray(q), np.ascontiguousarray(qd))
    rd.set_contact(flags)
    rd.calc_contact_constraint()
    for l, f in enumerate(fs):
        rd.set_task_space(l, np.ascontiguousarray(f))
    out = rd.calc_control_torque_lqp()
    refs = lqp_oracle_states(q, qd, flags, fs)
    check_lqp(rd, out, refs, flags)


def run_reduced_lqp(make_rd, D, H, B=3):
    """ConfigureLQP_R + CalcControlTorqueLQP_R (reference src/dwbc.cpp:4504-4632) vs the oracle: hierarchy = COM position +
    pelvis rotation (contact-chain level) followed by a non-contact level that the reduced LQP leaves out."""
    from libdwbc_b200 import CONTACT_6D, SOLVE_GRAV, SOLVE_TASK, TASK_LINK_6D, TASK_LINK_POSITION, TASK_LINK_ROTATION

    q, qd, rng = synth_states(B, 20249)
    fs = lqp_fstars(rng, B, False)
    fs[0][0, 1] = 12.0  # one state far beyond the friction cone
    rd = make_rd(B)
    rd.add_contact_constraint(LEFT_FOOT, CONTACT_6D, FOOT_POINT, (0, 0, 1), 0.13, 0.06)
    rd.add_contact_constraint(RIGHT_FOOT, CONTACT_6D, FOOT_POINT, (0, 0, 1), 0.13, 0.06)
    rd.add_task_space(0, TASK_LINK_POSITION, "COM")
    rd.add_task_space(0, TASK_LINK_ROTATION, "pelvis_link")
    rd.add_task_space(1, TASK_LINK_6D, "L_Wrist2_Link")
    rd.update_kinematics(D(q), D(qd))
    rd.set_contact(D(np.ones((B, 2), dtype=np.uint8)))
    rd.reduced_dynamics_calculate((1, 1))
    rd.reduced_calc_contact_constraint()
    for l, f in enumerate(fs):
        rd.set_task_space(l, D(f))
    rd.reduced_solve(SOLVE_GRAV | SOLVE_TASK)
    y, tau, st = [H(x) for x in rd.calc_control_torque_lqp_r()]
    ync, tau33, stnc = [H(x) for x in rd.calc_control_torque_lqp_r_nc()]
    n = 24
    for b in range(B):
        r = orc.OracleRobot(URDF)
        r.add_contact(LEFT_FOOT, FOOT_POINT, (0, 0, 1), 0.13, 0.06)
        r.add_contact(RIGHT_FOOT, FOOT_POINT, (0, 0, 1), 0.13, 0.06)
        r.add_task(0, orc.TASK_LINK_POSITION, "COM")
        r.add_task(0, orc.TASK_LINK_ROTATION, "pelvis_link")
        r.add_task(1, orc.TASK_LINK_6D, "L_Wrist2_Link")
        r.update_kinematics(q[b], qd[b])
        r.set_contact(1, 1)
        for l, f in enumerate(fs):
            r.set_task(l, f[b])
        r.reduced_dynamics()
        assert r.reduced_contact_constraint() == 1
        assert r.reduced_task_space() == 1
        nl = r.configure_lqp(reduced=True)
        ok = r.calc_lqp()
        assert nl == 3 and ok == 1 and int(st[b]) == 3, (b, nl, ok, st[b])
        for l in range(nl):
            yl = H(rd.get_lqp("y", l, reduced=True))[b]
            assert relerr(yl[:n + 12], r.get(f"lqp_y{l}")[:, 0]) < 1e-6, (b, l)
        assert relerr(tau[b], r.get("lqp_torque")[:, 0]) < 1e-6, b
        assert relerr(y[b][:n + 12], r.get(f"lqp_y{nl - 1}")[:, 0]) < 1e-6
        # ConfigureLQP_R_NC + CalcControlTorqueLQP_R_NC (reference src/dwbc.cpp:4634-4760, 4479-4502) and the stitched torque
        r.configure_lqp_r_nc()
        assert r.calc_lqp_r_nc() == 1 and int(stnc[b]) == 2, (b, stnc[b])
        assert relerr(ync[b], r.get("lqpnc_y2")[:, 0]) < 1e-6, b
        assert relerr(tau33[b], r.get("lqp_stitched_torque")[:, 0]) < 1e-6, b


def test_emul_reduced_lqp_vs_oracle():
    run_reduced_lqp(lambda B: RobotDataBatch(URDF, B, _backend=EmulBackend()), np.ascontiguousarray, lambda a: a)
