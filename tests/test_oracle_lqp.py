"""a10 LQP variant (reference src/dwbc.cpp:4304-4454, src/dwbc_hqp.cpp:23-85,291-403): the oracle's restatement.

PARITY UNPINNED: the reference holds no golden vector for this path and solves each level with OSQP (tolerance 1e-3).
What is checked here, independently of the oracle's solver (numpy only):
  * the level QPs are the ones the reference assembles (structure identities on H, A, ubA);
  * each level's answer satisfies the KKT conditions of that QP with sigma = 1e-6 on the null-space coordinates;
  * hierarchy: a later level never changes an earlier level's equality residual w or violates its frozen slack v;
  * physics: level 0 (floating-base rows of the equation of motion) holds exactly, and when nothing saturates the
    slack of the friction / acceleration rows is zero.
"""
import numpy as np
import pytest

from oracle import oracle as orc
from tests.common import *

SIGMA = 1e-6


def lqp_oracle(q, qd, flags, fstars, lqp_tasks="jacc"):
    """Tasks of reference tests/sp_test/jacc_compare.cpp:319-334: L0 = COM position + pelvis rotation (t = 6),
    L1 = left hand 6D."""
    r = orc.OracleRobot(URDF)
    r.add_contact(LEFT_FOOT, FOOT_POINT, (0, 0, 1), 0.13, 0.06)
    r.add_contact(RIGHT_FOOT, FOOT_POINT, (0, 0, 1), 0.13, 0.06)
    r.add_task(0, orc.TASK_LINK_POSITION, "COM")
    r.add_task(0, orc.TASK_LINK_ROTATION, "pelvis_link")
    r.add_task(1, orc.TASK_LINK_6D, "L_Wrist2_Link")
    r.update_kinematics(q, qd)
    r.set_contact(*flags)
    assert r.calc_contact_constraint() == 1
    for l, f in enumerate(fstars):
        r.set_task(l, f)
    r.update_task_space()
    nl = r.configure_lqp()
    ok = r.calc_lqp()
    return r, nl, ok


def kkt_residual(H, g, A, ub, z, nu):
    """Residual of the KKT conditions of min 1/2 z'(H + sigma I_u) z + g'z s.t. A z <= ub, with multipliers from NNLS."""
    from scipy.optimize import nnls
    Hs = H.copy()
    Hs[:nu, :nu] += SIGMA * np.eye(nu)
    grad = Hs @ z + g
    slack = ub - A @ z
    assert slack.min() > -1e-7, slack.min()
    act = np.where(slack < 1e-7)[0]
    if len(act) == 0:
        return np.linalg.norm(grad), act
    lam, rn = nnls(A[act].T, -grad)
    return rn, act


@pytest.mark.parametrize("flags", [(1, 1), (1, 0)])
@pytest.mark.parametrize("hard", [False, True])
def test_lqp_oracle_kkt_and_hierarchy(flags, hard):
    q, qd, rng = synth_states(3, 20243)
    for b in range(3):
        f0 = np.concatenate([0.5 * rng.standard_normal(3), 1.0 * rng.standard_normal(3)])
        f1 = 1.0 * rng.standard_normal(6)
        if hard:
            f0[1] = 12.0  # lateral COM acceleration far beyond the friction cone: contact rows saturate, slack v > 0 possible
            f1[:3] *= 30.0
        r, nl, ok = lqp_oracle(q[b], qd[b], flags, [f0, f1])
        assert nl == 4 and ok == 1
        n, c = 39, 6 * sum(flags)
        A, Bn, Jc = r.get("A"), r.get("B")[:, 0], r.get("J_C")
        ys = [r.get(f"lqp_y{l}")[:, 0] for l in range(nl)]
        # level 0: preset answer (reference src/dwbc.cpp:4381)
        assert relerr(ys[0][:n], -np.linalg.solve(A, Bn)) < 1e-10 and np.all(ys[0][n:] == 0)
        for l in range(1, nl):
            H, g = r.get(f"lqp_qpH{l}"), r.get(f"lqp_qpg{l}")[:, 0]
            QA, ub = r.get(f"lqp_qpA{l}"), r.get(f"lqp_qpub{l}")[:, 0]
            Zp = r.get(f"lqp_Z{l - 1}")
            nu = Zp.shape[1]
            assert np.allclose(Zp.T @ Zp, np.eye(nu), atol=1e-12)  # orthonormal null-space basis
            u, v = r.get(f"lqp_u{l}")[:, 0], r.get(f"lqp_v{l}")[:, 0]
            z = np.concatenate([u, v])
            assert H.shape == (nu + len(v), nu + len(v)) and QA.shape[1] == nu + len(v)
            Bl, bl = r.get(f"lqp_B{l}"), r.get(f"lqp_b{l}")[:, 0]
            E = Bl @ Zp
            Hc = Zp.T @ r.get("lqp_H") @ Zp
            assert relerr(H[:nu, :nu], E.T @ E + Hc) < 1e-12  # reference src/dwbc_hqp.cpp:337,351
            res, act = kkt_residual(H, g, QA, ub, z, nu)
            assert res < 1e-8 * max(1.0, np.linalg.norm(g)), (l, res)
            assert relerr(ys[l], ys[l - 1] + Zp @ u) < 1e-12
            assert len(v) == 0 or v.min() > -1e-9
        # hierarchy: earlier levels' equality residuals and slacks are untouched by later levels
        for l in range(nl):
            Bl, bl = r.get(f"lqp_B{l}"), r.get(f"lqp_b{l}")[:, 0]
            wl = Bl @ ys[max(l, 0)] + bl if l > 0 else Bl @ ys[0] + bl
            for k in range(l + 1, nl):
                assert np.linalg.norm(Bl @ ys[k] + bl - wl) < 1e-8, (l, k)
            if l < 2:
                Al, al = r.get(f"lqp_A{l}"), r.get(f"lqp_a{l}")[:, 0]
                vl = r.get(f"lqp_v{l}")[:, 0]
                for k in range(max(l, 1), nl):
                    assert (Al @ ys[k] + al - vl).max() < 1e-7, (l, k)
        # physics: floating-base rows of the equation of motion hold for the final answer
        y = ys[-1]
        eom = A @ y[:n] + Jc.T @ y[n:] + Bn
        assert np.linalg.norm(eom[:6]) < 1e-8 * np.linalg.norm(Bn)
        assert relerr(r.get("lqp_torque")[:, 0], eom[6:]) < 1e-12
        if not hard:
            # the cost 5 A/|A| shares each level with its equality rows (reference src/dwbc_hqp.cpp:351), so the rows are met
            # in the least-squares sense only; the slack of the friction / acceleration rows stays zero
            if sum(flags) == 2:  # single support: the ZMP rows of a random posture can be physically unattainable
                assert np.linalg.norm(r.get("lqp_v1")) < 1e-3  # sigma-regularisation bias only
            assert np.linalg.norm(Jc @ y[:n]) < 0.1 * np.linalg.norm(Jc @ ys[0][:n])


def test_reduced_lqp_oracle_kkt_and_physics():
    """ConfigureLQP_R (reference src/dwbc.cpp:4504-4632): the same hierarchy on the reduced model.  KKT of every level, the
    reduced floating-base equation of motion (the reduced variant uses gravity G_R only) and the null-space hierarchy."""
    q, _, rng = synth_states(3, 20248)
    for b in range(3):
        f0 = np.concatenate([0.5 * rng.standard_normal(3), 1.0 * rng.standard_normal(3)])
        res = []
        for reduced in (False, True):
            r = orc.OracleRobot(URDF)
            r.add_contact(LEFT_FOOT, FOOT_POINT, (0, 0, 1), 0.13, 0.06)
            r.add_contact(RIGHT_FOOT, FOOT_POINT, (0, 0, 1), 0.13, 0.06)
            r.add_task(0, orc.TASK_LINK_POSITION, "COM")
            r.add_task(0, orc.TASK_LINK_ROTATION, "pelvis_link")
            r.update_kinematics(q[b], np.zeros(39))
            r.set_contact(1, 1)
            r.set_task(0, f0)
            if reduced:
                r.reduced_dynamics()
                assert r.reduced_contact_constraint() == 1
                assert r.reduced_task_space() == 1
            else:
                assert r.calc_contact_constraint() == 1
                r.update_task_space()
            nl = r.configure_lqp(reduced)
            assert nl == 3 and r.calc_lqp() == 1
            res.append(r)
        full, red = res
        n = 24
        for l in range(1, 3):
            H, g = red.get(f"lqp_qpH{l}"), red.get(f"lqp_qpg{l}")[:, 0]
            QA, ub = red.get(f"lqp_qpA{l}"), red.get(f"lqp_qpub{l}")[:, 0]
            u, v = red.get(f"lqp_u{l}")[:, 0], red.get(f"lqp_v{l}")[:, 0]
            resid, act = kkt_residual(H, g, QA, ub, np.concatenate([u, v]), len(u))
            assert resid < 1e-8 * max(1.0, np.linalg.norm(g)), (l, resid)
        y = red.get("lqp_y2")[:, 0]
        AR, GR, JCR = red.get("A_R"), red.get("G_R")[:, 0], red.get("J_CR")
        eom = AR @ y[:n] + JCR.T @ y[n:] + GR
        assert np.linalg.norm(eom[:6]) < 1e-8 * np.linalg.norm(GR)
        assert relerr(red.get("lqp_torque")[:, 0], eom[6:]) < 1e-12
        # hierarchy: the task level works in the null space of the contact level (its equality residual is unchanged)
        B1, y1 = red.get("lqp_B1"), red.get("lqp_y1")[:, 0]
        assert np.linalg.norm(B1 @ (y - y1)) < 1e-8
        assert full.get("lqp_y2").shape[0] == 51 and len(y) == 36


def reduced_lqp_nc_oracle(q, f0, f1):
    """Reduced LQP + its non-contact half as reference tests/sp_test/jacc_compare.cpp:367-384, 485-487: L0 = COM position + pelvis
    rotation (contact chain), L1 = left hand 6D (non-contact chain; ts_[1] is what ConfigureLQP_R_NC hard-codes)."""
    r = orc.OracleRobot(URDF)
    r.add_contact(LEFT_FOOT, FOOT_POINT, (0, 0, 1), 0.13, 0.06)
    r.add_contact(RIGHT_FOOT, FOOT_POINT, (0, 0, 1), 0.13, 0.06)
    r.add_task(0, orc.TASK_LINK_POSITION, "COM")
    r.add_task(0, orc.TASK_LINK_ROTATION, "pelvis_link")
    r.add_task(1, orc.TASK_LINK_6D, "L_Wrist2_Link")
    r.update_kinematics(q, np.zeros(39))
    r.set_contact(1, 1)
    r.set_task(0, f0)
    r.set_task(1, f1)
    r.reduced_dynamics()
    assert r.reduced_contact_constraint() == 1
    assert r.reduced_task_space() == 1
    assert r.configure_lqp(True) == 3 and r.calc_lqp() == 1
    r.configure_lqp_r_nc()
    ok = r.calc_lqp_r_nc()
    return r, ok


@pytest.mark.parametrize("hard", [False, True])
def test_reduced_lqp_nc_oracle(hard):
    """ConfigureLQP_R_NC + CalcControlTorqueLQP_R_NC (reference src/dwbc.cpp:4634-4760, 4479-4502): 21 joint accelerations of the
    non-contact chain; level 0 (solvefirst) = centroidal equality J_I_nc qdd = fstar_gnc + torque rows, level 1 = hand task in the
    pelvis-relative form + |qdd| <= 5.  KKT of both QPs, the hierarchy, the data of the levels and the stitched torque."""
    q, _, rng = synth_states(3, 20249)
    for b in range(3):
        f0 = np.concatenate([0.5 * rng.standard_normal(3), 1.0 * rng.standard_normal(3)])
        f1 = 1.0 * rng.standard_normal(6)
        if hard:
            f1[:3] *= 40.0  # the acceleration bound |qdd| <= 5 saturates
        r, ok = reduced_lqp_nc_oracle(q[b], f0, f1)
        assert ok == 1
        nc, acc = 21, 24
        yR = r.get("lqp_y2")[:, 0]
        JI, ANC, G = r.get("J_I_nc"), r.get("A_NC"), r.get("G")[:, 0]
        Ann, Gnc = ANC[6:, 6:], G[39 - nc:]
        # level data as the reference writes them
        B0, b0, A0, a0 = (r.get(f"lqpnc_{k}1") for k in ("B", "b", "A", "a"))
        assert relerr(B0, JI) < 1e-14 and relerr(b0[:, 0], -yR[acc - 6:acc]) < 1e-14
        assert relerr(A0, np.vstack([Ann, -Ann])) < 1e-14 and relerr(a0[:, 0], np.concatenate([-200 + Gnc, -200 - Gnc])) < 1e-14
        assert relerr(r.get("lqp_nc_H"), 5 * Ann / np.linalg.norm(Ann)) < 1e-14
        Jt = r.get("J_task1")
        p = r.get("link_pos")
        hand = 23
        d = p[:, hand] - p[:, 0]
        Ja = np.eye(6)
        Ja[:3, 3:] = np.array([[0, -d[2], d[1]], [d[2], 0, -d[0]], [-d[1], d[0], 0]])
        assert relerr(r.get("lqpnc_B2"), Jt[:, 39 - nc:]) < 1e-14
        assert relerr(r.get("lqpnc_b2")[:, 0], -Ja @ (f1 - yR[:6])) < 1e-12
        ys = [r.get(f"lqpnc_y{l}")[:, 0] for l in range(3)]
        assert not ys[0].any()
        for l in (1, 2):
            H, g = r.get(f"lqpnc_qpH{l}"), r.get(f"lqpnc_qpg{l}")[:, 0]
            QA, ub = r.get(f"lqpnc_qpA{l}"), r.get(f"lqpnc_qpub{l}")[:, 0]
            Zp = r.get(f"lqpnc_Z{l - 1}")
            nu = Zp.shape[1]
            assert nu == (21 if l == 1 else 15)
            z = np.concatenate([np.linalg.lstsq(Zp, ys[l] - ys[l - 1], rcond=None)[0], r.get(f"lqpnc_v{l}")[:, 0]])
            res, act = kkt_residual(H, g, QA, ub, z, nu)
            assert res < 1e-8 * max(1.0, np.linalg.norm(g)), (l, res)
        # hierarchy: level 1 works in the null space of the centroidal equality and respects the frozen torque-row slack
        assert np.linalg.norm(B0 @ (ys[2] - ys[1])) < 1e-9
        v0 = r.get("lqpnc_v1")[:, 0]
        assert (A0 @ ys[2] + a0[:, 0] - v0).max() < 1e-7
        if not hard:
            # (the cost 5 A_NC / |A_NC| shares level 0 with its equality rows -- HQP::solvefirst adds H_ to the Hessian, src/dwbc_hqp.cpp:264 --
            # so the centroidal equality is met in the least-squares sense only)
            assert np.abs(ys[2]).max() <= 5 + 1e-6
        else:
            # the bound is a soft row (slack with unit cost against the task residual): it bends, and the slack says by how much
            v2 = r.get("lqpnc_v2")[:, 0]
            assert v2.max() > 1e-3 and abs((np.abs(ys[2]).max() - 5.0) - v2.max()) < 1e-6
        # stitched torque: contact-chain joints from the reduced LQP, the rest from the non-contact dynamics
        tau = r.get("lqp_stitched_torque")[:, 0]
        assert relerr(tau[:12], r.get("lqp_torque")[:12, 0]) < 1e-14
        assert relerr(tau[12:], Ann @ ys[2] + Gnc) < 1e-12
