"""generic-DoF models -- the reference's own scaling-study family (reference tests/dof_test/
dyros_tocabi_dof{18..45}.urdf, tests/sp_test/dof_comparison.cpp:41-146; three of them reduced to their dynamics data under
tests/golden/dof/).  Kernel logic (emulation build) vs the oracle on 18-, 25- and 45-joint models: gravity compensation,
contact-consistent dynamics, closed-form task torque and the 2-level HQP + redistribution with torque limit 200."""
import os

import numpy as np
import pytest

from libdwbc_b200 import CONTACT_6D, SOLVE_GRAV, SOLVE_HQP, SOLVE_REDIST, SOLVE_TASK, TASK_LINK_POSITION, TASK_LINK_ROTATION, RobotDataBatch
from oracle import oracle as orc
from tests.common import *


def dof_urdf(dof):
    return os.path.join(GOLDEN, "dof", f"tocabi_dof{dof}.urdf")


def variant_urdf(name):
    return os.path.join(GOLDEN, "variants", f"tocabi_{name}.urdf")


def dof_states(dof, B, seed, joints=None):
    """the reference's q2 (tests/sp_test/dof_comparison.cpp:74-81) truncated / zero-extended to the model, plus the section 8(d) noise"""
    rng = np.random.Generator(np.random.PCG64(seed))
    base = np.array(Q_CASE[1][6:39])
    if joints is None:
        joints = np.zeros(dof)
        joints[:min(dof, 33)] = base[:min(dof, 33)]
    q = np.zeros((B, dof + 7))
    q[:, 6:6 + dof] = joints + 0.05 * rng.uniform(-1, 1, (B, dof))
    q[:, 0:3] = 0.01 * rng.uniform(-1, 1, (B, 3))
    rpy = 0.1 * rng.uniform(-1, 1, (B, 3))
    from scipy.spatial.transform import Rotation
    qu = Rotation.from_euler("xyz", rpy).as_quat()  # x y z w
    q[:, 3:6] = qu[:, :3]
    q[:, dof + 6] = qu[:, 3]
    qd = 0.1 * rng.uniform(-1, 1, (B, dof + 6))
    return q, qd, rng


def run_generic(dof, make_rd, D, H, B=3, urdf=None, joints=None):
    urdf = urdf or dof_urdf(dof)
    q, qd, rng = dof_states(dof, B, 20260 + dof, joints)
    fs = [0.3 * rng.standard_normal((B, 3)), 0.3 * rng.standard_normal((B, 3))]
    rd = make_rd(urdf, B)
    assert (rd.model_dof_, rd.system_dof_) == (dof, dof + 6)
    rd.add_contact_constraint(LEFT_FOOT, CONTACT_6D, FOOT_POINT, (0, 0, 1), 0.15, 0.075)
    rd.add_contact_constraint(RIGHT_FOOT, CONTACT_6D, FOOT_POINT, (0, 0, 1), 0.15, 0.075)
    rd.add_task_space(0, TASK_LINK_POSITION, "COM")
    rd.add_task_space(1, TASK_LINK_ROTATION, "pelvis_link")
    rd.set_torque_limit(200.0)
    rd.update_kinematics(D(q), D(qd))
    flags = np.ones((B, 2), dtype=np.uint8)
    flags[B - 1] = [1, 0]
    rd.set_contact(D(flags))
    rd.calc_contact_constraint()
    for l, f in enumerate(fs):
        rd.set_task_space(l, D(f))
    closed = [H(x) for x in rd.solve(SOLVE_GRAV | SOLVE_TASK)]
    hqp = [H(x) for x in rd.solve(SOLVE_GRAV | SOLVE_HQP | SOLVE_REDIST)]
    for b in range(B):
        ref = orc.OracleRobot(urdf)
        ref.add_contact(LEFT_FOOT, FOOT_POINT, (0, 0, 1), 0.15, 0.075)
        ref.add_contact(RIGHT_FOOT, FOOT_POINT, (0, 0, 1), 0.15, 0.075)
        ref.add_task(0, orc.TASK_LINK_POSITION, "COM")
        ref.add_task(1, orc.TASK_LINK_ROTATION, "pelvis_link")
        ref.set_torque_limit(200.0)
        ref.update_kinematics(q[b], qd[b])
        ref.set_contact(*flags[b])
        assert ref.calc_contact_constraint() == 1
        for l, f in enumerate(fs):
            ref.set_task(l, f[b])
        ref.calc_grav()
        ref.update_task_space()
        for name in ("A", "A_inv", "Lambda_contact", "W_inv"):
            a = H(rd.get(name))[b]
            g = ref.get(name)
            assert relerr(a[:g.shape[0], :g.shape[1]], g) < 1e-9, (dof, b, name)
        ref.calc_task_torque(False)
        assert relerr(closed[0][b], ref.get("torque_grav")[:, 0]) < 1e-9 and relerr(closed[1][b], ref.get("torque_task")[:, 0]) < 1e-8, (dof, b)
        ok = ref.calc_task_torque(True)
        ok2 = ref.calc_redistribute(True)
        assert (hqp[3][b] & 7) == (1 | (2 if ok else 0) | (4 if ok2 else 0)), (dof, b, hqp[3][b], ok, ok2)
        if ok:
            assert relerr(hqp[1][b], ref.get("torque_task")[:, 0]) < 1e-6, (dof, b)
        if ok and ok2:
            tc = ref.get("torque_contact")[:, 0]
            assert np.abs(hqp[2][b] - tc).max() < 1e-6 * max(1.0, np.abs(tc).max()), (dof, b)


@pytest.mark.parametrize("dof", [18, 25, 45])
def test_emul_generic_dof(dof):
    from tests.emul.emul_backend import EmulBackend

    run_generic(dof, lambda urdf, B: RobotDataBatch(urdf, B, _backend=EmulBackend()), np.ascontiguousarray, lambda a: a)


# ---- model variants (reference tests/dyros_tocabi_{wo_head,fixed_head,ub}.urdf) ----
HEAD_JOINT = 24  # index of Head_Joint among TOCABI's 33 joints: removed (wo_head) or fixed and merged into the neck link (fixed_head)


def run_variant(name, make_rd, D, H, B=3):
    """fixed_head (32 joints, head inertia merged into the neck link): the whole path on the reference's case-2 posture without the
    head joint.  wo_head (32 joints, head link deleted): stage a1 only -- its 0.1 kg neck link leaves A_ a 5.8e-5 diagonal entry, W
    singular values from 1.7e4 down to 8.6e-3, and the reference's COD threshold 1e-6 (include/dwbc_wbd.h:10) then cuts real directions
    of W ("Contact Space Factorization Error", src/wbd.cpp:130-135; the reference's only use of the model is commented out,
    tests/sp_test/test_mod.cpp:525).  The oracle reproduces that failure; the product's closed-form null space does not fail."""
    joints = np.delete(np.array(Q_CASE[1][6:39]), HEAD_JOINT)
    if name == "fixed_head":
        return run_generic(32, make_rd, D, H, B=B, urdf=variant_urdf(name), joints=joints)
    run_kinematics_only(variant_urdf(name), 32, joints, make_rd, D, H, B)
    ref = orc.OracleRobot(variant_urdf(name))
    ref.add_contact(LEFT_FOOT, FOOT_POINT, (0, 0, 1), 0.15, 0.075)
    ref.add_contact(RIGHT_FOOT, FOOT_POINT, (0, 0, 1), 0.15, 0.075)
    q, qd, _ = dof_states(32, 1, 20281, joints)
    ref.update_kinematics(q[0], qd[0])
    ref.set_contact(1, 1)
    assert ref.calc_contact_constraint() == 0  # the reference's rank test fails on this model


def run_kinematics_only(urdf, dof, joints, make_rd, D, H, B, names=("A", "A_inv", "G", "B", "com_pos")):
    q, qd, _ = dof_states(dof, B, 20281, joints)
    rd = make_rd(urdf, B)
    assert (rd.model_dof_, rd.system_dof_) == (dof, dof + 6)
    rd.update_kinematics(D(q), D(qd))
    for b in range(B):
        ref = orc.OracleRobot(urdf)
        ref.update_kinematics(q[b], qd[b])
        for name in names:
            a = np.asarray(H(rd.get(name))[b], dtype=float)
            g = ref.get(name)
            a = a.reshape(g.shape) if a.size == g.size else a[:g.shape[0], :g.shape[1]]
            assert relerr(a, g) < (1e-7 if name == "A_inv" else 1e-9), (b, name)


def run_upper_body(make_rd, D, H, B=3):
    """ub (pelvis + 21 upper-body joints, no legs): UpdateKinematics only, the posture of reference tests/sp_test/rot_test.cpp:54-58
    (A_ and the COM are what that program prints).  A_inv_ is left out: the model's pelvis is massless and the first waist joint is a
    yaw joint, so pelvis yaw and waist yaw are one motion and A_ is singular (smallest eigenvalue 4e-15) in the reference as well."""
    joints = np.array([0, 0, 0, 0.3, 0.3, 1.5, -1.27, -1, 0, -1, 0, 0, 0, -0.3, -0.3, -1.5, 1.27, 1, 0, 1, 0], dtype=float)
    run_kinematics_only(variant_urdf("ub"), 21, joints, make_rd, D, H, B, names=("A", "G", "B", "com_pos"))


@pytest.mark.parametrize("name", ["wo_head", "fixed_head"])
def test_emul_model_variant(name):
    from tests.emul.emul_backend import EmulBackend

    run_variant(name, lambda urdf, B: RobotDataBatch(urdf, B, _backend=EmulBackend()), np.ascontiguousarray, lambda a: a)


def test_emul_upper_body_model():
    from tests.emul.emul_backend import EmulBackend

    run_upper_body(lambda urdf, B: RobotDataBatch(urdf, B, _backend=EmulBackend()), np.ascontiguousarray, lambda a: a)


def test_fixed_head_merges_head_inertia():
    """fixing a joint must not lose mass: fixed_head has the full robot's mass, wo_head is lighter by the head link"""
    from oracle.urdf_model import load_urdf

    full, fixed, wo = (load_urdf(p).total_mass for p in (URDF, variant_urdf("fixed_head"), variant_urdf("wo_head")))
    assert abs(full - fixed) < 1e-12 and wo < full - 0.1
