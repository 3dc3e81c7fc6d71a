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


def dof_states(dof, B, seed):
    """the reference's q2 (tests/sp_test/dof_comparison.cpp:74-81) truncated / zero-extended to the model, plus the section 8(d) noise"""
    rng = np.random.Generator(np.random.PCG64(seed))
    base = np.array(Q_CASE[1][6:39])
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


def run_generic(dof, make_rd, D, H, B=3):
    urdf = dof_urdf(dof)
    q, qd, rng = dof_states(dof, B, 20260 + dof)
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
