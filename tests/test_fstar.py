"""f* generation (task PD on quintic / slerp trajectories) and the MuJoCo / float32 wire
format, kernel logic (emulation build) vs the numpy restatement in oracle/fstar.py, plus closed-form properties of the
restatement itself (the reference holds no golden vector for this step)."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation

from libdwbc_b200 import (SOLVE_GRAV, SOLVE_TASK, TASK_LINK_6D, TASK_LINK_6D_CUSTOM_FRAME, TASK_LINK_POSITION, TASK_LINK_POSITION_COM_FRAME,
                          TASK_LINK_ROTATION, RobotDataBatch)
from oracle import fstar as ofs
from oracle import oracle as orc
from oracle.urdf_model import load_urdf
from tests.common import *

TASKS = [  # (level, mode, link, point)
    (0, TASK_LINK_POSITION, "COM", (0, 0, 0)), (0, TASK_LINK_ROTATION, "pelvis_link", (0, 0, 0)),
    (1, TASK_LINK_6D, "L_Wrist2_Link", (0, 0, 0)),
    (2, TASK_LINK_6D_CUSTOM_FRAME, "upperbody_link", (0.05, -0.02, 0.3)),
    (3, TASK_LINK_POSITION_COM_FRAME, "R_Wrist2_Link", (0, 0, 0)),
]
GAINS = dict(pos_p=np.array([400.0, 380.0, 420.0]), pos_d=np.array([40.0, 38.0, 42.0]), pos_a=np.array([1.0, 0.9, 1.1]),
             rot_p=np.array([300.0, 310.0, 290.0]), rot_d=np.array([30.0, 31.0, 29.0]), rot_a=np.array([1.0, 1.0, 1.0]))


def test_quintic_and_rotation_properties():
    """Boundary conditions of QuinticSpline, clamping outside [t0, tf], and the slerp end points."""
    s0 = ofs.quintic_spline(1.0, 1.0, 3.0, 0.2, -0.1, 0.0, 0.9, 0.3, 0.0)
    s1 = ofs.quintic_spline(3.0, 1.0, 3.0, 0.2, -0.1, 0.0, 0.9, 0.3, 0.0)
    assert np.allclose(s0, [0.2, -0.1, 0.0], atol=1e-12) and np.allclose(s1, [0.9, 0.3, 0.0], atol=1e-10)
    assert np.allclose(ofs.quintic_spline(0.5, 1.0, 3.0, 0.2, -0.1, 0.0, 0.9, 0.3, 0.0), [0.2, -0.1, 0.0])
    assert np.allclose(ofs.quintic_spline(3.5, 1.0, 3.0, 0.2, -0.1, 0.0, 0.9, 0.3, 0.0), [0.9, 0.3, 0.0])
    h = 1e-6  # velocity / acceleration are the derivatives of the position polynomial
    a, b = ofs.quintic_spline(2.0 - h, 1.0, 3.0, 0.2, -0.1, 0.0, 0.9, 0.3, 0.0), ofs.quintic_spline(2.0 + h, 1.0, 3.0, 0.2, -0.1, 0.0, 0.9, 0.3, 0.0)
    mid = ofs.quintic_spline(2.0, 1.0, 3.0, 0.2, -0.1, 0.0, 0.9, 0.3, 0.0)
    assert abs((b[0] - a[0]) / (2 * h) - mid[1]) < 1e-7 and abs((b[1] - a[1]) / (2 * h) - mid[2]) < 1e-7
    R0, R1 = Rotation.from_euler("zyx", [0.3, -0.2, 0.1]).as_matrix(), Rotation.from_euler("zyx", [-0.4, 0.5, 0.2]).as_matrix()
    g = dict(GAINS, rot_d=np.zeros(3))
    assert np.allclose(ofs.fstar_rot_pd(g, 0.0, (0.0, 1.0, R0, R1), R0, np.zeros(3)), 0, atol=1e-12)  # on the trajectory: no error
    assert np.allclose(ofs.fstar_rot_pd(g, 1.0, (0.0, 1.0, R0, R1), R1, np.zeros(3)), 0, atol=1e-12)
    assert np.allclose(ofs.get_phi(np.eye(3), Rotation.from_rotvec([0, 0, 1e-4]).as_matrix()), [0, 0, 1e-4], atol=1e-10)  # small-angle error


def _states(B, seed):
    q, qd, rng = synth_states(B, seed)
    return q, qd, rng


def _oracle_state(ref, model, q, qd):
    ref.update_kinematics(q, qd)
    nb = model.nb
    A = ref.get("A")
    lr = ref.get("link_rot")
    rots = [lr[:, i].reshape(3, 3).T for i in range(nb)]  # column-major 3x3 per column
    M = model.total_mass
    cpos = ref.get("com_pos")[:, 0]
    c = cpos - q[0:3]
    R0 = rots[0]
    lin = A[0:3, :] @ qd / M
    # angular rows of jac_ = SI_body^{-1} CMM_  (reference src/dwbc.cpp:331-356)
    sk = lambda v: np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
    Icom = R0 @ A[3:6, 3:6] @ R0.T - M * sk(c) @ sk(c).T
    h = sk(c).T @ (A[0:3, :] @ qd) + R0 @ (A[3:6, :] @ qd)
    return dict(nb=nb, link_pos=ref.get("link_pos"), link_vel=ref.get("link_vel"), link_rot=rots, com_local=model.com, com=(cpos, lin, np.linalg.solve(Icom, h)))


def test_emul_fstar_and_wire_format():
    from tests.emul.emul_backend import EmulBackend

    run_fstar_case(lambda B: RobotDataBatch(URDF, B, _backend=EmulBackend()), np.ascontiguousarray, lambda a: a)


def run_fstar_case(make_rd, D, H):
    """make_rd(B) -> RobotDataBatch; D: host array -> backend array; H: backend array -> numpy."""
    B = 5
    q, qd, rng = _states(B, 20251)
    rd = make_rd(B)
    rd.add_contact_constraint(LEFT_FOOT, 0, FOOT_POINT, (0, 0, 1), 0.15, 0.075)
    rd.add_contact_constraint(RIGHT_FOOT, 0, FOOT_POINT, (0, 0, 1), 0.15, 0.075)
    for lvl, mode, link, pt in TASKS:
        rd.add_task_space(lvl, mode, link, pt)
    # MuJoCo layout in, RBDL layout inside (reference src/pybind/rl_wbc_bridge.cpp:88-98)
    qpos = np.concatenate([q[:, 0:3], q[:, 39:40], q[:, 3:6], q[:, 6:39]], axis=1)
    rd.update_kinematics_mujoco(D(qpos), D(qd))
    rd2 = make_rd(B)
    rd2.update_kinematics(D(q), D(qd))
    assert np.array_equal(H(rd.get("A")), H(rd2.get("A"))) and np.array_equal(H(rd.get("B")), H(rd2.get("B")))
    # trajectories: per state, per task link
    now = np.ascontiguousarray(0.2 + 1.2 * rng.uniform(0, 1, B))  # some states before / inside / after the window
    now[0], now[1] = 0.05, 1.9
    traj = {}
    index_in_level = {}
    for k, (lvl, mode, link, pt) in enumerate(TASKS):
        li = index_in_level.get(lvl, 0)
        index_in_level[lvl] = li + 1
        time = np.ascontiguousarray(np.tile([0.1, 1.5], (B, 1)) + 0.05 * rng.uniform(-1, 1, (B, 2)))
        g = GAINS
        rd.set_task_gain(lvl, li, g["pos_p"], g["pos_d"], g["pos_a"], g["rot_p"], g["rot_d"], g["rot_a"])
        ent = dict(time=time)
        if mode <= 5:
            ent["p"] = [np.ascontiguousarray(rng.uniform(-1, 1, (B, 3)) * s) for s in (0.5, 0.2, 0.5, 0.2)]
            rd.set_trajectory_quintic(lvl, li, D(time), *[D(x) for x in ent["p"]])
        if mode <= 2 or mode >= 6:
            Ri = Rotation.from_rotvec(0.4 * rng.uniform(-1, 1, (B, 3))).as_matrix()
            Rd = Rotation.from_rotvec(1.2 * rng.uniform(-1, 1, (B, 3))).as_matrix()
            ent["r"] = [np.ascontiguousarray(Ri), np.ascontiguousarray(Rd)]
            rd.set_trajectory_rotation(lvl, li, D(time), *[D(x) for x in ent["r"]])
        traj[k] = ent
    dofs = [6, 6, 6, 3]
    nowd = D(now)
    fsd = [rd.calc_fstar(l, nowd, D(np.full((B, dofs[l]), 7.0))) for l in range(4)]
    fs = [H(f) for f in fsd]
    # oracle
    model = load_urdf(URDF)
    ref = orc.OracleRobot(URDF)
    for b in range(B):
        st = _oracle_state(ref, model, q[b], qd[b])
        row = {0: 0, 1: 0, 2: 0, 3: 0}
        for k, (lvl, mode, link, pt) in enumerate(TASKS):
            lid = model.link_id(link)
            p, v, R, w = ofs.link_frame(mode, lid, np.array(pt, dtype=float), st)
            ent = traj[k]
            t0, t1 = ent["time"][b]
            r = row[lvl]
            if mode <= 5:
                want = ofs.fstar_pos_pd(GAINS, now[b], (t0, t1, *[x[b] for x in ent["p"]]), p, v)
                assert np.allclose(fs[lvl][b, r:r + 3], want, rtol=1e-9, atol=1e-9), (b, k)
                r += 3
            if mode <= 2 or mode >= 6:
                want = ofs.fstar_rot_pd(GAINS, now[b], (t0, t1, ent["r"][0][b], ent["r"][1][b]), R, w)
                assert np.allclose(fs[lvl][b, r:r + 3], want, rtol=1e-9, atol=1e-9), (b, k)
                r += 3
            row[lvl] = r
    # float32 torque command = grav + task + contact of the last solve (getTorqueCommand)
    flags = np.ones((B, 2), dtype=np.uint8)
    rd.set_contact(D(flags))
    rd.calc_contact_constraint()
    for l in range(4):
        rd.set_task_space(l, fsd[l])
    tg, tt, tc, stt = [H(x) for x in rd.solve(SOLVE_GRAV | SOLVE_TASK)]
    cmd = H(rd.get_torque_command())
    assert cmd.dtype == np.float32 and np.array_equal(cmd, (tg + tt + tc).astype(np.float32))
