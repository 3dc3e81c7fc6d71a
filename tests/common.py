"""Shared test inputs: the reference's own test configurations (tests/dwbc_test.cpp:48-77,152-179)
and the synthetic state generator of SURVEY.md section 8(d)."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
URDF = os.path.join(GOLDEN, "tocabi_model.urdf")

# reference tests/dwbc_test.cpp:48-54 and :152-158
Q_CASE = {
    1: [0, 0, 0.92983, 0, 0, 0, 0.0, 0.0, -0.24, 0.6, -0.36, 0.0, 0.0, 0.0, -0.24, 0.6, -0.36, 0.0, 0, 0, 0,
        0.3, 0.3, 1.5, -1.27, -1, 0, -1, 0, 0, 0, -0.3, -0.3, -1.5, 1.27, 1, 0, 1, 0, 1.0],
    2: [0, 0, 0.92983, 0, 0, 0, 0.1, 0.0, -0.24, 0.5, -0.6, 0.0, 0.05, 0.0, -0.21, 0.7, -0.31, 0.0, 0, 0, 0,
        0.2, 0.5, 1.5, -1.27, -1.2, 0, -1, 0, 0, 0, -0.3, -0.3, -1.5, 1.27, 1.3, 0.1, 1.3, 0, 1.0],
}
# reference tests/dwbc_test.cpp:56-57 and :160-161
FSTAR_CASE = {1: ([0.1, 4.0, 0.1, 0.1, -0.1, 0.1], [0.1, -0.1, 0.1]), 2: ([0.4, 2.0, 0.1, 0.3, -0.1, 0.1], [0.1, 0.1, 0.1])}
# reference tests/sp_test/data_confirmation.cpp:41-57: a state and f* captured from a walking controller (four registered contacts, feet
# active; COM / pelvis rotation / upper-body rotation / both hands 6D; no torque limit)
Q_WALKING = [-0.0325, -0.0579, 0.7273, 0.0194, -0.0118, -0.0008, -0.0006, 0.0698, -0.7835, 1.6487, -0.8420, -0.0911,
             -0.0007, 0.0767, -0.7963, 1.6742, -0.8549, -0.1150, -0.0001, -0.0003, 0.0204,
             0.2998, 0.3001, 1.5000, -1.2701, -1.0507, 0.0000, -1.0000, 0.0000,
             -0.0000, 0.0003, -0.2998, -0.3060, -1.5001, 1.2700, 1.0848, 0.0000, 1.0000, 0.0000, 0.9997]
FSTAR_WALKING = ([0.3142, -1.8202, -1.7750], [-17.8677, 8.4977, 1.0850], [-8.5340, 8.5992, 1.2655],
                 [4.0251, 3.9975, 7.5672, -8.2841, 30.3652, 0.8954, 2.7585, 3.7898, 9.3234, -9.5724, 43.8036, 2.5202])
FOOT_POINT = (0.03, 0.0, -0.1585)  # reference tests/dwbc_test.cpp:66-67
LEFT_FOOT, RIGHT_FOOT, PELVIS, UPPERBODY = 6, 12, 0, 15

# golden file -> oracle/product matrix name
GOLDEN_NAMES = [
    ("A", "Acontact_mat"), ("A_inv", "A_inv_"), ("J_C", "J_C"), ("Lambda_contact", "Lambda_contact"),
    ("J_C_INV_T", "J_C_INV_T"), ("N_C", "N_C"), ("W", "W"), ("W_inv", "W_inv"), ("NwJw", "NwJw"),
    ("torque_grav", "torque_grav_"),
]


def synth_states(B, seed, q0=None):
    """SURVEY.md section 8(d) 'Synthetic inputs': PCG64(seed); joints q0+0.05U, base xyz +0.01U, base rpy 0.1U,
    qdot 0.1U."""
    rng = np.random.Generator(np.random.PCG64(seed))
    q0 = np.asarray(Q_CASE[1] if q0 is None else q0, dtype=np.float64)
    q = np.tile(q0, (B, 1))
    q[:, 6:39] += 0.05 * rng.uniform(-1, 1, (B, 33))
    q[:, 0:3] = np.array([0, 0, 0.92983]) + 0.01 * rng.uniform(-1, 1, (B, 3))
    rpy = 0.1 * rng.uniform(-1, 1, (B, 3))
    cr, sr = np.cos(rpy[:, 0] / 2), np.sin(rpy[:, 0] / 2)
    cp, sp = np.cos(rpy[:, 1] / 2), np.sin(rpy[:, 1] / 2)
    cy, sy = np.cos(rpy[:, 2] / 2), np.sin(rpy[:, 2] / 2)
    q[:, 39] = cr * cp * cy + sr * sp * sy
    q[:, 3] = sr * cp * cy - cr * sp * sy
    q[:, 4] = cr * sp * cy + sr * cp * sy
    q[:, 5] = cr * cp * sy - sr * sp * cy
    qd = 0.1 * rng.uniform(-1, 1, (B, 39))
    return q, qd, rng


def synth_fstar(rng, B, dofs=(3, 6, 21), scales=(0.5, 1.0, 2.0), hard_frac=0.05):
    fs = [s * rng.standard_normal((B, d)) for d, s in zip(dofs, scales)]
    k = int(round(hard_frac * B))
    if k > 0:
        idx = rng.choice(B, size=k, replace=False)
        fs[0][idx, 1] = 4.0  # infeasible lateral acceleration, reference tests/dwbc_test.cpp:56
    return fs


def posture_selector(n=39, first_joint=12, count=21):
    """TASK_CUSTOM Jacobian [0 | S] selecting the non-leg joints."""
    J = np.zeros((count, n))
    for i in range(count):
        J[i, 6 + first_joint + i] = 1.0
    return J


# ---------------------------------------------------------------------------------------------
# Configurations of BASELINE.json / SURVEY.md section 8(d), applied identically to the product
# (libdwbc_b200.RobotDataBatch) and to the oracle (oracle.oracle.OracleRobot).
# ---------------------------------------------------------------------------------------------
def configure_product(rd, config, ncontacts=2):
    from libdwbc_b200 import CONTACT_6D, TASK_LINK_6D, TASK_LINK_POSITION, TASK_LINK_ROTATION
    if config == "walking":  # reference tests/sp_test/data_confirmation.cpp:59-72
        for link, lx, ly in ((LEFT_FOOT, 0.12, 0.06), (RIGHT_FOOT, 0.12, 0.06), (23, 0.04, 0.04), (31, 0.04, 0.04)):
            rd.add_contact_constraint(link, CONTACT_6D, FOOT_POINT, (0, 0, 1), lx, ly)
        rd.add_task_space(0, TASK_LINK_POSITION, "COM")
        rd.add_task_space(1, TASK_LINK_ROTATION, "pelvis_link")
        rd.add_task_space(2, TASK_LINK_ROTATION, "upperbody_link")
        rd.add_task_space(3, TASK_LINK_6D, "L_Wrist2_Link")
        rd.add_task_space(3, TASK_LINK_6D, "R_Wrist2_Link")
        return rd
    rd.add_contact_constraint(LEFT_FOOT, CONTACT_6D, FOOT_POINT, (0, 0, 1), 0.15, 0.075)
    rd.add_contact_constraint(RIGHT_FOOT, CONTACT_6D, FOOT_POINT, (0, 0, 1), 0.15, 0.075)
    if ncontacts == 4:
        rd.add_contact_constraint(23, CONTACT_6D, FOOT_POINT, (0, 0, 1), 0.04, 0.04)
        rd.add_contact_constraint(31, CONTACT_6D, FOOT_POINT, (0, 0, 1), 0.04, 0.04)
    if config == "golden":  # reference tests/dwbc_test.cpp:71-77
        rd.add_task_space(0, TASK_LINK_6D, PELVIS)
        rd.add_task_space(1, TASK_LINK_ROTATION, UPPERBODY)
        rd.set_torque_limit(300.0)
    elif config == 1:  # COM position + pelvis rotation, closed form
        rd.add_task_space(0, TASK_LINK_POSITION, "COM")
        rd.add_task_space(1, TASK_LINK_ROTATION, "pelvis_link")
    elif config == 2:  # 3-level HQP: COM / pelvis+upper-body rotation / posture
        rd.add_task_space(0, TASK_LINK_POSITION, "COM")
        rd.add_task_space(1, TASK_LINK_ROTATION, "pelvis_link")
        rd.add_task_space(1, TASK_LINK_ROTATION, "upperbody_link")
        rd.add_custom_task_space(2, 21)
        rd.set_torque_limit(300.0)
    elif config == "lqp":  # reference tests/sp_test/jacc_compare.cpp:319-334: COM position + pelvis rotation / left hand 6D
        rd.add_task_space(0, TASK_LINK_POSITION, "COM")
        rd.add_task_space(0, TASK_LINK_ROTATION, "pelvis_link")
        rd.add_task_space(1, TASK_LINK_6D, "L_Wrist2_Link")
    return rd


def configure_oracle(ref, config, ncontacts=2):
    from oracle import oracle as orc
    if config == "walking":
        for link, lx, ly in ((LEFT_FOOT, 0.12, 0.06), (RIGHT_FOOT, 0.12, 0.06), (23, 0.04, 0.04), (31, 0.04, 0.04)):
            ref.add_contact(link, FOOT_POINT, (0, 0, 1), lx, ly)
        ref.add_task(0, orc.TASK_LINK_POSITION, "COM")
        ref.add_task(1, orc.TASK_LINK_ROTATION, "pelvis_link")
        ref.add_task(2, orc.TASK_LINK_ROTATION, "upperbody_link")
        ref.add_task(3, orc.TASK_LINK_6D, "L_Wrist2_Link")
        ref.add_task(3, orc.TASK_LINK_6D, "R_Wrist2_Link")
        return ref
    ref.add_contact(LEFT_FOOT, FOOT_POINT, (0, 0, 1), 0.15, 0.075)
    ref.add_contact(RIGHT_FOOT, FOOT_POINT, (0, 0, 1), 0.15, 0.075)
    if ncontacts == 4:
        ref.add_contact(23, FOOT_POINT, (0, 0, 1), 0.04, 0.04)
        ref.add_contact(31, FOOT_POINT, (0, 0, 1), 0.04, 0.04)
    if config == "golden":
        ref.add_task(0, orc.TASK_LINK_6D, PELVIS)
        ref.add_task(1, orc.TASK_LINK_ROTATION, UPPERBODY)
        ref.set_torque_limit(300.0)
    elif config == 1:
        ref.add_task(0, orc.TASK_LINK_POSITION, "COM")
        ref.add_task(1, orc.TASK_LINK_ROTATION, "pelvis_link")
    elif config == 2:
        ref.add_task(0, orc.TASK_LINK_POSITION, "COM")
        ref.add_task(1, orc.TASK_LINK_ROTATION, "pelvis_link")
        ref.add_task(1, orc.TASK_LINK_ROTATION, "upperbody_link")
        ref.add_custom_task(2, 21)
        ref.set_torque_limit(300.0)
        ref.set_task(2, np.zeros(21), posture_selector())
    elif config == "lqp":
        ref.add_task(0, orc.TASK_LINK_POSITION, "COM")
        ref.add_task(0, orc.TASK_LINK_ROTATION, "pelvis_link")
        ref.add_task(1, orc.TASK_LINK_6D, "L_Wrist2_Link")
    return ref


def walking_inputs(B=1):
    """The captured walking state of reference tests/sp_test/data_confirmation.cpp:41-57 (B copies): q with the base quaternion
    normalised (the capture is rounded to four decimals; RBDL expects a unit quaternion, and with a non-unit one the rotation matrix
    is not orthogonal -- the oracle, which follows RBDL's body-frame recursions, and the kernels, which work in one common frame, then
    differ by O(1 - |quat|) = 3e-7), zero velocity, the four f* vectors, contact flags feet on / hands off."""
    q = np.array([Q_WALKING] * B, dtype=np.float64)
    q[:, [3, 4, 5, 39]] /= np.linalg.norm(q[0, [3, 4, 5, 39]])
    fs = [np.array([f] * B, dtype=np.float64) for f in FSTAR_WALKING]
    flags = np.array([[1, 1, 0, 0]] * B, dtype=np.uint8)
    return q, np.zeros((B, 39)), fs, flags


def check_walking_state(rd, out, to_np):
    """Product results for walking_inputs() against the oracle: torques 1e-9 (the QPs are active: the right foot is unloaded to zero
    wrench, a vertex where all ten rows of that contact and one row of the left foot hold with equality), identical active sets."""
    from oracle import oracle as orc

    q, qd, fs, flags = walking_inputs(1)
    o = configure_oracle(orc.OracleRobot(URDF), "walking").run_batch(q, qd, fs, flags, sets=True)
    tg, tt, tc, st = [to_np(x) for x in out]
    assert int(st[0]) & 7 == 7 and int(o["status"][0]) & 7 == 7
    assert relerr(tg[0], o["torque_grav"][0]) < 1e-9
    assert relerr(tt[0], o["torque_task"][0]) < 1e-9 and relerr(tc[0], o["torque_contact"][0]) < 1e-9
    assert np.abs(tt[0]).max() > 10 and np.abs(tc[0]).max() > 10
    for l in range(4):
        rows = [i for i in range(128) if (int(o["active"][0, l, i // 32]) >> (i % 32)) & 1]
        got = list(rd.active_set(l)[0])
        assert got == rows and len(rows) >= 10, (l, got, rows)


def config_dofs(config):
    return {"golden": (6, 3), 1: (3, 3), 2: (3, 6, 21), "lqp": (6, 6), "walking": (3, 3, 3, 12)}[config]


def lqp_fstars(rng, B, hard=False):
    """f* of the LQP cases: level 0 = [COM acceleration; pelvis angular acceleration], level 1 = left-hand 6D.  `hard`: a lateral
    COM acceleration far beyond the friction cone and a large hand acceleration (contact / acceleration rows saturate)."""
    f0 = np.concatenate([0.5 * rng.standard_normal((B, 3)), 1.0 * rng.standard_normal((B, 3))], axis=1)
    f1 = 1.0 * rng.standard_normal((B, 6))
    if hard:
        f0[:, 1] = 12.0
        f1[:, :3] *= 30.0
    return [f0, f1]


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
