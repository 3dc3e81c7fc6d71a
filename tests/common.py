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


def config_dofs(config):
    return {"golden": (6, 3), 1: (3, 3), 2: (3, 6, 21), "lqp": (6, 6)}[config]


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
