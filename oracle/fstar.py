"""TEST INFRASTRUCTURE (oracle): numpy restatement of the reference's f* generation -- TaskLink::GetFstarPosPD / GetFstarRotPD
(reference src/task.cpp:268-334), QuinticSpline and GetPhi (src/math.cpp:127-183, 272-287) and the per-LINK_MODE frame
selection of UpdateTaskSpace (src/dwbc.cpp:714-785).  Independent of the CUDA code: the spline solves the 3x3 system as the
reference does, the rotation interpolation goes through scipy's rotation vectors instead of quaternion slerp.
PARITY: no golden vector exists for this step in the reference; pinned by closed-form properties in tests/test_fstar.py."""
import numpy as np
from scipy.spatial.transform import Rotation


def quintic_spline(time, t0, tf, x0, v0, a0, xf, vf, af):
    """reference src/math.cpp:127-183"""
    if time < t0:
        return np.array([x0, v0, a0])
    if time > tf:
        return np.array([xf, vf, af])
    T = tf - t0
    a1, a2, a3 = x0, v0, a0 / 2.0
    M = np.array([[T**3, T**4, T**5], [3 * T**2, 4 * T**3, 5 * T**4], [6 * T, 12 * T**2, 20 * T**3]])
    r = np.array([xf - x0 - v0 * T - a0 * T**2 / 2.0, vf - v0 - a0 * T, af - a0])
    a4, a5, a6 = np.linalg.inv(M) @ r
    s = time - t0
    return np.array([a1 + a2 * s + a3 * s**2 + a4 * s**3 + a5 * s**4 + a6 * s**5,
                     a2 + 2 * a3 * s + 3 * a4 * s**2 + 4 * a5 * s**3 + 5 * a6 * s**4,
                     2 * a3 + 6 * a4 * s + 12 * a5 * s**2 + 20 * a6 * s**3])


def get_phi(cur, des):
    """reference src/math.cpp:272-287 (returns -phi with phi = -1/2 sum v_i x w_i)"""
    return 0.5 * sum(np.cross(cur[:, i], des[:, i]) for i in range(3))


def fstar_pos_pd(gains, now, traj, cur_pos, cur_vel):
    t0, t1, p0, v0, p1, v1 = traj
    q = np.array([quintic_spline(now, t0, t1, p0[j], v0[j], 0.0, p1[j], v1[j], 0.0) for j in range(3)])
    return gains["pos_a"] * q[:, 2] + gains["pos_p"] * (q[:, 0] - cur_pos) + gains["pos_d"] * (q[:, 1] - cur_vel)


def fstar_rot_pd(gains, now, traj, cur_rot, cur_w):
    t0, t1, R0, R1 = traj
    qs = quintic_spline(now, t0, t1, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0)
    rv = Rotation.from_matrix(R1 @ R0.T).as_rotvec()  # angle * axis of rq_desired * rq_init^{-1}
    rot_traj = Rotation.from_rotvec(qs[0] * rv).as_matrix() @ R0  # slerp(qs(0))
    w_traj = qs[1] * rv
    return gains["rot_p"] * get_phi(cur_rot, rot_traj) + gains["rot_d"] * (w_traj - cur_w)


def link_frame(mode, link, point, st):
    """(xpos, v, rotm, w) the reference hands to GetFstar*PD for this LINK_MODE (src/dwbc.cpp:714-785).  st: dict with
    link_pos [3, nb], link_rot [nb][3, 3], link_vel [6, nb], com_local [nb, 3], com = (pos, v, w), nb."""
    if link == st["nb"]:
        p, v, w = st["com"]
        return p, v, st["link_rot"][0], w
    R = st["link_rot"][link]
    p, v, w = st["link_pos"][:, link], st["link_vel"][:3, link], st["link_vel"][3:, link]
    if mode in (1, 4):  # *_COM_FRAME: xipos, vi (reference src/link.cpp:83, 95)
        r = R @ st["com_local"][link]
        return p + r, v + np.cross(w, r), R, w
    if mode in (2, 5):  # *_CUSTOM_FRAME: reference quirk -- w x task_point_ with the local point (src/dwbc.cpp:741, 768)
        return p + R @ point, v + np.cross(w, point), R, w
    return p, v, R, w
