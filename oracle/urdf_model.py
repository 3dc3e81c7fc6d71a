"""URDF -> flat rigid-body model arrays (TEST INFRASTRUCTURE: part of the CPU oracle).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.  The product has its own, independent URDF reader in
libdwbc_b200/csrc/urdf_model.cpp.

What is restated here is the model construction that the reference delegates to RBDL
(`RigidBodyDynamics::Addons::URDFReadFromFile`, called at reference src/dwbc.cpp:115 with
floating_base=true) followed by `RobotData::InitModelData` (reference src/dwbc.cpp:125-236):

* joints are visited depth first, the children of a link sorted by *joint name*
  (confirmed by the link ids used in reference tests/dwbc_test.cpp:63-72),
* URDF `rpy` means R = Rz(yaw) * Ry(pitch) * Rx(roll),
* links attached through `fixed` joints are merged into their parent body,
* the floating base is a 3-DoF translation followed by a spherical joint, so
  q = [p(3) | quat xyz(3) | theta(m) | quat w], qdot = [v_world(3) | omega_body(3) | theta_dot(m)]
  (reference src/pybind/rl_wbc_bridge.cpp:88-98),
* link i (reference `link_[i]`) = i-th movable body with non-zero mass, link 0 = base body,
  and a pseudo link "COM" is appended last (reference src/dwbc.cpp:158-162, 230-231).
"""
from __future__ import annotations

import xml.etree.ElementTree as ET
from dataclasses import dataclass, field

import numpy as np


def rpy_to_rot(rpy):
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]], dtype=float)


def _vec(s, n=3):
    if s is None:
        return np.zeros(n)
    return np.array([float(x) for x in s.split()], dtype=float)


@dataclass
class Model:
    """Flat arrays for a floating-base tree.  Body 0 is the base; body i>=1 is the child of the
    revolute joint whose velocity column is 5+i."""

    names: list = field(default_factory=list)  # body names (URDF link names)
    parent: np.ndarray = None  # [nb] int32, parent body (-1 for base)
    E_T: np.ndarray = None  # [nb,3,3] rotation parent-frame -> joint-frame coords is E_T^T; here R_parent<-joint
    r_T: np.ndarray = None  # [nb,3] joint origin in the parent body frame
    axis: np.ndarray = None  # [nb,3] joint axis in the joint (= child body) frame
    mass: np.ndarray = None  # [nb]
    com: np.ndarray = None  # [nb,3] centre of mass in the body frame
    inertia: np.ndarray = None  # [nb,3,3] rotational inertia about the COM, body frame

    @property
    def nb(self):
        return len(self.names)

    @property
    def model_dof(self):
        return self.nb - 1

    @property
    def system_dof(self):
        return self.nb + 5

    @property
    def total_mass(self):
        return float(self.mass.sum())

    def link_id(self, name):
        low = [n.lower() for n in self.names]
        if name.lower() == "com":
            return self.nb
        return low.index(name.lower())

    def flat(self):
        """Contiguous float64 / int32 arrays in the layout oracle/dwbc_oracle.cpp expects."""
        return dict(
            nb=np.int32(self.nb),
            parent=np.ascontiguousarray(self.parent, dtype=np.int32),
            E_T=np.ascontiguousarray(self.E_T, dtype=np.float64),
            r_T=np.ascontiguousarray(self.r_T, dtype=np.float64),
            axis=np.ascontiguousarray(self.axis, dtype=np.float64),
            mass=np.ascontiguousarray(self.mass, dtype=np.float64),
            com=np.ascontiguousarray(self.com, dtype=np.float64),
            inertia=np.ascontiguousarray(self.inertia, dtype=np.float64),
        )


def load_urdf(path: str) -> Model:
    root = ET.parse(path).getroot()
    links = {}
    for ln in root.findall("link"):
        inert = ln.find("inertial")
        if inert is None:
            m, c, I, Ri = 0.0, np.zeros(3), np.zeros((3, 3)), np.eye(3)
        else:
            org = inert.find("origin")
            c = _vec(org.get("xyz")) if org is not None else np.zeros(3)
            Ri = rpy_to_rot(_vec(org.get("rpy"))) if org is not None and org.get("rpy") else np.eye(3)
            m = float(inert.find("mass").get("value"))
            it = inert.find("inertia")
            g = lambda k: float(it.get(k))
            I = np.array([[g("ixx"), g("ixy"), g("ixz")], [g("ixy"), g("iyy"), g("iyz")], [g("ixz"), g("iyz"), g("izz")]])
            I = Ri @ I @ Ri.T
        links[ln.get("name")] = dict(mass=m, com=c, inertia=I)

    children = {}
    child_links = set()
    for jn in root.findall("joint"):
        org = jn.find("origin")
        xyz = _vec(org.get("xyz")) if org is not None else np.zeros(3)
        rpy = _vec(org.get("rpy")) if org is not None and org.get("rpy") else np.zeros(3)
        ax = jn.find("axis")
        axis = _vec(ax.get("xyz")) if ax is not None else np.array([1.0, 0, 0])
        j = dict(name=jn.get("name"), type=jn.get("type"), parent=jn.find("parent").get("link"),
                 child=jn.find("child").get("link"), xyz=xyz, R=rpy_to_rot(rpy), axis=axis)
        children.setdefault(j["parent"], []).append(j)
        child_links.add(j["child"])
    roots = [n for n in links if n not in child_links]
    assert len(roots) == 1, roots
    for v in children.values():
        v.sort(key=lambda j: j["name"])

    bodies = []  # dicts: name,parent,E,r,axis,mass,com,inertia

    def merge_fixed(body, R_pc, r_pc, link):
        """Merge `link`, rigidly attached at (R_pc, r_pc) w.r.t. the body frame, into `body`."""
        m2 = link["mass"]
        if m2 <= 0:
            return
        c2 = r_pc + R_pc @ link["com"]
        I2 = R_pc @ link["inertia"] @ R_pc.T
        m1, c1, I1 = body["mass"], body["com"], body["inertia"]
        m = m1 + m2
        c = (m1 * c1 + m2 * c2) / m
        d1, d2 = c1 - c, c2 - c
        I = I1 + m1 * (np.dot(d1, d1) * np.eye(3) - np.outer(d1, d1)) + I2 + m2 * (np.dot(d2, d2) * np.eye(3) - np.outer(d2, d2))
        body["mass"], body["com"], body["inertia"] = m, c, I

    def visit(link_name, body_idx, R_acc, r_acc):
        # R_acc, r_acc: pose of `link_name`'s frame relative to body `body_idx` (identity unless reached via fixed joints)
        for j in children.get(link_name, []):
            R_j = R_acc @ j["R"]
            r_j = r_acc + R_acc @ j["xyz"]
            if j["type"] == "fixed":
                merge_fixed(bodies[body_idx], R_j, r_j, links[j["child"]])
                visit(j["child"], body_idx, R_j, r_j)
            elif j["type"] in ("revolute", "continuous"):
                L = links[j["child"]]
                bodies.append(dict(name=j["child"], parent=body_idx, E=R_j, r=r_j, axis=j["axis"].copy(),
                                   mass=L["mass"], com=L["com"].copy(), inertia=L["inertia"].copy()))
                visit(j["child"], len(bodies) - 1, np.eye(3), np.zeros(3))
            else:
                raise ValueError("unsupported joint type " + j["type"])

    L0 = links[roots[0]]
    bodies.append(dict(name=roots[0], parent=-1, E=np.eye(3), r=np.zeros(3), axis=np.zeros(3),
                       mass=L0["mass"], com=L0["com"].copy(), inertia=L0["inertia"].copy()))
    visit(roots[0], 0, np.eye(3), np.zeros(3))

    md = Model()
    md.names = [b["name"] for b in bodies]
    md.parent = np.array([b["parent"] for b in bodies], dtype=np.int32)
    md.E_T = np.stack([b["E"] for b in bodies])
    md.r_T = np.stack([b["r"] for b in bodies])
    md.axis = np.stack([b["axis"] for b in bodies])
    md.mass = np.array([b["mass"] for b in bodies])
    md.com = np.stack([b["com"] for b in bodies])
    md.inertia = np.stack([b["inertia"] for b in bodies])
    return md


# --------------------------------------------------------------------------------------------
# Small NumPy kinematics used to cross-check the C++ oracle (first-principles formulas, no
# spatial algebra): FK, geometric Jacobians, and the mass matrix as a sum over bodies.
# --------------------------------------------------------------------------------------------

def quat_to_rot(x, y, z, w):
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def axis_rot(axis, th):
    a = axis / np.linalg.norm(axis)
    K = skew(a)
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def fk(md: Model, q):
    """World pose of every body: R[i] (world<-body), p[i] (body origin, world)."""
    nb = md.nb
    R = np.zeros((nb, 3, 3))
    p = np.zeros((nb, 3))
    R[0] = quat_to_rot(q[3], q[4], q[5], q[md.system_dof])
    p[0] = q[0:3]
    for i in range(1, nb):
        l = md.parent[i]
        R[i] = R[l] @ md.E_T[i] @ axis_rot(md.axis[i], q[5 + i])
        p[i] = p[l] + R[l] @ md.r_T[i]
    return R, p


def point_jacobian(md: Model, R, p, body, point_local):
    """6 x n Jacobian, rows [linear; angular], of a point fixed in `body` (reference
    src/contact_constraint.cpp:56-62 / src/link.cpp:98-107 after the row swap)."""
    n = md.system_dof
    J = np.zeros((6, n))
    pw = p[body] + R[body] @ point_local
    J[0:3, 0:3] = np.eye(3)
    J[0:3, 3:6] = -skew(pw - p[0]) @ R[0]
    J[3:6, 3:6] = R[0]
    i = body
    while i > 0:
        a = R[i] @ md.axis[i]
        J[0:3, 5 + i] = np.cross(a, pw - p[i])
        J[3:6, 5 + i] = a
        i = md.parent[i]
    return J


def mass_matrix_sum(md: Model, R, p):
    n = md.system_dof
    A = np.zeros((n, n))
    for i in range(md.nb):
        J = point_jacobian(md, R, p, i, md.com[i])
        Iw = R[i] @ md.inertia[i] @ R[i].T
        A += md.mass[i] * J[0:3].T @ J[0:3] + J[3:6].T @ Iw @ J[3:6]
    return A


def read_golden(path):
    """Reference golden format (tests/dwbc_test_util.h:15-28): int64 rows, cols, col-major f64."""
    raw = np.fromfile(path, dtype=np.uint8)
    rows, cols = np.frombuffer(raw[:16], dtype=np.int64)
    data = np.frombuffer(raw[16:16 + 8 * rows * cols], dtype=np.float64)
    return data.reshape((cols, rows)).T.copy()
