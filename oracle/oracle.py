"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may import
this.  It mirrors the reference call order on one `RobotData` (reference include/dwbc.h:59-430):
UpdateKinematics -> SetContact -> CalcContactConstraint -> SetTaskSpace -> CalcGravCompensation
-> CalcTaskControlTorque -> CalcContactRedistribute.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .urdf_model import Model, load_urdf

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "build", "liboracle.so")

# LINK_MODE values, reference include/dwbc_task.h:23-33
TASK_LINK_6D, TASK_LINK_6D_COM_FRAME, TASK_LINK_6D_CUSTOM_FRAME = 0, 1, 2
TASK_LINK_POSITION, TASK_LINK_POSITION_COM_FRAME, TASK_LINK_POSITION_CUSTOM_FRAME = 3, 4, 5
TASK_LINK_ROTATION, TASK_LINK_ROTATION_CUSTOM_FRAME = 6, 7

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("dwbc_oracle.cpp", "linalg.hpp", "qp.hpp")]
    if force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _LIB


_lib = None
_NATIVE = os.path.join(os.path.dirname(_LIB), "liboracle_native.so")
_use_native = False
flavour = "portable (-march=x86-64-v3)"


def use_native():
    """bench.py's CPU arms: build the oracle with -march=native ON THIS MACHINE and time that one (falls back to the portable
    library if the compiler is missing).  Must be called before the first use of the library."""
    global _use_native, flavour
    try:
        subprocess.check_call(["make", "-s", "-C", _HERE, "native"])
        C.CDLL(_NATIVE)
        _use_native = True
        flavour = "built on this host with -O3 -march=native"
    except Exception:
        _use_native = False
    return _use_native


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_NATIVE if _use_native else _LIB)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, _ip, _dp, _dp, _dp, _dp, _dp, _dp]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_add_contact.argtypes = [C.c_void_p, C.c_int, _dp, _dp, C.c_double, C.c_double]
        L.orc_add_task.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _dp]
        L.orc_add_custom_task.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_set_torque_limit.argtypes = [C.c_void_p, _dp]
        L.orc_set_friction.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double]
        L.orc_update_kinematics.argtypes = [C.c_void_p, _dp, _dp]
        L.orc_set_contact.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_calc_contact_constraint.argtypes = [C.c_void_p]
        L.orc_set_task.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
        L.orc_calc_grav.argtypes = [C.c_void_p]
        L.orc_calc_task_torque.argtypes = [C.c_void_p, C.c_int]
        L.orc_calc_redistribute.argtypes = [C.c_void_p, C.c_int]
        L.orc_get.argtypes = [C.c_void_p, C.c_char_p, _dp, _ip, _ip]
        L.orc_get_active.argtypes = [C.c_void_p, C.c_int, _ip, C.c_int]
        L.orc_get_qp_info.argtypes = [C.c_void_p, C.c_int, _ip, _ip]
        L.orc_lexqp.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _ip]
        L.orc_cod_pinv.argtypes = [C.c_int, C.c_int, _dp, C.c_double, _dp, _ip]
        L.orc_run_batch.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp, C.c_char_p, C.c_int, _dp, _dp, _dp, _ip, C.c_int, _dp, _dp]
        L.orc_run_batch_sets.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp, C.c_char_p, C.c_int, _dp, _dp, _dp, _ip, C.c_int, C.POINTER(C.c_uint), C.POINTER(C.c_uint)]
        L.orc_contact_force.argtypes = [C.c_void_p, _dp, _dp]
        L.orc_zmp.argtypes = [C.c_void_p, _dp, _dp]
        L.orc_redis_eta.argtypes = [C.c_void_p]
        L.orc_redis_eta.restype = C.c_double
        L.orc_num_threads.restype = C.c_int
        for f in ("orc_reduced_dynamics", "orc_reduced_contact_constraint", "orc_reduced_grav", "orc_reduced_task_space"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_reduced_task_torque.argtypes = [C.c_void_p, C.c_int]
        L.orc_reduced_redistribute.argtypes = [C.c_void_p, C.c_int]
        L.orc_set_reduced_torque_limit.argtypes = [C.c_void_p, _dp, C.c_int]
        L.orc_reduced_dims.argtypes = [C.c_void_p, _ip]
        for f in ("orc_configure_lqp", "orc_configure_lqp_r", "orc_calc_lqp", "orc_lqp_levels", "orc_update_task_space"):
            getattr(L, f).argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


class OracleRobot:
    """One reference-shaped RobotData evaluated by the CPU oracle."""

    def __init__(self, model: Model | str):
        if isinstance(model, str):
            model = load_urdf(model)
        self.model = model
        f = model.flat()
        self._keep = f
        L = lib()
        self.h = L.orc_create(int(f["nb"]), f["parent"].ctypes.data_as(_ip), *[f[k].ctypes.data_as(_dp) for k in
                                                                                ("E_T", "r_T", "axis", "mass", "com", "inertia")])
        self.n = model.system_dof
        self.m = model.model_dof
        self.task_dofs = []
        self.ncontact = 0

    def __del__(self):
        try:
            lib().orc_destroy(self.h)
        except Exception:
            pass

    # ---- setup ----
    def add_contact(self, link, point, normal=(0, 0, 1), lx=0.0, ly=0.0):
        if isinstance(link, str):
            link = self.model.link_id(link)
        _, p = _d(point)
        _, nrm = _d(normal)
        r = lib().orc_add_contact(self.h, int(link), p, nrm, float(lx), float(ly))
        if r >= 0:
            self.ncontact += 1
        return r

    def add_task(self, level, mode, link, point=(0, 0, 0)):
        if isinstance(link, str):
            link = self.model.link_id(link)
        _, p = _d(point)
        r = lib().orc_add_task(self.h, int(level), int(mode), int(link), p)
        if r == 0:
            dof = 6 if mode <= 2 else 3
            if level == len(self.task_dofs):
                self.task_dofs.append(dof)
            else:
                self.task_dofs[level] += dof
        return r

    def add_custom_task(self, level, dof):
        r = lib().orc_add_custom_task(self.h, int(level), int(dof))
        if r == 0:
            self.task_dofs.append(int(dof))
        return r

    def set_torque_limit(self, tlim):
        a, p = _d(np.broadcast_to(np.asarray(tlim, dtype=np.float64), (self.m,)))
        lib().orc_set_torque_limit(self.h, p)

    # ---- cycle ----
    def update_kinematics(self, q, qd=None):
        qa, qp = _d(q)
        qda, qdp = _d(np.zeros(self.n) if qd is None else qd)
        assert qa.size == self.n + 1 and qda.size == self.n
        lib().orc_update_kinematics(self.h, qp, qdp)

    def set_contact(self, *flags):
        fl = bytes(bytearray([1 if f else 0 for f in flags] + [0] * (self.ncontact - len(flags))))
        lib().orc_set_contact(self.h, fl)

    def calc_contact_constraint(self):
        return lib().orc_calc_contact_constraint(self.h)

    def set_task(self, level, fstar, J=None):
        _, fp = _d(fstar)
        if J is None:
            lib().orc_set_task(self.h, int(level), fp, None)
        else:
            Ja = np.asfortranarray(np.asarray(J, dtype=np.float64))
            lib().orc_set_task(self.h, int(level), fp, Ja.ctypes.data_as(_dp))

    def calc_grav(self):
        lib().orc_calc_grav(self.h)
        return self.get("torque_grav")[:, 0]

    def calc_task_torque(self, hqp=True):
        return lib().orc_calc_task_torque(self.h, 1 if hqp else 0)

    def calc_redistribute(self, hqp=True):
        return lib().orc_calc_redistribute(self.h, 1 if hqp else 0)

    # ---- a9: reduced-dimension variant, reference src/dwbc.cpp:2752-3446 (call order of tests/sp_test/redu_dyn_test.cpp:262-267) ----
    def reduced_dynamics(self):
        lib().orc_reduced_dynamics(self.h)

    def reduced_contact_constraint(self):
        return lib().orc_reduced_contact_constraint(self.h)

    def reduced_grav(self):
        lib().orc_reduced_grav(self.h)
        return self.get("torque_grav")[:, 0]

    def reduced_task_space(self):
        return lib().orc_reduced_task_space(self.h)

    def reduced_task_torque(self, hqp=False):
        return lib().orc_reduced_task_torque(self.h, 1 if hqp else 0)

    def reduced_redistribute(self, hqp=True):
        """ReducedCalcContactRedistribute(hqp), reference src/dwbc.cpp:3758-3770 (hqp=False: the analytic two-foot branch)"""
        return lib().orc_reduced_redistribute(self.h, 1 if hqp else 0)

    def set_reduced_torque_limit(self, tlim):
        """torque limit of reduced_model_dof_ entries, as CalcSingleTaskTorqueWithQP_R expects (reference src/dwbc.cpp:3462-3468)"""
        a = np.ascontiguousarray(np.asarray(tlim, dtype=np.float64).reshape(-1))
        lib().orc_set_reduced_torque_limit(self.h, a.ctypes.data_as(_dp), len(a))

    def reduced_dims(self):
        out = (C.c_int * 5)()
        lib().orc_reduced_dims(self.h, out)
        return dict(zip(("co_dof", "nc_dof", "vc_dof", "reduced_system_dof", "reduced_model_dof"), out))

    # ---- a10: LQP variant, reference src/dwbc.cpp:4304-4454 (call order of tests/sp_test/jacc_compare.cpp:319-384) ----
    def update_task_space(self):
        lib().orc_update_task_space(self.h)

    def configure_lqp(self, reduced=False):
        """ConfigureLQP (reference src/dwbc.cpp:4304-4430) or, reduced=True, ConfigureLQP_R (:4504-4632; needs reduced_dynamics,
        reduced_contact_constraint and reduced_task_space first)"""
        (lib().orc_configure_lqp_r if reduced else lib().orc_configure_lqp)(self.h)
        return lib().orc_lqp_levels(self.h)

    def calc_lqp(self):
        return lib().orc_calc_lqp(self.h)

    def configure_lqp_r_nc(self):
        """ConfigureLQP_R_NC (reference src/dwbc.cpp:4634-4760) from the final answer of the reduced LQP (call configure_lqp(True) and
        calc_lqp() first); levels: 0 = whole space (virtual), 1 = centroidal equality + torque rows, 2 = the 6D task of ts_[1]"""
        lib().orc_configure_lqp_r_nc(self.h)

    def calc_lqp_r_nc(self):
        """CalcControlTorqueLQP_R_NC (:4479-4502) + the torque stitching of tests/sp_test/jacc_compare.cpp:543-569"""
        return lib().orc_calc_lqp_r_nc(self.h)

    def contact_force(self, torque):
        """getContactForce(command_torque), reference src/dwbc.cpp:891-896"""
        _, tp = _d(torque)
        out = np.zeros(6 * self.ncontact)
        k = lib().orc_contact_force(self.h, tp, out.ctypes.data_as(_dp))
        return out[:k]

    def zmp(self, contact_force):
        """getZMP(contact_force), reference src/dwbc.cpp:898-939"""
        cf = np.zeros(6 * self.ncontact)
        cf[:len(contact_force)] = contact_force
        out = np.zeros(3)
        lib().orc_zmp(self.h, cf.ctypes.data_as(_dp), out.ctypes.data_as(_dp))
        return out

    # ---- results ----
    def get(self, name):
        r, c = C.c_int(), C.c_int()
        nb = lib().orc_get(self.h, name.encode(), None, C.byref(r), C.byref(c))
        if nb < 0:
            raise KeyError(name)
        out = np.zeros((c.value, r.value), dtype=np.float64)
        lib().orc_get(self.h, name.encode(), out.ctypes.data_as(_dp), None, None)
        return out.T.copy()

    def active_set(self, level):
        buf = (C.c_int * 256)()
        k = lib().orc_get_active(self.h, int(level), buf, 256)
        return list(buf[:k])

    def qp_info(self, level):
        s, p = C.c_int(), C.c_int()
        lib().orc_get_qp_info(self.h, int(level), C.byref(s), C.byref(p))
        return s.value, p.value

    def run_batch(self, q, qd, fstars, flags, stages=0xF, nthreads=0, sets=False):
        """Whole cycle over B states. fstars: list over levels of [B, t_l]. Returns dict.  sets=True: also the active-set bit masks
        `active` / `gray` [B, levels + 1, 4] uint32 (last = redistribution QP; gray = rows whose relative slack lies within a decade
        of the 1e-9 reporting threshold), instead of contact_force / zmp."""
        q = np.ascontiguousarray(q, dtype=np.float64)
        qd = np.ascontiguousarray(qd, dtype=np.float64)
        B = q.shape[0]
        fs = np.concatenate([np.ascontiguousarray(f, dtype=np.float64).reshape(-1) for f in fstars]) if fstars else np.zeros(1)
        fl = np.ascontiguousarray(flags, dtype=np.uint8)
        assert fl.shape == (B, self.ncontact)
        tg = np.zeros((B, self.m)); tt = np.zeros((B, self.m)); tc = np.zeros((B, self.m))
        st = np.zeros(B, dtype=np.int32)
        if sets:
            nq = len(fstars) + 1
            act = np.zeros((B, nq, 4), dtype=np.uint32); gray = np.zeros((B, nq, 4), dtype=np.uint32)
            up = C.POINTER(C.c_uint)
            lib().orc_run_batch_sets(self.h, B, q.ctypes.data_as(_dp), qd.ctypes.data_as(_dp), fs.ctypes.data_as(_dp),
                                     fl.ctypes.data_as(C.c_char_p), int(stages), tg.ctypes.data_as(_dp), tt.ctypes.data_as(_dp),
                                     tc.ctypes.data_as(_dp), st.ctypes.data_as(_ip), int(nthreads), act.ctypes.data_as(up), gray.ctypes.data_as(up))
            return dict(torque_grav=tg, torque_task=tt, torque_contact=tc, status=st, active=act, gray=gray)
        cf = np.zeros((B, 6 * self.ncontact)); zmp = np.zeros((B, 3))
        lib().orc_run_batch(self.h, B, q.ctypes.data_as(_dp), qd.ctypes.data_as(_dp), fs.ctypes.data_as(_dp),
                            fl.ctypes.data_as(C.c_char_p), int(stages), tg.ctypes.data_as(_dp), tt.ctypes.data_as(_dp),
                            tc.ctypes.data_as(_dp), st.ctypes.data_as(_ip), int(nthreads), cf.ctypes.data_as(_dp), zmp.ctypes.data_as(_dp))
        return dict(torque_grav=tg, torque_task=tt, torque_contact=tc, status=st, contact_force=cf, zmp=zmp)


def lexqp(A, ub, t):
    A = np.asfortranarray(np.asarray(A, dtype=np.float64))
    ub = np.ascontiguousarray(ub, dtype=np.float64)
    x = np.zeros(A.shape[1])
    it = C.c_int()
    st = lib().orc_lexqp(A.shape[0], A.shape[1], int(t), A.ctypes.data_as(_dp), ub.ctypes.data_as(_dp), x.ctypes.data_as(_dp), C.byref(it))
    return st, x, it.value


def cod_pinv(A, thr=1e-6):
    A = np.asfortranarray(np.asarray(A, dtype=np.float64))
    out = np.zeros((A.shape[0], A.shape[1]), dtype=np.float64)  # holds (cols x rows) column-major
    rk = C.c_int()
    lib().orc_cod_pinv(A.shape[0], A.shape[1], A.ctypes.data_as(_dp), float(thr), out.ctypes.data_as(_dp), C.byref(rk))
    return out.reshape(-1).reshape((A.shape[0], A.shape[1]), order="C").reshape(-1).reshape((A.shape[1], A.shape[0]), order="F"), rk.value


def num_threads():
    return lib().orc_num_threads()
