"""RobotDataBatch: Python host over the C ABI, mirroring the reference's RobotData call sequence for B states.

reference include/dwbc.h:59-430 (RobotData) -- same method names in snake_case, same argument meaning:
    UpdateKinematics -> update_kinematics(q, qdot)        SetContact -> set_contact(flags)
    CalcContactConstraint -> calc_contact_constraint()    SetTaskSpace -> set_task_space(level, fstar, J=None)
    CalcGravCompensation / CalcTaskControlTorque / CalcContactRedistribute -> solve(stages)
Public RobotData fields (A_, A_inv_, J_C, Lambda_contact, ... torque_task_) are read with get(name).

Device memory and streams come from PyTorch (plumbing only): every tensor argument must be a CUDA tensor on the
batch's device, float64 (uint8 for contact flags) and contiguous.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import DwbcError, load_library

CONTACT_6D = 0
(TASK_LINK_6D, TASK_LINK_6D_COM_FRAME, TASK_LINK_6D_CUSTOM_FRAME, TASK_LINK_POSITION, TASK_LINK_POSITION_COM_FRAME,
 TASK_LINK_POSITION_CUSTOM_FRAME, TASK_LINK_ROTATION, TASK_LINK_ROTATION_CUSTOM_FRAME) = range(8)
SOLVE_GRAV, SOLVE_TASK, SOLVE_HQP, SOLVE_REDIST, SOLVE_REDIST_ANALYTIC = 1, 2, 4, 8, 16
ST_CONTACT_OK, ST_TASK_OK, ST_REDIST_OK, ST_QP_SLOW = 1, 2, 4, 16
OUT_AUTO, OUT_TORQUES, OUT_ALL = -1, 0, 1

# name -> (matrix id, per-level?)
_MATS = {
    "A": (0, False), "A_inv": (1, False), "G": (2, False), "B": (3, False), "J_C": (4, False), "Lambda_contact": (5, False),
    "J_C_INV_T": (6, False), "N_C": (7, False), "A_inv_N_C": (8, False), "W": (9, False), "W_inv": (10, False), "V2": (11, False),
    "NwJw": (12, False), "P_C": (13, False), "torque_grav": (14, False), "torque_task": (15, False), "torque_contact": (16, False),
    "com_pos": (17, False), "link_pos": (18, False), "link_rot": (19, False), "link_vel": (20, False), "cf_redis_qp": (21, False),
    "J_task": (64, True), "Lambda_task": (65, True), "J_kt": (66, True), "Null_task": (67, True), "f_star_qp": (68, True),
    "contact_qp": (69, True), "torque_h": (70, True),
}

# reduced-dimension variant: name -> (matrix id, per-level?)   (DWBC_RMAT_* of include/dwbc_b200.h)
_RMATS = {
    "A_R_inv": (0, False), "A_R": (1, False), "J_I_nc": (2, False), "J_I_nc_inv_T": (3, False), "G_R": (4, False), "J_CR": (5, False),
    "Lambda_CR": (6, False), "J_CR_INV_T": (7, False), "A_R_inv_N_CR": (8, False), "W_R": (9, False), "W_R_inv": (10, False),
    "NwJw_R": (11, False), "P_CR": (12, False), "torque_grav_R": (13, False), "torque_task_R": (14, False), "J_base_R_kt": (15, False),
    "SI_nc": (16, False), "J_task_R": (64, True), "Lambda_task": (65, True), "J_kt_R": (66, True), "Null_task_R": (67, True),
}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class _CudaBackend:
    """Device buffers = torch CUDA tensors."""

    name = "cuda"

    def __init__(self, device):
        import torch

        if not torch.cuda.is_available():
            raise DwbcError("libdwbc_b200 needs a CUDA device (no CPU fallback)")
        self.torch = torch
        self.device = torch.device("cuda", device)
        self.index = device

    def lib(self):
        return load_library()

    def empty(self, shape, dtype="f8"):
        t = self.torch
        dt = {"f8": t.float64, "f4": t.float32, "u1": t.uint8, "i4": t.int32, "u4": t.int32}[dtype]
        return t.empty(shape, dtype=dt, device=self.device)

    def check(self, x, dtype="f8"):
        t = self.torch
        dt = {"f8": t.float64, "u1": t.uint8, "i4": t.int32}[dtype]
        if not isinstance(x, t.Tensor) or not x.is_cuda or x.device != self.device or x.dtype != dt or not x.is_contiguous():
            raise DwbcError(f"expected a contiguous {dt} CUDA tensor on {self.device}")
        return x

    def ptr(self, x):
        return C.c_void_p(x.data_ptr())

    def stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def to_numpy(self, x):
        return x.cpu().numpy()

    def from_numpy(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a)).to(self.device)


class RobotDataBatch:
    def __init__(self, urdf_path: str, batch: int, device: int = 0, floating: bool = True, _backend=None):
        self.be = _backend if _backend is not None else _CudaBackend(device)
        self.L = self.be.lib()
        self._model = self.L.dwbc_model_load_urdf(urdf_path.encode(), 1 if floating else 0)
        if not self._model:
            raise DwbcError(self._err())
        n, m, nb, mass = C.c_int(), C.c_int(), C.c_int(), C.c_double()
        self.L.dwbc_model_info(self._model, C.byref(n), C.byref(m), C.byref(nb), C.byref(mass))
        self.system_dof_, self.model_dof_, self.n_links, self.total_mass_ = n.value, m.value, nb.value, mass.value
        self.batch = int(batch)
        self._h = self.L.dwbc_batch_create(self._model, self.batch, getattr(self.be, "index", 0))
        if not self._h:
            raise DwbcError(self._err())
        cm, rm, ml, mt = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self.L.dwbc_batch_limits(self._h, C.byref(cm), C.byref(rm), C.byref(ml), C.byref(mt))
        self.cmax, self.rmax = cm.value, rm.value
        self.task_dofs = []
        self.ncontact = 0
        self._keep = {}

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self.L.dwbc_batch_destroy(self._h)
            if getattr(self, "_model", None):
                self.L.dwbc_model_destroy(self._model)
        except Exception:
            pass

    def _err(self):
        return self.L.dwbc_last_error().decode()

    def _ck(self, rc):
        if rc < 0:
            raise DwbcError(f"dwbc error {rc}: {self._err()}")
        return rc

    # ---- model queries ----
    def link_id(self, name: str) -> int:
        return self.L.dwbc_model_link_id(self._model, name.encode())

    def link_name(self, link: int) -> str:
        s = self.L.dwbc_model_link_name(self._model, link)
        return s.decode() if s else ""

    # ---- setup ----
    def _link(self, link):
        if isinstance(link, str):
            i = self.link_id(link)
            if i < 0:
                raise DwbcError("Link Name is Wrong : " + link)
            return i
        return int(link)

    def add_contact_constraint(self, link, contact_type, point, normal=(0, 0, 1), lx=0.0, ly=0.0):
        p = (C.c_double * 3)(*point)
        nrm = (C.c_double * 3)(*normal)
        idx = self._ck(self.L.dwbc_batch_add_contact(self._h, self._link(link), int(contact_type), p, nrm, float(lx), float(ly)))
        self.ncontact += 1
        return idx

    def set_friction_ratio(self, contact, mu, muz):
        self._ck(self.L.dwbc_batch_set_friction(self._h, int(contact), float(mu), float(muz)))

    def add_task_space(self, level, task_mode, link, point=(0, 0, 0)):
        p = (C.c_double * 3)(*point)
        self._ck(self.L.dwbc_batch_add_task(self._h, int(level), int(task_mode), self._link(link), p))
        dof = 6 if task_mode <= TASK_LINK_6D_CUSTOM_FRAME else 3
        if level == len(self.task_dofs):
            self.task_dofs.append(dof)
        else:
            self.task_dofs[level] += dof

    def add_custom_task_space(self, level, task_dof):
        self._ck(self.L.dwbc_batch_add_custom_task(self._h, int(level), int(task_dof)))
        self.task_dofs.append(int(task_dof))

    def clear_task_space(self):
        self._ck(self.L.dwbc_batch_clear_tasks(self._h))
        self.task_dofs = []
        # the C side drops the gains / trajectory pointers of the removed task links: drop the buffers kept alive for them
        for k in [k for k in self._keep if isinstance(k, tuple) and k[0] in ("trajp", "trajr", "f", "J")]:
            del self._keep[k]

    def set_torque_limit(self, tlim):
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(tlim, dtype=np.float64), (self.model_dof_,)))
        self._ck(self.L.dwbc_batch_set_torque_limit(self._h, a.ctypes.data_as(_dp)))

    # ---- per cycle (device tensors) ----
    def update_kinematics(self, q, qdot):
        be = self.be
        q, qdot = be.check(q), be.check(qdot)
        if tuple(q.shape) != (self.batch, self.system_dof_ + 1) or tuple(qdot.shape) != (self.batch, self.system_dof_):
            raise DwbcError("q size is not matched")
        self._keep["q"], self._keep["qd"] = q, qdot
        self._ck(self.L.dwbc_batch_update_kinematics(self._h, be.ptr(q), be.ptr(qdot), be.stream()))

    def set_contact(self, flags):
        be = self.be
        flags = be.check(flags, "u1")
        if tuple(flags.shape) != (self.batch, self.ncontact):
            raise DwbcError("contact flags must be [batch, number of registered contacts]")
        self._keep["flags"] = flags
        self._ck(self.L.dwbc_batch_set_contact(self._h, be.ptr(flags), be.stream()))

    def calc_contact_constraint(self):
        self._ck(self.L.dwbc_batch_calc_contact_constraint(self._h, self.be.stream()))

    def set_task_space(self, level, fstar, J=None):
        be = self.be
        fstar = be.check(fstar)
        t = self.task_dofs[level]
        if tuple(fstar.shape) != (self.batch, t):
            raise DwbcError("ERROR : task dof not matching!")
        self._keep[("f", level)] = fstar
        jp, per_state = None, 0
        if J is not None:
            J = be.check(J)
            # J is given in the natural (rows = task dof, cols = system dof) orientation; the ABI wants column-major
            if J.ndim == 2:
                Jc = J.t().contiguous()
            else:
                Jc = J.transpose(1, 2).contiguous()
                per_state = 1
            self._keep[("J", level)] = Jc
            jp = be.ptr(Jc)
        self._ck(self.L.dwbc_batch_set_task(self._h, int(level), be.ptr(fstar), jp, per_state, be.stream()))

    def solve(self, stages=SOLVE_GRAV | SOLVE_TASK | SOLVE_HQP | SOLVE_REDIST, out=None):
        """Runs the selected stages; returns (torque_grav, torque_task, torque_contact, status) device tensors."""
        be = self.be
        B, m = self.batch, self.model_dof_
        if out is None:
            out = (be.empty((B, m)), be.empty((B, m)), be.empty((B, m)), be.empty((B,), "i4"))
        tg, tt, tc, st = out
        self._ck(self.L.dwbc_batch_solve(self._h, int(stages), be.ptr(tg), be.ptr(tt), be.ptr(tc), be.ptr(st), be.stream()))
        return tg, tt, tc, st

    # ---- a9: reduced-dimension variant (reference src/dwbc.cpp:2752-3446; call order of tests/sp_test/redu_dyn_test.cpp:262-267) ----
    def reduced_dynamics_calculate(self, pattern):
        """ReducedDynamicsCalculate for the contact pattern (one 0/1 flag per registered contact) every state shows."""
        pat = (C.c_uint8 * max(self.ncontact, 1))(*[1 if f else 0 for f in pattern])
        self._ck(self.L.dwbc_batch_reduced_dynamics(self._h, pat, self.be.stream()))
        d = [C.c_int() for _ in range(5)]
        self._ck(self.L.dwbc_batch_reduced_dims(self._h, *[C.byref(x) for x in d]))
        self.co_dof, self.nc_dof, self.vc_dof, self.reduced_system_dof_, self.reduced_model_dof_ = [x.value for x in d]

    def set_reduced_torque_limit(self, tlim):
        """torque limit of the reduced QPs (reduced_model_dof_ entries, reference src/dwbc.cpp:3462-3468); None clears it"""
        if tlim is None:
            self._ck(self.L.dwbc_batch_set_reduced_torque_limit(self._h, None, 0))
            return
        a = np.ascontiguousarray(np.asarray(tlim, dtype=np.float64).reshape(-1))
        self._ck(self.L.dwbc_batch_set_reduced_torque_limit(self._h, a.ctypes.data_as(_dp), len(a)))

    def reduced_calc_contact_constraint(self):
        self._ck(self.L.dwbc_batch_reduced_calc_contact_constraint(self._h, self.be.stream()))

    def reduced_solve(self, stages=SOLVE_GRAV | SOLVE_TASK, out=None):
        """ReducedCalcGravCompensation + ReducedCalcTaskSpace + ReducedCalcTaskControlTorque (SOLVE_HQP: hqp=true) +
        ReducedCalcContactRedistribute (SOLVE_REDIST); returns (torque_grav, torque_task, torque_contact, status) like solve()."""
        be = self.be
        B, m = self.batch, self.model_dof_
        if out is None:
            out = (be.empty((B, m)), be.empty((B, m)), be.empty((B, m)), be.empty((B,), "i4"))
        tg, tt, tc, st = out
        self._ck(self.L.dwbc_batch_reduced_solve(self._h, int(stages), be.ptr(tg), be.ptr(tt), be.ptr(tc), be.ptr(st), be.stream()))
        return tg, tt, tc, st

    def get_reduced(self, name, level=0):
        return self._get(self.L.dwbc_batch_get_reduced_matrix, _RMATS[name][0], level)

    # ---- f* from trajectories, MuJoCo layout, float32 torque command ----
    def set_task_gain(self, level, link_index, pos_p, pos_d, pos_a, rot_p, rot_d, rot_a):
        """TaskLink::SetTaskGain (reference src/task.cpp:253-264)"""
        g = np.concatenate([np.broadcast_to(np.asarray(x, dtype=np.float64), (3,)) for x in (pos_p, pos_d, pos_a, rot_p, rot_d, rot_a)])
        g = np.ascontiguousarray(g)
        self._ck(self.L.dwbc_batch_set_task_gain(self._h, int(level), int(link_index), g.ctypes.data_as(_dp)))

    def set_trajectory_quintic(self, level, link_index, time, pos_init, vel_init, pos_desired, vel_desired):
        """TaskLink::SetTrajectoryQuintic per state: time [B, 2] = (start, end), the rest [B, 3] (device tensors, kept alive here)."""
        be = self.be
        a = [be.check(x) for x in (time, pos_init, vel_init, pos_desired, vel_desired)]
        self._keep[("trajp", level, link_index)] = a
        self._ck(self.L.dwbc_batch_set_trajectory_quintic(self._h, int(level), int(link_index), *[be.ptr(x) for x in a]))

    def set_trajectory_rotation(self, level, link_index, time, rot_init, rot_desired):
        """TaskLink::SetTrajectoryRotation per state: time [B, 2], rotations [B, 3, 3] (row-major)."""
        be = self.be
        a = [be.check(x) for x in (time, rot_init, rot_desired)]
        self._keep[("trajr", level, link_index)] = a
        self._ck(self.L.dwbc_batch_set_trajectory_rotation(self._h, int(level), int(link_index), *[be.ptr(x) for x in a]))

    def calc_fstar(self, level, control_time, fstar=None):
        """GetFstarPosPD / GetFstarRotPD of every task link of `level` whose trajectory is set; returns f* [B, t]."""
        be = self.be
        ct = be.check(control_time)
        if fstar is None:
            fstar = be.empty((self.batch, self.task_dofs[level]))
        self._ck(self.L.dwbc_batch_calc_fstar(self._h, int(level), be.ptr(ct), be.ptr(be.check(fstar)), be.stream()))
        return fstar

    def update_kinematics_mujoco(self, qpos, qvel):
        """RlWBCBridge::UpdateKinematics (reference src/pybind/rl_wbc_bridge.cpp:82-107): MuJoCo-ordered qpos / qvel."""
        be = self.be
        qpos, qvel = be.check(qpos), be.check(qvel)
        if tuple(qpos.shape) != (self.batch, self.system_dof_ + 1) or tuple(qvel.shape) != (self.batch, self.system_dof_):
            raise DwbcError("q size is not matched")
        self._keep["q"], self._keep["qd"] = qpos, qvel
        self._ck(self.L.dwbc_batch_update_kinematics_mujoco(self._h, be.ptr(qpos), be.ptr(qvel), be.stream()))

    def get_torque_command(self):
        """RlWBCBridge::getTorqueCommand (reference src/pybind/rl_wbc_bridge.cpp:131-139): float32 [B, m]."""
        be = self.be
        out = be.empty((self.batch, self.model_dof_), "f4")
        self._ck(self.L.dwbc_batch_get_torque_command_f32(self._h, be.ptr(out), be.stream()))
        return out

    # ---- a10: LQP variant (reference src/dwbc.cpp:4304-4454, src/dwbc_hqp.cpp; call order of tests/sp_test/jacc_compare.cpp:367-384) ----
    def calc_control_torque_lqp(self, out=None):
        """ConfigureLQP + CalcControlTorqueLQP for every state: returns (y = [qdd; Fc] [B, n + cmax], torque [B, m], status [B])."""
        be = self.be
        B, m, nv = self.batch, self.model_dof_, self.system_dof_ + self.cmax
        if out is None:
            out = (be.empty((B, nv)), be.empty((B, m)), be.empty((B,), "i4"))
        y, tau, st = out
        self._ck(self.L.dwbc_batch_calc_control_torque_lqp(self._h, be.ptr(y), be.ptr(tau), be.ptr(st), be.stream()))
        return y, tau, st

    def calc_control_torque_lqp_r(self, out=None):
        """ConfigureLQP_R + CalcControlTorqueLQP_R (after reduced_dynamics_calculate, reduced_calc_contact_constraint and
        reduced_solve(SOLVE_TASK)): returns (y [B, reduced_system_dof + cmax], torque [B, reduced_model_dof], status [B])."""
        be = self.be
        B = self.batch
        if out is None:
            out = (be.empty((B, self.reduced_system_dof_ + self.cmax)), be.empty((B, self.reduced_model_dof_)), be.empty((B,), "i4"))
        y, tau, st = out
        self._ck(self.L.dwbc_batch_calc_control_torque_lqp_r(self._h, be.ptr(y), be.ptr(tau), be.ptr(st), be.stream()))
        return y, tau, st

    def calc_control_torque_lqp_r_nc(self, out=None):
        """ConfigureLQP_R_NC + CalcControlTorqueLQP_R_NC (reference src/dwbc.cpp:4634-4760, 4479-4502) after
        calc_control_torque_lqp_r: returns (qdd_nc [B, nc_dof], stitched joint torque [B, m] as reference
        tests/sp_test/jacc_compare.cpp:543-569, status [B])."""
        be = self.be
        B = self.batch
        if out is None:
            out = (be.empty((B, self.nc_dof)), be.empty((B, self.model_dof_)), be.empty((B,), "i4"))
        y, tau, st = out
        self._ck(self.L.dwbc_batch_calc_control_torque_lqp_r_nc(self._h, be.ptr(y), be.ptr(tau), be.ptr(st), be.stream()))
        return y, tau, st

    def get_lqp(self, name, level=0, reduced=False):
        """"y" (answer after `level`), "v1" (slack of the level-1 inequality rows), "torque"; reduced=True: ConfigureLQP_R results."""
        return self._get(self.L.dwbc_batch_get_lqp, {"y": 0, "v1": 1, "torque": 2}[name] | (256 if reduced else 0), level)

    def lqp_iterations(self):
        buf = self.be.empty((self.batch,), "i4")
        self._ck(self.L.dwbc_batch_get_lqp_iterations(self._h, self.be.ptr(buf), self.be.stream()))
        return buf

    def get(self, name, level=0):
        """Public RobotData field `name` for all states: tensor [B, rows, cols] (vectors: [B, rows])."""
        return self._get(self.L.dwbc_batch_get_matrix, _MATS[name][0], level)

    def _get(self, fn, mid, level):
        be = self.be
        r, c = C.c_int(), C.c_int()
        self._ck(fn(self._h, mid, int(level), None, C.byref(r), C.byref(c), be.stream()))
        buf = be.empty((self.batch, c.value, r.value))  # column-major per state
        self._ck(fn(self._h, mid, int(level), be.ptr(buf), None, None, be.stream()))
        out = buf.transpose(1, 2) if hasattr(buf, "transpose") and not isinstance(buf, np.ndarray) else np.swapaxes(buf, 1, 2)
        if c.value == 1:
            out = out[:, :, 0]
        return out

    def get_contact_force(self, torque=None):
        """getContactForce(command_torque) (reference src/dwbc.cpp:891-896): [B, cmax]; torque=None uses
        torque_grav_ + torque_task_ + torque_contact_ of the last solve."""
        be = self.be
        out = be.empty((self.batch, self.cmax))
        tp = None
        if torque is not None:
            torque = be.check(torque)
            if tuple(torque.shape) != (self.batch, self.model_dof_):
                raise DwbcError("command torque must be [batch, model_dof]")
            tp = be.ptr(torque)
        self._ck(self.L.dwbc_batch_get_contact_force(self._h, tp, be.ptr(out), be.stream()))
        return out

    def get_zmp(self, contact_force):
        """getZMP(contact_force) (reference src/dwbc.cpp:898-939): [B, 3]"""
        be = self.be
        cf = be.check(contact_force)
        if tuple(cf.shape) != (self.batch, self.cmax):
            raise DwbcError("contact force must be [batch, cmax]")
        out = be.empty((self.batch, 3))
        self._ck(self.L.dwbc_batch_get_zmp(self._h, be.ptr(cf), be.ptr(out), be.stream()))
        return out

    def active_set(self, level):
        """Active rows of the level's QP (level = -1: redistribution QP) as a list of row-index lists."""
        be = self.be
        buf = be.empty((self.batch, 4), "u4")
        self._ck(self.L.dwbc_batch_get_active_set(self._h, int(level), be.ptr(buf), be.stream()))
        words = be.to_numpy(buf).astype(np.uint32)
        res = []
        for b in range(self.batch):
            rows = [i for i in range(128) if (int(words[b, i >> 5]) >> (i & 31)) & 1]
            res.append(rows)
        return res

    def qp_iterations(self):
        buf = self.be.empty((self.batch,), "i4")
        self._ck(self.L.dwbc_batch_get_qp_iterations(self._h, self.be.ptr(buf), self.be.stream()))
        return buf

    # ---- output groups ----
    def set_outputs(self, outputs):
        """dwbc_batch_set_outputs: OUT_ALL (every public RobotData field, generic kernels), OUT_TORQUES (torques + status only, the
        one-kernel cycle; the staged calls then only record their arguments until solve()), OUT_AUTO (default)."""
        self._ck(self.L.dwbc_batch_set_outputs(self._h, int(outputs)))

    def last_path(self):
        """1 if the last solve / cycle_host ran the one-kernel cycle (lean_cycle_kernel), 0 for the generic kernels"""
        return self.L.dwbc_batch_last_path(self._h)

    def workspace_bytes(self):
        """device memory the handle has allocated so far (lazy: grows when DWBC_OUT_ALL is first used)"""
        return int(self.L.dwbc_batch_workspace_bytes(self._h))

    # ---- per-kernel device timing (CUDA events recorded by the library around each of its launches) ----
    def set_profiling(self, enable=True):
        self._ck(self.L.dwbc_batch_set_profiling(self._h, 1 if enable else 0))

    def kernel_times(self, reset=True):
        """{kernel name: (summed ms while profiling was on, launches since the last reset)}"""
        K = 0
        while self.L.dwbc_kernel_name(K):
            K += 1
        ms = (C.c_double * K)()
        cnt = (C.c_longlong * K)()
        self._ck(self.L.dwbc_batch_get_kernel_times(self._h, ms, cnt, 1 if reset else 0))
        return {self.L.dwbc_kernel_name(k).decode(): (ms[k], cnt[k]) for k in range(K)}

    # ---- whole cycle from host buffers (pinned numpy / torch CPU tensors) ----
    def set_host_slices(self, slices):
        """dwbc_batch_set_host_slices: slices cycle_host pipelines the batch in (0 = automatic, 1 = no slicing)"""
        self._ck(self.L.dwbc_batch_set_host_slices(self._h, int(slices)))

    def cycle_host(self, stages, q, qdot, fstars, flags, out=None):
        """H2D + all stages + D2H through dwbc_batch_cycle_host.  All arguments are host arrays (numpy or torch CPU tensors):
        q [B, n+1] float64, qdot [B, n] float64, fstars one [B, t_l] float64 array per level (or their concatenation), flags
        [B, ncontact] uint8.  Wrong dtypes are converted, wrong shapes raise: the C side reads raw pointers."""
        B, m, n = self.batch, self.model_dof_, self.system_dof_

        def host(a, dtype, shape, what):
            if not isinstance(a, np.ndarray):
                a = a.numpy() if hasattr(a, "numpy") else np.asarray(a)
            if a.dtype != dtype and dtype == np.uint8 and a.dtype not in (np.bool_, np.int8, np.int32, np.int64):
                raise DwbcError(f"cycle_host: {what} must be an integer / bool array")
            a = np.ascontiguousarray(a, dtype=dtype)
            if a.size != int(np.prod(shape)):
                raise DwbcError(f"cycle_host: {what} has {a.size} entries, expected {shape}")
            return a

        q = host(q, np.float64, (B, n + 1), "q")
        qdot = host(qdot, np.float64, (B, n), "qdot")
        flags = host(flags, np.uint8, (B, self.ncontact), "flags")
        fcat = None
        if fstars is not None and not (isinstance(fstars, (list, tuple)) and len(fstars) == 0):
            dofs = list(self.task_dofs)
            if isinstance(fstars, (list, tuple)):
                if len(fstars) != len(dofs):
                    raise DwbcError(f"cycle_host: {len(fstars)} f* arrays for {len(dofs)} hierarchy levels")
                fcat = np.concatenate([host(f, np.float64, (B, dofs[l]), f"fstar[{l}]").reshape(-1) for l, f in enumerate(fstars)])
            else:
                fcat = host(fstars, np.float64, (B * sum(dofs),), "fstar")
        if out is None:
            out = (np.empty((B, m)), np.empty((B, m)), np.empty((B, m)), np.empty((B,), dtype=np.int32))
        tg, tt, tc, st = out
        for a, dt, sh, what in ((tg, np.float64, B * m, "tau_grav out"), (tt, np.float64, B * m, "tau_task out"),
                                (tc, np.float64, B * m, "tau_contact out"), (st, np.int32, B, "status out")):
            if a is None:
                continue
            an = a if isinstance(a, np.ndarray) else a.numpy()
            if an.dtype != dt or an.size != sh or not an.flags["C_CONTIGUOUS"]:
                raise DwbcError(f"cycle_host: {what} must be a contiguous {np.dtype(dt).name} array with {sh} entries")

        def hp(a):
            if a is None:
                return None
            return C.c_void_p(a.ctypes.data if isinstance(a, np.ndarray) else a.data_ptr())

        self._ck(self.L.dwbc_batch_cycle_host(self._h, int(stages), hp(q), hp(qdot), hp(fcat), hp(flags),
                                             hp(tg), hp(tt), hp(tc), hp(st), self.be.stream()))
        return tg, tt, tc, st
