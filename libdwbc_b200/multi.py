"""RobotDataMultiBatch: ONE host batch over several GPUs of one box through dwbc_multi_* (include/dwbc_b200.h).

SURVEY.md section 8(e): the states are independent, so the batch is cut into contiguous slices, one per GPU; every slice is driven by
its own host thread, stream and dwbc_batch_t inside the library, reads and writes its own part of the caller's (pinned) host arrays and
the call returns when all slices are done.  There is no collective on the data path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import DwbcError, load_library

_dp = C.POINTER(C.c_double)


def declare_multi(L):
    vp = C.c_void_p
    L.dwbc_multi_create.restype = vp
    L.dwbc_multi_create.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_int)]
    L.dwbc_multi_destroy.argtypes = [vp]
    L.dwbc_multi_gpus.argtypes = [vp]
    L.dwbc_multi_shard.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.dwbc_multi_add_contact.argtypes = [vp, C.c_int, C.c_int, _dp, _dp, C.c_double, C.c_double]
    L.dwbc_multi_add_task.argtypes = [vp, C.c_int, C.c_int, C.c_int, _dp]
    L.dwbc_multi_add_custom_task.argtypes = [vp, C.c_int, C.c_int, _dp]
    L.dwbc_multi_set_torque_limit.argtypes = [vp, _dp]
    L.dwbc_multi_set_outputs.argtypes = [vp, C.c_int]
    L.dwbc_multi_cycle_host.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp]
    L.dwbc_multi_last_shard_ms.argtypes = [vp, _dp]
    L.dwbc_host_alloc.restype = vp
    L.dwbc_host_alloc.argtypes = [C.c_size_t]
    L.dwbc_host_free.argtypes = [vp]
    L.dwbc_batch_last_path.argtypes = [vp]
    return L


class RobotDataMultiBatch:
    def __init__(self, urdf_path: str, batch: int, ngpus: int, devices=None, floating: bool = True, _lib=None):
        self.L = declare_multi(_lib if _lib is not None else load_library())
        self._model = self.L.dwbc_model_load_urdf(urdf_path.encode(), 1 if floating else 0)
        if not self._model:
            raise DwbcError(self.L.dwbc_last_error().decode())
        n, m, nb, mass = C.c_int(), C.c_int(), C.c_int(), C.c_double()
        self.L.dwbc_model_info(self._model, C.byref(n), C.byref(m), C.byref(nb), C.byref(mass))
        self.system_dof_, self.model_dof_, self.n_links = n.value, m.value, nb.value
        dev = (C.c_int * ngpus)(*devices) if devices is not None else None
        self._h = self.L.dwbc_multi_create(self._model, int(batch), int(ngpus), dev)
        if not self._h:
            raise DwbcError(self.L.dwbc_last_error().decode())
        self.batch, self.ngpus = int(batch), self.L.dwbc_multi_gpus(self._h)
        self.task_dofs, self.ncontact = [], 0
        self._pinned = []

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self.L.dwbc_multi_destroy(self._h)
            for p in getattr(self, "_pinned", []):
                self.L.dwbc_host_free(p)
            if getattr(self, "_model", None):
                self.L.dwbc_model_destroy(self._model)
        except Exception:
            pass

    def _ck(self, rc):
        if rc < 0:
            raise DwbcError(f"dwbc error {rc}: {self.L.dwbc_last_error().decode()}")
        return rc

    def link_id(self, name):
        return self.L.dwbc_model_link_id(self._model, name.encode())

    def _link(self, link):
        if isinstance(link, str):
            i = self.link_id(link)
            if i < 0:
                raise DwbcError("Link Name is Wrong : " + link)
            return i
        return int(link)

    # ---- setup, same meaning as RobotDataBatch ----
    def add_contact_constraint(self, link, contact_type, point, normal=(0, 0, 1), lx=0.0, ly=0.0):
        idx = self._ck(self.L.dwbc_multi_add_contact(self._h, self._link(link), int(contact_type), (C.c_double * 3)(*point), (C.c_double * 3)(*normal), float(lx), float(ly)))
        self.ncontact += 1
        return idx

    def add_task_space(self, level, task_mode, link, point=(0, 0, 0)):
        self._ck(self.L.dwbc_multi_add_task(self._h, int(level), int(task_mode), self._link(link), (C.c_double * 3)(*point)))
        dof = 6 if task_mode <= 2 else 3
        if level == len(self.task_dofs):
            self.task_dofs.append(dof)
        else:
            self.task_dofs[level] += dof

    def add_custom_task_space(self, level, task_dof, jacobian):
        """TASK_CUSTOM level with its [t, n] Jacobian (numpy, shared by all states)"""
        J = np.asfortranarray(np.asarray(jacobian, dtype=np.float64))
        if J.shape != (task_dof, self.system_dof_):
            raise DwbcError("custom task Jacobian must be [task_dof, system_dof]")
        self._ck(self.L.dwbc_multi_add_custom_task(self._h, int(level), int(task_dof), J.ctypes.data_as(_dp)))
        self.task_dofs.append(int(task_dof))

    def set_torque_limit(self, tlim):
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(tlim, dtype=np.float64), (self.model_dof_,)))
        self._ck(self.L.dwbc_multi_set_torque_limit(self._h, a.ctypes.data_as(_dp)))

    def set_outputs(self, outputs):
        self._ck(self.L.dwbc_multi_set_outputs(self._h, int(outputs)))

    def shards(self):
        out = []
        for g in range(self.ngpus):
            h, d, f, c = C.c_void_p(), C.c_int(), C.c_int(), C.c_int()
            self._ck(self.L.dwbc_multi_shard(self._h, g, C.byref(h), C.byref(d), C.byref(f), C.byref(c)))
            out.append({"device": d.value, "first": f.value, "count": c.value, "last_path": self.L.dwbc_batch_last_path(h)})
        return out

    # ---- pinned host arrays every GPU can copy from / to ----
    def pinned(self, shape, dtype=np.float64):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = self.L.dwbc_host_alloc(max(n, 16))
        if not p:
            raise DwbcError(self.L.dwbc_last_error().decode())
        self._pinned.append(p)
        return np.ctypeslib.as_array((C.c_ubyte * n).from_address(p)).view(dtype).reshape(shape)

    # ---- the whole cycle for the whole batch ----
    def cycle_host(self, stages, q, qdot, fstars, flags, out=None):
        """q [B, n+1], qdot [B, n], fstars: one [B, t_l] array per level, flags [B, ncontact] uint8 (numpy, ideally from pinned());
        returns (torque_grav, torque_task, torque_contact [B, m], status [B])."""
        B, m, n = self.batch, self.model_dof_, self.system_dof_
        q = np.ascontiguousarray(q, dtype=np.float64)
        qdot = np.ascontiguousarray(qdot, dtype=np.float64)
        flags = np.ascontiguousarray(flags, dtype=np.uint8)
        if q.shape != (B, n + 1) or qdot.shape != (B, n) or flags.shape != (B, self.ncontact):
            raise DwbcError("cycle_host: q / qdot / flags have the wrong shape")
        fcat = None
        if fstars is not None and len(fstars):
            if isinstance(fstars, np.ndarray) and fstars.ndim == 1:
                fcat = np.ascontiguousarray(fstars, dtype=np.float64)
                if fcat.size != B * sum(self.task_dofs):
                    raise DwbcError("cycle_host: f* has the wrong size")
            else:
                if len(fstars) != len(self.task_dofs):
                    raise DwbcError("cycle_host: one f* array per hierarchy level")
                parts = []
                for l, f in enumerate(fstars):
                    f = np.ascontiguousarray(f, dtype=np.float64)
                    if f.shape != (B, self.task_dofs[l]):
                        raise DwbcError(f"cycle_host: fstar[{l}] must be [B, {self.task_dofs[l]}]")
                    parts.append(f.reshape(-1))
                fcat = np.concatenate(parts)
        if out is None:
            out = (np.empty((B, m)), np.empty((B, m)), np.empty((B, m)), np.empty((B,), dtype=np.int32))
        tg, tt, tc, st = out
        hp = lambda a: None if a is None else C.c_void_p(a.ctypes.data)
        self._ck(self.L.dwbc_multi_cycle_host(self._h, int(stages), hp(q), hp(qdot), hp(fcat), hp(flags), hp(tg), hp(tt), hp(tc), hp(st)))
        return tg, tt, tc, st

    def last_shard_ms(self):
        a = (C.c_double * self.ngpus)()
        self._ck(self.L.dwbc_multi_last_shard_ms(self._h, a))
        return list(a)
