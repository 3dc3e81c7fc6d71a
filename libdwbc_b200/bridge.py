"""RlWBCBridge / RlWBCBridgeBatch: the reference's Python entry point (reference src/pybind/rl_wbc_bridge.cpp:8-158) over the
batched C ABI -- same method names, argument meaning and wire format (MuJoCo-ordered qpos / qvel in, float32 torque command
out).  `RlWBCBridge(env_id)` is the single-environment class of the reference; `RlWBCBridgeBatch(num_envs)` steps many
environments per call with device-resident tensors.

Behaviour kept from the reference: CalcTorque() passes `task_init = false` as the FIRST
positional argument of CalcTaskControlTorque / CalcContactRedistribute (src/pybind/rl_wbc_bridge.cpp:123-129), i.e. hqp = false:
closed-form task torque and the analytic two-foot contact redistribution.  `hqp=True` selects the QP branches instead.
There is no CPU fallback: construction raises DwbcError without the CUDA library and a GPU."""
import os

import numpy as np

from ._lib import DwbcError
from .batch import (CONTACT_6D, SOLVE_GRAV, SOLVE_HQP, SOLVE_REDIST, SOLVE_REDIST_ANALYTIC, SOLVE_TASK, TASK_LINK_6D, TASK_LINK_ROTATION,
                    RobotDataBatch)

# reference src/pybind/rl_wbc_bridge.cpp:22-28 (initial posture, RBDL order)
Q_INIT = [0, 0, 0.92983, 0, 0, 0, 0.0, 0.0, -0.24, 0.6, -0.36, 0.0, 0.0, 0.0, -0.24, 0.6, -0.36, 0.0, 0, 0, 0,
          0.3, 0.3, 1.5, -1.27, -1, 0, -1, 0, 0, 0, -0.3, -0.3, -1.5, 1.27, 1, 0, 1, 0, 1]


def _default_urdf():
    d = os.environ.get("DWBC_URDF_DIR")  # the reference bakes URDF_DIR in at compile time (src/pybind/rl_wbc_bridge.cpp:11-13)
    if not d:
        raise DwbcError("pass urdf_path or set DWBC_URDF_DIR to the directory holding dyros_tocabi.urdf")
    return os.path.join(d, "dyros_tocabi.urdf")


class RlWBCBridgeBatch:
    def __init__(self, num_envs, device=0, urdf_path=None, hqp=False, _backend=None):
        self.rd_ = RobotDataBatch(urdf_path or _default_urdf(), int(num_envs), device, _backend=_backend)
        rd = self.rd_
        self.num_envs = int(num_envs)
        self.hqp = bool(hqp)
        # reference src/pybind/rl_wbc_bridge.cpp:31-44
        rd.add_contact_constraint(6, CONTACT_6D, (0.03, 0, -0.1585), (0, 0, 1), 0.15, 0.075)
        rd.add_contact_constraint(12, CONTACT_6D, (0.03, 0, -0.1585), (0, 0, 1), 0.15, 0.075)
        rd.add_task_space(0, TASK_LINK_6D, 0, (0, 0, 0))
        rd.add_task_space(1, TASK_LINK_ROTATION, 15, (0, 0, 0))
        rd.set_torque_limit(300.0)
        self.task_init = True
        be = rd.be
        self._flags = be.empty((self.num_envs, 2), "u1")
        self._out = None
        q0 = np.tile(np.asarray(Q_INIT, dtype=np.float64), (self.num_envs, 1))
        qpos0 = np.concatenate([q0[:, 0:3], q0[:, 39:40], q0[:, 3:6], q0[:, 6:39]], axis=1)
        self.UpdateKinematics(be.from_numpy(qpos0), be.from_numpy(np.zeros((self.num_envs, 39))), None)

    def reflectAction(self, action):
        """no-op, as in the reference (the body is commented out, src/pybind/rl_wbc_bridge.cpp:52-80)"""

    def UpdateKinematics(self, qpos, qvel, qacc=None):
        """qpos [B, 40] MuJoCo order (xyz, qw qx qy qz, joints), qvel [B, 39]; qacc is accepted and unused (the reference passes it
        to UpdateKinematics, which ignores it)."""
        self.rd_.update_kinematics_mujoco(qpos, qvel)

    def SetContact(self, left, right):
        """left / right: bool (all environments) or per-environment arrays / tensors [B].  Device tensors stay on the device: the
        [B, 2] uint8 flag array is assembled there (no D2H + H2D round trip on the RL step path)."""
        rd, be = self.rd_, self.rd_.be
        torch = getattr(be, "torch", None)
        on_dev = torch is not None and any(isinstance(v, torch.Tensor) and v.is_cuda for v in (left, right))
        if on_dev:
            cols = []
            for v in (left, right):
                if isinstance(v, torch.Tensor):
                    cols.append(v.to(device=be.device).reshape(self.num_envs).ne(0).to(torch.uint8))
                else:
                    cols.append(torch.full((self.num_envs,), 1 if v else 0, dtype=torch.uint8, device=be.device))
            self._flags = torch.stack(cols, dim=1).contiguous()
        else:
            fl = np.empty((self.num_envs, 2), dtype=np.uint8)
            for k, v in enumerate((left, right)):
                fl[:, k] = np.asarray(be.to_numpy(v) if hasattr(v, "shape") and not isinstance(v, np.ndarray) else v).astype(np.uint8).reshape(-1)
            self._flags = be.from_numpy(fl)
        rd.set_contact(self._flags)
        rd.calc_contact_constraint()  # + CalcTaskSpace: done inside solve()

    def SetTaskSpace(self, heirarchy, f_star):
        self.rd_.set_task_space(int(heirarchy), f_star)

    def CalcTorque(self):
        stages = SOLVE_GRAV | SOLVE_TASK | ((SOLVE_HQP | SOLVE_REDIST) if self.hqp else SOLVE_REDIST_ANALYTIC)
        self.task_init = False
        self._out = self.rd_.solve(stages, self._out)
        return self._out[3]

    def getTorqueCommand(self):
        """float32 [B, 33] device tensor: torque_grav_ + torque_task_ + torque_contact_"""
        return self.rd_.get_torque_command()

    def Reset(self):
        self.task_init = True


class RlWBCBridge:
    """Single-environment class with the reference's exact Python signature: numpy in, list[float] out."""

    def __init__(self, env_id=0, device=0, urdf_path=None, _backend=None):
        self.env_id = env_id
        self._b = RlWBCBridgeBatch(1, device, urdf_path, _backend=_backend)

    def reflectAction(self, action):
        pass

    def _dev(self, a, n):
        a = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(1, n))
        return self._b.rd_.be.from_numpy(a)

    def UpdateKinematics(self, qpos, qvel, qacc):
        self._b.UpdateKinematics(self._dev(qpos, 40), self._dev(qvel, 39), None)

    def SetContact(self, left, right):
        self._b.SetContact(bool(left), bool(right))

    def SetTaskSpace(self, heirarchy, f_star):
        f = np.asarray(f_star, dtype=np.float64).reshape(1, -1)
        self._b.SetTaskSpace(heirarchy, self._b.rd_.be.from_numpy(np.ascontiguousarray(f)))

    def CalcTorque(self):
        self._b.CalcTorque()

    def getTorqueCommand(self):
        return [float(x) for x in self._b.rd_.be.to_numpy(self._b.getTorqueCommand())[0]]

    def Reset(self):
        self._b.Reset()
