"""RlWBCBridge / RlWBCBridgeBatch (reference src/pybind/rl_wbc_bridge.cpp) over the emulation build: same call sequence as the
reference's Python users, checked against the oracle running the calls the reference bridge makes (hqp = false by the
reference's own argument mix-up: closed-form task torque + analytic two-foot redistribution)."""
import os

import numpy as np

from libdwbc_b200 import RlWBCBridge, RlWBCBridgeBatch
from oracle import oracle as orc
from tests.common import *
from tests.emul.emul_backend import EmulBackend


def _oracle_cmd(q, qd, flags, f0, f1):
    r = orc.OracleRobot(URDF)
    r.add_contact(LEFT_FOOT, FOOT_POINT, (0, 0, 1), 0.15, 0.075)
    r.add_contact(RIGHT_FOOT, FOOT_POINT, (0, 0, 1), 0.15, 0.075)
    r.add_task(0, orc.TASK_LINK_6D, PELVIS)
    r.add_task(1, orc.TASK_LINK_ROTATION, UPPERBODY)
    r.set_torque_limit(300.0)
    r.update_kinematics(q, qd)
    r.set_contact(*flags)
    assert r.calc_contact_constraint() == 1
    r.set_task(0, f0)
    r.set_task(1, f1)
    r.calc_grav()
    r.update_task_space()
    r.calc_task_torque(False)
    r.calc_redistribute(False)
    return (r.get("torque_grav")[:, 0] + r.get("torque_task")[:, 0] + r.get("torque_contact")[:, 0]).astype(np.float32)


def test_bridge_batch_and_single():
    B = 3
    q, qd, rng = synth_states(B, 20252)
    qpos = np.ascontiguousarray(np.concatenate([q[:, 0:3], q[:, 39:40], q[:, 3:6], q[:, 6:39]], axis=1))
    f0, f1 = 0.3 * rng.standard_normal((B, 6)), 0.3 * rng.standard_normal((B, 3))
    br = RlWBCBridgeBatch(B, urdf_path=URDF, _backend=EmulBackend())
    br.UpdateKinematics(qpos, np.ascontiguousarray(qd), None)
    br.SetContact(True, True)
    br.SetTaskSpace(0, np.ascontiguousarray(f0))
    br.SetTaskSpace(1, np.ascontiguousarray(f1))
    st = br.CalcTorque()
    cmd = br.getTorqueCommand()
    assert cmd.dtype == np.float32 and cmd.shape == (B, 33)
    for b in range(B):
        want = _oracle_cmd(q[b], qd[b], (1, 1), f0[b], f1[b])
        assert np.abs(cmd[b] - want).max() <= 1e-5 * max(1.0, np.abs(want).max()), b
    one = RlWBCBridge(0, urdf_path=URDF, _backend=EmulBackend())
    one.UpdateKinematics(qpos[1], qd[1], np.zeros(39))
    one.SetContact(True, True)
    one.SetTaskSpace(0, f0[1])
    one.SetTaskSpace(1, f1[1])
    one.CalcTorque()
    out = one.getTorqueCommand()
    assert isinstance(out, list) and len(out) == 33 and np.array_equal(np.float32(out), cmd[1])
    one.Reset()


def run_pybind_module(lib, out_dir):
    """The compiled Python module `dwbc` (libdwbc_b200/csrc/pybind_dwbc.cpp; reference src/pybind/rl_wbc_bridge.cpp:148-158) linked against
    `lib`: RlWBCBridge(env_id) with the reference's exact call signatures and RlWBCBridgeBatch, against the oracle."""
    import importlib
    import sys

    import __graft_entry__ as g

    path = g.build_pybind(lib, out_dir)
    sys.path.insert(0, os.path.dirname(path))
    try:
        sys.modules.pop("dwbc", None)
        dwbc = importlib.import_module("dwbc")
    finally:
        sys.path.pop(0)
    B = 3
    q, qd, rng = synth_states(B, 20253)
    qpos = np.ascontiguousarray(np.concatenate([q[:, 0:3], q[:, 39:40], q[:, 3:6], q[:, 6:39]], axis=1))
    f0, f1 = 0.3 * rng.standard_normal((B, 6)), 0.3 * rng.standard_normal((B, 3))
    assert sorted(n for n in dir(dwbc.RlWBCBridge) if not n.startswith("_")) == sorted(
        ["reflectAction", "UpdateKinematics", "SetContact", "SetTaskSpace", "CalcTorque", "getTorqueCommand", "Reset"])
    one = dwbc.RlWBCBridge(0)
    one.reflectAction(np.zeros(13))
    for b in range(B):
        one.UpdateKinematics(qpos[b], qd[b], np.zeros(39))
        one.SetContact(True, True)
        one.SetTaskSpace(0, f0[b])
        one.SetTaskSpace(1, f1[b])
        one.CalcTorque()
        out = one.getTorqueCommand()
        want = _oracle_cmd(q[b], qd[b], (1, 1), f0[b], f1[b])
        assert isinstance(out, list) and len(out) == 33
        assert np.abs(np.float32(out) - want).max() <= 1e-5 * max(1.0, np.abs(want).max()), b
    one.Reset()
    bb = dwbc.RlWBCBridgeBatch(B, 0)
    bb.UpdateKinematics(qpos, qd, None)
    bb.SetContact(np.ones(B, dtype=np.uint8), np.ones(1, dtype=np.uint8))
    bb.SetTaskSpace(0, f0)
    bb.SetTaskSpace(1, f1)
    bb.CalcTorque()
    cmd = bb.getTorqueCommand()
    assert cmd.dtype == np.float32 and cmd.shape == (B, 33)
    for b in range(B):
        want = _oracle_cmd(q[b], qd[b], (1, 1), f0[b], f1[b])
        assert np.abs(cmd[b] - want).max() <= 1e-5 * max(1.0, np.abs(want).max()), b


def test_pybind_module_emulation():
    import os as _os

    from tests.emul.build_emul import build

    lib = build()
    run_pybind_module(lib, _os.path.dirname(lib))
