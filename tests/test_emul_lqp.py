"""a10 LQP variant: kernel logic (emulation build, no GPU) vs the oracle's exact level-by-level solve.  The authoritative
comparison is tests/test_gpu_parity.py::test_gpu_lqp; PARITY UNPINNED w.r.t. the reference (OSQP, no goldens)."""
import numpy as np
import pytest

from libdwbc_b200 import RobotDataBatch
from oracle import oracle as orc
from tests.common import *
from tests.emul.emul_backend import EmulBackend


def lqp_oracle_states(q, qd, flags, fstars):
    res = []
    for b in range(q.shape[0]):
        r = configure_oracle(orc.OracleRobot(URDF), "lqp")
        r.update_kinematics(q[b], qd[b])
        r.set_contact(*flags[b])
        ok_cc = r.calc_contact_constraint()
        for l, f in enumerate(fstars):
            r.set_task(l, f[b])
        r.update_task_space()
        nl = r.configure_lqp()
        ok = r.calc_lqp()
        res.append((r, nl, ok if ok_cc == 1 else 0))
    return res


def check_lqp(rd, out, refs, flags, tol=1e-6):
    y, tau, st = out
    n = 39
    for b, (r, nl, ok) in enumerate(refs):
        c = 6 * int(sum(flags[b]))
        assert int(st[b]) == (nl if ok else st[b]) and (ok == 1) == (int(st[b]) == nl), (b, st[b], nl, ok)
        if not ok:
            continue
        for l in range(nl):
            yl = rd.get_lqp("y", l)[b]
            yr = r.get(f"lqp_y{l}")[:, 0]
            assert relerr(yl[:n + c], yr) < tol, (b, l, relerr(yl[:n + c], yr))
            assert np.all(yl[n + c:] == 0)
        v1 = rd.get_lqp("v1")[b]
        vr = r.get("lqp_v1")[:, 0]
        assert np.linalg.norm(v1[:len(vr)] - vr) < tol * max(1.0, np.linalg.norm(vr)), b
        assert relerr(tau[b], r.get("lqp_torque")[:, 0]) < tol, b
        assert relerr(y[b][:n + c], r.get(f"lqp_y{nl - 1}")[:, 0]) < tol


@pytest.mark.parametrize("hard", [False, True])
def test_emul_lqp_vs_oracle(hard):
    B = 4
    q, qd, rng = synth_states(B, 20243)
    fs = lqp_fstars(rng, B, hard)
    flags = np.ones((B, 2), dtype=np.uint8)
    flags[2] = [1, 0]
    flags[3] = [0, 1]
    rd = configure_product(RobotDataBatch(URDF, B, _backend=EmulBackend()), "lqp")
    rd.update_kinematics(np.ascontiguousarray(q), np.ascontiguousarray(qd))
    rd.set_contact(flags)
    rd.calc_contact_constraint()
    for l, f in enumerate(fs):
        rd.set_task_space(l, np.ascontiguousarray(f))
    out = rd.calc_control_torque_lqp()
    refs = lqp_oracle_states(q, qd, flags, fs)
    check_lqp(rd, out, refs, flags)
