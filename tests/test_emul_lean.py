"""The one-kernel cycle (DWBC_OUT_TORQUES, libdwbc_b200/csrc/kernels_lean.cuh) against the generic kernels and the oracle,
through the CUDA emulation (test infrastructure).  The `-m gpu` twins are in tests/test_gpu_parity.py."""
import numpy as np
import pytest

from libdwbc_b200 import OUT_ALL, OUT_TORQUES, SOLVE_GRAV, SOLVE_HQP, SOLVE_REDIST, SOLVE_TASK, DwbcError, RobotDataBatch
from oracle import oracle as orc
from tests.common import *
from tests.emul.emul_backend import EmulBackend


def run(config, q, qd, fstars, flags, stages, outputs, ncontacts=2):
    B = q.shape[0]
    rd = configure_product(RobotDataBatch(URDF, B, _backend=EmulBackend()), config, ncontacts)
    rd.set_outputs(outputs)
    rd.update_kinematics(np.ascontiguousarray(q), np.ascontiguousarray(qd))
    rd.set_contact(np.ascontiguousarray(flags, dtype=np.uint8))
    rd.calc_contact_constraint()
    for l, f in enumerate(fstars):
        rd.set_task_space(l, np.ascontiguousarray(f), np.ascontiguousarray(posture_selector()) if (config == 2 and l == 2) else None)
    out = [np.array(x) for x in rd.solve(stages)]
    return rd, out


def test_lean_config2_vs_generic_and_oracle():
    """3-level HQP + redistribution; double and single support; QP-active and inactive states; a state without contact."""
    B = 5
    q, qd, rng = synth_states(B, 20242)
    fs = synth_fstar(rng, B)
    fs[0][0, 1] = 4.0
    flags = np.ones((B, 2), dtype=np.uint8)
    flags[2] = [1, 0]
    flags[3] = [0, 1]
    flags[4] = [0, 0]
    stages = SOLVE_GRAV | SOLVE_HQP | SOLVE_REDIST
    rdl, (tg, tt, tc, st) = run(2, q, qd, fs, flags, stages, OUT_TORQUES)
    assert rdl.last_path() == 1
    rdg, (tg0, tt0, tc0, st0) = run(2, q, qd, fs, flags, stages, OUT_ALL)
    assert rdg.last_path() == 0
    assert np.array_equal(st & 7, st0 & 7)
    assert st[4] == 0 and not tg[4].any() and not tt[4].any() and not tc[4].any()
    for b in range(B):
        assert relerr(tg[b], tg0[b]) < 1e-10 or not tg0[b].any()
        assert np.abs(tt[b] - tt0[b]).max() < 1e-8 * max(1.0, np.abs(tt0[b]).max())
        assert np.abs(tc[b] - tc0[b]).max() < 1e-8 * max(1.0, np.abs(tc0[b]).max())
    for l in (0, 1, 2, -1):
        assert rdl.active_set(l) == rdg.active_set(l)
    o = configure_oracle(orc.OracleRobot(URDF), 2).run_batch(q, qd, fs, flags)
    assert np.array_equal(st[:4] & 7, o["status"][:4] & 7)  # (the state without contact is the product's own edge case: status 0)
    for b in range(B - 1):
        if (o["status"][b] & 7) != 7:
            continue
        assert relerr(tg[b], o["torque_grav"][b]) < 1e-9
        assert relerr(tt[b], o["torque_task"][b]) < 1e-6
        assert np.abs(tc[b] - o["torque_contact"][b]).max() < 1e-6 * max(1.0, np.abs(o["torque_contact"][b]).max())
    # only torque-group fields can be read back
    assert np.abs(rdl.get("f_star_qp", 0) - rdg.get("f_star_qp", 0)).max() < 1e-8
    with pytest.raises(DwbcError):
        rdl.get("A_inv")


@pytest.mark.parametrize("case", [1, 2])
def test_lean_golden(case):
    import os

    from oracle.urdf_model import read_golden
    q = np.array([Q_CASE[case]], dtype=np.float64)
    flags = np.array([[1, 1, 0, 0]], dtype=np.uint8)
    fs = [np.array([FSTAR_CASE[case][0]]), np.array([FSTAR_CASE[case][1]])]
    rd, (tg, tt, tc, st) = run("golden", q, np.zeros((1, 39)), fs, flags, SOLVE_GRAV | SOLVE_HQP | SOLVE_REDIST, OUT_TORQUES, 4)
    assert rd.last_path() == 1 and st[0] == 7
    g = lambda n: read_golden(os.path.join(GOLDEN, "cases", str(case), n))
    assert relerr(tg[0], g("torque_grav_")[:, 0]) < 1e-9
    assert relerr(tt[0], g("torque_task_")[:, 0]) < 1e-9
    assert relerr(tc[0], g("torque_contact_")[:, 0]) < (1e-9 if case == 1 else 1e-4)
    if case == 1:
        assert rd.active_set(0)[0] == [66, 67, 68, 69, 70, 71, 72, 73, 74, 75, 78]
        assert rd.active_set(1)[0] == [67, 68, 71, 72, 75, 78]


def test_lean_config1_and_config0():
    """closed-form task torque (config 1) and gravity compensation only (config 0) through the one-kernel cycle"""
    B = 3
    q, qd, rng = synth_states(B, 20241)
    fs = synth_fstar(rng, B, dofs=(3, 3), scales=(0.5, 1.0), hard_frac=0)
    flags = np.ones((B, 2), dtype=np.uint8)
    flags[2] = [1, 0]
    rdl, (tg, tt, tc, st) = run(1, q, qd, fs, flags, SOLVE_GRAV | SOLVE_TASK, OUT_TORQUES)
    rdg, (tg0, tt0, tc0, st0) = run(1, q, qd, fs, flags, SOLVE_GRAV | SOLVE_TASK, OUT_ALL)
    assert rdl.last_path() == 1 and np.array_equal(st & 7, st0 & 7)
    assert relerr(tg, tg0) < 1e-10 and relerr(tt, tt0) < 1e-9 and not tc.any()
    # the torque path never allocates the public-field workspace (dwbc_batch_workspace_bytes: lazy allocation)
    assert 0 < rdl.workspace_bytes() < 40 * 1024 * B and rdg.workspace_bytes() > 2 * rdl.workspace_bytes()
    rdl, (tg, tt, tc, st) = run(1, q, qd, fs, flags, SOLVE_GRAV, OUT_TORQUES)
    assert rdl.last_path() == 1 and np.array_equal(st & 7, np.ones(B, dtype=np.int32))
    assert relerr(tg, tg0) < 1e-10 and not tt.any()


def test_lean_cycle_host_matches_staged():
    B = 3
    q, qd, rng = synth_states(B, 20243)
    fs = synth_fstar(rng, B)
    flags = np.ones((B, 2), dtype=np.uint8)
    flags[1] = [0, 1]
    stages = SOLVE_GRAV | SOLVE_HQP | SOLVE_REDIST
    _, (tg, tt, tc, st) = run(2, q, qd, fs, flags, stages, OUT_TORQUES)
    rd = configure_product(RobotDataBatch(URDF, B, _backend=EmulBackend()), 2)
    rd.set_task_space(2, np.zeros((B, 21)), np.ascontiguousarray(posture_selector()))
    rd.set_host_slices(2)
    out = rd.cycle_host(stages, q, qd, fs, flags)
    assert rd.last_path() == 1
    for a, b in zip(out, (tg, tt, tc, st)):
        assert np.array_equal(a, b)
