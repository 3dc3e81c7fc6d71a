"""dwbc_multi_* (one host batch over several GPUs, one host thread + stream + handle per GPU) through the CUDA emulation: the
"GPUs" are slices of the batch driven by the library's own worker threads; results must equal the single-handle call bit for bit."""
import ctypes as C

import numpy as np
import pytest

from libdwbc_b200 import CONTACT_6D, SOLVE_GRAV, SOLVE_HQP, SOLVE_REDIST, TASK_LINK_POSITION, TASK_LINK_ROTATION, RobotDataBatch, RobotDataMultiBatch
from libdwbc_b200._lib import declare
from tests.common import *
from tests.emul.build_emul import build
from tests.emul.emul_backend import EmulBackend


@pytest.mark.parametrize("ngpus,B", [(2, 3), (3, 4)])  # small batches: a state costs seconds in the emulation
def test_multi_equals_single(ngpus, B):
    lib = declare(C.CDLL(build()))
    q, qd, rng = synth_states(B, 20245)
    fs = synth_fstar(rng, B)
    fs[0][0, 1] = 4.0
    flags = np.ones((B, 2), dtype=np.uint8)
    flags[1] = [1, 0]
    stages = SOLVE_GRAV | SOLVE_HQP | SOLVE_REDIST
    mh = RobotDataMultiBatch(URDF, B, ngpus, _lib=lib)
    assert mh.ngpus == ngpus and sum(s["count"] for s in mh.shards()) == B
    mh.add_contact_constraint(LEFT_FOOT, CONTACT_6D, FOOT_POINT, (0, 0, 1), 0.15, 0.075)
    mh.add_contact_constraint(RIGHT_FOOT, CONTACT_6D, FOOT_POINT, (0, 0, 1), 0.15, 0.075)
    mh.add_task_space(0, TASK_LINK_POSITION, "COM")
    mh.add_task_space(1, TASK_LINK_ROTATION, "pelvis_link")
    mh.add_task_space(1, TASK_LINK_ROTATION, "upperbody_link")
    mh.add_custom_task_space(2, 21, posture_selector())
    mh.set_torque_limit(300.0)
    qp = mh.pinned((B, 40))
    qp[:] = q
    out = mh.cycle_host(stages, qp, qd, fs, flags)
    assert all(s["last_path"] == 1 for s in mh.shards())
    rd = configure_product(RobotDataBatch(URDF, B, _backend=EmulBackend()), 2)
    rd.set_task_space(2, np.zeros((B, 21)), np.ascontiguousarray(posture_selector()))
    ref = rd.cycle_host(stages, q, qd, fs, flags)
    for a, b in zip(out, ref):
        assert np.array_equal(a, b)
    assert len(mh.last_shard_ms()) == ngpus
    with pytest.raises(Exception):
        mh.cycle_host(stages, q[:-1], qd, fs, flags)
