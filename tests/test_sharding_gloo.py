"""Multi-rank host logic on CPU (gloo, world_size 2): contiguous slices of independent states, one handle per rank,
no data-path collective; the off-path gather reproduces the single-process result bit for bit.  The kernels run
through the host emulation build (tests/emul, test infrastructure) because this container has no GPU."""
import os
import socket

import numpy as np
import pytest

from libdwbc_b200.sharding import shard_bounds, shard_sizes


def test_shard_bounds_cover_and_order():
    for total in (0, 1, 7, 65536, 1048576 + 3):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(total, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == total
            assert all(b[r][1] == b[r + 1][0] for r in range(world - 1))
            s = shard_sizes(total, world)
            assert sum(s) == total and max(s) - min(s) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, outdir):
    import torch
    import torch.distributed as dist

    from libdwbc_b200 import SOLVE_GRAV, SOLVE_HQP, SOLVE_REDIST, RobotDataBatch
    from libdwbc_b200.sharding import gather_states, shard
    from tests.common import URDF, configure_product, posture_selector, synth_fstar, synth_states
    from tests.emul.emul_backend import EmulBackend

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    q, qd, rng = synth_states(total, 20242)
    fs = synth_fstar(rng, total)
    flags = np.ones((total, 2), dtype=np.uint8)
    flags[1] = [1, 0]
    ql, qdl, fl = shard(q, world, rank), shard(qd, world, rank), shard(flags, world, rank)
    fsl = [shard(f, world, rank) for f in fs]
    rd = configure_product(RobotDataBatch(URDF, len(ql), _backend=EmulBackend()), 2)
    rd.update_kinematics(np.ascontiguousarray(ql), np.ascontiguousarray(qdl))
    rd.set_contact(np.ascontiguousarray(fl))
    rd.calc_contact_constraint()
    for l, f in enumerate(fsl):
        rd.set_task_space(l, np.ascontiguousarray(f), np.ascontiguousarray(posture_selector()) if l == 2 else None)
    tg, tt, tc, st = rd.solve(SOLVE_GRAV | SOLVE_HQP | SOLVE_REDIST)
    local = torch.from_numpy(np.concatenate([tg, tt, tc, st[:, None].astype(np.float64)], axis=1))
    full = gather_states(local, total)
    dist.barrier()
    if rank == 0:
        np.save(os.path.join(outdir, "gathered.npy"), full.numpy())
    dist.destroy_process_group()


def test_two_rank_slices_match_single_process(tmp_path):
    import torch.multiprocessing as mp

    from libdwbc_b200 import SOLVE_GRAV, SOLVE_HQP, SOLVE_REDIST, RobotDataBatch
    from tests.common import URDF, configure_product, posture_selector, synth_fstar, synth_states
    from tests.emul.build_emul import build
    from tests.emul.emul_backend import EmulBackend

    build()  # compile once, before the ranks race for it
    total, world = 3, 2  # ragged: 2 + 1 states (a state costs seconds in the emulation)
    mp.spawn(_worker, args=(world, _free_port(), total, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "gathered.npy")

    q, qd, rng = synth_states(total, 20242)
    fs = synth_fstar(rng, total)
    flags = np.ones((total, 2), dtype=np.uint8)
    flags[1] = [1, 0]
    rd = configure_product(RobotDataBatch(URDF, total, _backend=EmulBackend()), 2)
    rd.update_kinematics(q, qd)
    rd.set_contact(flags)
    rd.calc_contact_constraint()
    for l, f in enumerate(fs):
        rd.set_task_space(l, np.ascontiguousarray(f), np.ascontiguousarray(posture_selector()) if l == 2 else None)
    tg, tt, tc, st = rd.solve(SOLVE_GRAV | SOLVE_HQP | SOLVE_REDIST)
    want = np.concatenate([tg, tt, tc, st[:, None].astype(np.float64)], axis=1)
    assert got.shape == want.shape
    assert np.array_equal(got, want)  # same code on the same states: bit-exact, independent of the slicing
