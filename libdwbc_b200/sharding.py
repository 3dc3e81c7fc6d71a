"""Multi-GPU host logic: the batch of independent robot states is cut into contiguous slices, one per rank
(SURVEY.md section 8(e)).  There is NO collective on the data path -- every rank runs its own `RobotDataBatch`
on its own slice; `gather_states` is the optional, off-path collection of per-state results on every rank
(torch.distributed all_gather over whatever backend the process group uses: NCCL on GPUs, gloo in the CPU tests).

The reference has no multi-device story (one `RobotData` per thread, reference tests/test_thread.cpp:174-176);
the slices play the role of its per-thread instances.
"""
from __future__ import annotations


def shard_bounds(total: int, world: int, rank: int) -> tuple[int, int]:
    """[lo, hi) of rank's contiguous slice: the first `total % world` ranks hold one state more."""
    if world <= 0 or not (0 <= rank < world) or total < 0:
        raise ValueError(f"shard_bounds: bad arguments total={total} world={world} rank={rank}")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(total: int, world: int) -> list[int]:
    return [shard_bounds(total, world, r)[1] - shard_bounds(total, world, r)[0] for r in range(world)]


def shard(array, world: int, rank: int):
    """rank's slice of a per-state array (first axis = state)."""
    lo, hi = shard_bounds(len(array), world, rank)
    return array[lo:hi]


def gather_states(local, total: int, group=None):
    """All-gathers per-state results (torch tensor, first axis = local states) into the global order.
    Off the hot path; ragged slices are padded to the largest one for the collective."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    sizes = shard_sizes(total, world)
    if local.shape[0] != sizes[dist.get_rank(group)]:
        raise ValueError(f"gather_states: local slice has {local.shape[0]} states, expected {sizes[dist.get_rank(group)]}")
    cap = max(sizes)
    pad = torch.zeros((cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)
