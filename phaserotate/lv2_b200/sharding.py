"""Sample-range sharding of one analysis pass across ranks (one process per GPU).

A per-angle peak is a maximum over samples, so a stream can be cut into shards,
each rank sweeps its shard (with one block of history in front of it), and the
per-rank tables combine with an element-wise max — the only collective on the
path (NCCL max all-reduce of [channels x angles] floats).  Cutting on multiples
of `align` frames (phaserot_shard_align) keeps the result bit-identical to a
single pass.
"""
import numpy as np


def plan_shards(n_frames, world, align):
    """[(start, n)] per rank: contiguous, every start a multiple of `align`; the last
    non-empty shard takes the remainder.  Ranks beyond the data get (n_frames, 0)."""
    if world < 1 or align < 1:
        raise ValueError("world and align must be >= 1")
    units = -(-n_frames // align)  # ceil: the tail unit may be short
    base, extra = divmod(units, world)
    plan, start = [], 0
    for r in range(world):
        n = min((base + (1 if r < extra else 0)) * align, n_frames - start)
        plan.append((start, n))
        start += n
    assert start == n_frames
    return plan


def shard_flags(plan, rank):
    """(first, last) for `rank`: does its shard start / end the stream?"""
    start, n = plan[rank]
    total = plan[-1][0] + plan[-1][1]
    owner_of_end = max([r for r, (_, nn) in enumerate(plan) if nn > 0], default=0)
    return start == 0 and (n > 0 or rank == 0), rank == owner_of_end


def sharded_sweep(compute_shard, fetch_history, n_frames, world, rank, align, table_shape, all_reduce_max):
    """Run this rank's shard and combine the tables of all ranks.

    compute_shard(start, n, hist, first, last) -> np.float32 array of table_shape
    fetch_history(start) -> the blksiz frames before `start` ([blksiz, channels])
    all_reduce_max(table) -> element-wise max over ranks
    A rank without data contributes zeros (peaks are >= 0) but still joins the collective.
    """
    plan = plan_shards(n_frames, world, align)
    start, n = plan[rank]
    first, last = shard_flags(plan, rank)
    if n == 0 and not last:
        table = np.zeros(table_shape, np.float32)
    else:
        hist = None if start == 0 else fetch_history(start)
        table = np.ascontiguousarray(compute_shard(start, n, hist, first, last), np.float32)
    return all_reduce_max(table)
