"""world_size-2 test of the multi-GPU path's host logic on CPU (gloo): shard
planning, history hand-over and the max all-reduce that combines per-rank peak
tables.  The per-shard compute is the oracle here (there is no GPU); on the GPU
box the same helper drives phaserot_sweep_shard_device (bench.py,
tests/test_gpu_parity.py)."""
import os
import socket

import numpy as np
import pytest

import oracle_lib as O
from phaserotate.lv2_b200 import sharding


def test_plan_shards():
    al = 24576
    for n, w in [(172800000, 8), (1000, 4), (0, 2), (24576 * 3, 2), (24576 * 3 + 5, 3), (100, 1)]:
        plan = sharding.plan_shards(n, w, al)
        assert len(plan) == w and plan[0][0] == 0
        assert sum(nn for _, nn in plan) == n
        for (s0, n0), (s1, _) in zip(plan, plan[1:]):
            assert s1 == s0 + n0
        for r, (s, nn) in enumerate(plan):
            if nn:
                assert s % al == 0
            first, last = sharding.shard_flags(plan, r)
            assert first == (s == 0 and (nn > 0 or r == 0))
        assert sum(sharding.shard_flags(plan, r)[1] for r in range(w)) == 1  # exactly one rank ends the stream
    sizes = [nn for _, nn in sharding.plan_shards(172800000, 8, al)]
    assert max(sizes) - min(sizes) <= al


def _worker(rank, world, port, n_frames, L, align, out_path):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x = O.harmonic(48000, n_frames / 48000.0, 2)[:n_frames]

    def compute(start, n, hist, first, last):
        return O.oracle_analyze_shard(x[start:start + n], L, hist, first, last)

    def reduce_max(table):
        t = torch.from_numpy(table)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.numpy()

    table = sharding.sharded_sweep(compute, lambda s: x[s - L:s], n_frames, world, rank, align, (2, 360), reduce_max)
    if rank == 0:
        np.save(out_path, table)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_frames,L", [(40000, 2048), (3 * 4096 + 17, 1024), (1500, 1024)])
def test_two_rank_shards_combine_to_single_pass(tmp_path, oracle_built, n_frames, L):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "table.npy")
    mp.spawn(_worker, args=(2, port, n_frames, L, 2 * L, out), nprocs=2, join=True)
    combined = np.load(out)
    x = O.harmonic(48000, n_frames / 48000.0, 2)[:n_frames]
    single = O.oracle_analyze(x, L)
    # the oracle convolves exactly, so the cut position cannot matter: bit-identical
    assert np.array_equal(combined, single)
