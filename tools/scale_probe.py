"""Where does the multi-GPU step spend its time?  torchrun --nproc-per-node N tools/scale_probe.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from phaserotate.lv2_b200 import capi  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
torch.cuda.set_stream(torch.cuda.Stream(device=dev))  # an explicit stream: handle 0 means "private stream" to phaserot_set_stream
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
frames = int(3600 * bench.SR); frames -= frames % (32768 - bench.BLKSIZ)
n_chunks = (frames + bench.GEN_CHUNK - 1) // bench.GEN_CHUNK
x = torch.cat([bench.gen_chunk_torch(torch, rank * n_chunks + k, dev) for k in range(n_chunks)])[:frames].contiguous()
hist = x[:bench.BLKSIZ].clone()
h = capi.Phaserot(n_channels=2, blksiz=bench.BLKSIZ, subsample=10, device=local)
h.set_stream(torch.cuda.current_stream().cuda_stream)
dummy = torch.zeros(3600, device=dev)


def table():
    ptr, nc, na = h.pending_table()

    class _D:
        __cuda_array_interface__ = {"shape": (nc * na + nc + 1,), "typestr": "<f4", "data": (ptr, False), "version": 2}
    return torch.as_tensor(_D(), device=dev)


def v_plain():
    h.reset(); h.sweep_device(x.data_ptr(), frames); h.peaks()


def v_shard():
    h.reset(); h.sweep_shard_device(x.data_ptr(), frames, hist.data_ptr() if rank else None, rank == 0, rank == world - 1); h.peaks()


def v_shard_ar():
    h.reset(); h.sweep_shard_device(x.data_ptr(), frames, hist.data_ptr() if rank else None, rank == 0, rank == world - 1)
    if world > 1:
        dist.all_reduce(table(), op=dist.ReduceOp.MAX)
    h.peaks()


def v_plain_dummy_ar():
    h.reset(); h.sweep_device(x.data_ptr(), frames)
    if world > 1:
        dist.all_reduce(dummy, op=dist.ReduceOp.MAX)
    h.peaks()


def v_ar_only():
    if world > 1:
        dist.all_reduce(dummy, op=dist.ReduceOp.MAX)
    torch.cuda.synchronize()


for name, fn in [("plain", v_plain), ("shard", v_shard), ("shard+allreduce", v_shard_ar), ("plain+dummy allreduce", v_plain_dummy_ar), ("allreduce only", v_ar_only)]:
    for _ in range(3):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 20 * 1e3
    print(f"rank {rank} {name:24s} {dt:.3f} ms/step", flush=True)
    if world > 1:
        dist.barrier()
