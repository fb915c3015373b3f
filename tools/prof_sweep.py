"""Small driver for ncu: a few device-resident sweeps (and optionally renders) of synthetic audio."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from phaserotate.lv2_b200 import capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=300.0)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--subsample", type=int, default=10)
ap.add_argument("--no-prune", action="store_true")
ap.add_argument("--render", action="store_true")
a = ap.parse_args()

dev = torch.device("cuda", 0)
frames = int(a.seconds * bench.SR)
frames -= frames % bench.BLKSIZ
n_chunks = (frames + bench.GEN_CHUNK - 1) // bench.GEN_CHUNK
x = torch.cat([bench.gen_chunk_torch(torch, k, dev) for k in range(n_chunks)])[:frames].contiguous()
torch.cuda.synchronize()
h = capi.Phaserot(n_channels=bench.CHANNELS, blksiz=bench.BLKSIZ, subsample=a.subsample, flags=capi.FLAG_NO_PRUNE if a.no_prune else 0)
for _ in range(a.steps):
    h.reset()
    h.sweep_device(x.data_ptr(), frames)
    pk = h.peaks()
print("argmin", pk.argmin(1), h.stats())
if a.render:
    out = torch.empty(((frames // bench.BLKSIZ + 1) * bench.BLKSIZ, bench.CHANNELS), device=dev)
    for _ in range(a.steps):
        h.render_device(x.data_ptr(), frames, [90, 180], 1, out.data_ptr())
    torch.cuda.synchronize()
    print("render ok", float(out.abs().max()))
