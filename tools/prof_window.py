"""ncu driver for dense mode: a few device-resident sweeps of a pure sine (every sample survives the radius filter)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

import bench  # noqa: E402
import input_legs as IL  # noqa: E402
from phaserotate.lv2_b200 import capi  # noqa: E402

dev = torch.device("cuda", 0)
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 600.0
frames = int(secs * bench.SR)
frames -= frames % (32768 - bench.BLKSIZ)
x = IL.tone_chunks(torch, dev, frames, bench.SR, [(0.5, 440.0, [0.0, 1.0])]).contiguous()
with capi.Phaserot(n_channels=2, blksiz=bench.BLKSIZ, subsample=10) as h:
    for _ in range(3):
        h.reset()
        h.sweep_device(x.data_ptr(), frames)
        h.peaks()
    print(h.stats())
