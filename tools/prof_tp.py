"""Small driver for ncu: true-peak sweeps (4x) of synthetic audio, device resident."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from phaserotate.lv2_b200 import capi  # noqa: E402

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 600.0
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
frames = int(secs * bench.SR)
frames -= frames % (32768 - bench.BLKSIZ)
n_chunks = (frames + bench.GEN_CHUNK - 1) // bench.GEN_CHUNK
x = torch.cat([bench.gen_chunk_torch(torch, k, dev) for k in range(n_chunks)])[:frames].contiguous()
torch.cuda.synchronize()
h = capi.Phaserot(n_channels=bench.CHANNELS, blksiz=bench.BLKSIZ, subsample=10, oversample=4)
h.set_profiling(True)
for _ in range(steps):
    h.reset()
    h.sweep_device(x.data_ptr(), frames)
    pk = h.peaks()
print("argmin", pk.argmin(1), h.stats(), {k: v for k, v in h.kernel_times().items() if v["launches"]})
