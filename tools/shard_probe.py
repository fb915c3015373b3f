"""Per-shard cost of the 1-hour sweep for the audio each rank of an 8-GPU run gets (single GPU)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from phaserotate.lv2_b200 import capi  # noqa: E402

dev = torch.device("cuda", 0)
frames = int(3600 * bench.SR); frames -= frames % (32768 - bench.BLKSIZ)
n_chunks = (frames + bench.GEN_CHUNK - 1) // bench.GEN_CHUNK
h = capi.Phaserot(n_channels=2, blksiz=bench.BLKSIZ, subsample=10, device=0)
for r in range(int(sys.argv[1]) if len(sys.argv) > 1 else 8):
    x = torch.cat([bench.gen_chunk_torch(torch, r * n_chunks + k, dev) for k in range(n_chunks)])[:frames].contiguous()
    for _ in range(2):
        h.reset(); h.sweep_device(x.data_ptr(), frames); pk = h.peaks()
    h.reset_stats()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        h.reset(); h.sweep_device(x.data_ptr(), frames); pk = h.peaks()
    dt = (time.perf_counter() - t0) / 5 * 1e3
    st = h.stats()
    h.set_profiling(True); h.reset(); h.sweep_device(x.data_ptr(), frames); h.peaks(); kt = h.kernel_times(); h.set_profiling(False)
    print(f"shard {r}: {dt:.3f} ms/step  survivors/step {st['points_evaluated'] // 5}  min peak {pk[:, 1:].min(1)} max {pk.max(1)}  "
          f"fft {kt['fftconv_filter']['ms']:.3f} sweep {kt['sweep']['ms']:.3f}", flush=True)
    del x
