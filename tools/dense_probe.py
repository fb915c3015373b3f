"""Dense mode under the magnifier: listed points / evaluations per listed point / kernel times for a pure sine or the two tones of config 1
(python tools/dense_probe.py seconds [sine|two_sine]; PHASEROT_DEBUG=1 prints the counters)."""
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
torch.cuda.set_stream(torch.cuda.Stream(device=dev))
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 600.0
material = sys.argv[2] if len(sys.argv) > 2 else "sine"          # sine | two_sine
frames = int(secs * bench.SR)
frames -= frames % (32768 - bench.BLKSIZ)
parts = [(0.5, 440.0, [0.0, 1.0])] if material == "sine" else [(0.5, 110.0, [0.0, 1.0]), (0.25, 1760.3, [0.0, 0.0])]
x = IL.tone_chunks(torch, dev, frames, bench.SR, parts).contiguous()
r, _ = IL.run_leg(torch, capi, x, frames, bench.BLKSIZ, 10, 0)
print(r)
with capi.Phaserot(n_channels=2, blksiz=bench.BLKSIZ, subsample=10) as h:
    h.set_stream(torch.cuda.current_stream().cuda_stream)
    for k in range(4):
        h.reset_stats()
        h.set_profiling(True)
        h.sweep_device(x.data_ptr(), frames)      # no reset in between: from the second pass on every threshold is the final peak
        pk = h.peaks()
        st = h.stats()
        print("pass", k, st, {n: (round(v["ms"], 3), v["launches"]) for n, v in h.kernel_times().items() if v["launches"]}, "peak min/max ch0", float(pk[0].min()), float(pk[0].max()))
