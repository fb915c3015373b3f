"""Small pass over every kernel of libphaserot_cuda for compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
Sizes are minimal (the tool slows kernels down 10-100x); every path is still the product path: digital and
true-peak sweeps (normal, brute force and dense mode), host PCM ingest (16 / 24 / 32 bit), shards, a device group,
CLI render, plugin small calls and a bulk call."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402
from phaserotate.lv2_b200 import capi  # noqa: E402

x = O.programme(48000, 1.2, 2)
q16 = np.clip(np.round(x * 32768.0), -32768, 32767).astype(np.int16)
q32 = q16.astype(np.int32) << 16
q24 = np.clip(np.round(x * 8388608.0), -8388608, 8388607).astype(np.int32)
b24 = np.stack([(q24 >> s) & 0xff for s in (0, 8, 16)], axis=-1).astype(np.uint8).reshape(-1)
with capi.Phaserot(n_channels=2, blksiz=8192) as h:
    h.sweep(x)
    a = h.peaks()
    for pcm in (q16, q32, b24):
        h.reset()
        h.sweep_pcm(pcm)
        h.peaks()
    h.reset()
    h.sweep(x, -3, 4, 1)
    h.peaks()
    al = h.shard_align()
    h.reset()
    h.sweep_shard(np.ascontiguousarray(x[:al]), al, None, True, False)
    h.peaks()
    y = h.render(x, [37, 181], 1)
    blk = x[:8192].copy()  # phaserot_apply works in place
    h.reset()
    h.apply(blk, [37, 181])
with capi.Phaserot(n_channels=2, blksiz=8192, flags=capi.FLAG_NO_PRUNE) as h:
    h.sweep(x)
    b = h.peaks()
    if not np.array_equal(b, a):
        d = np.argwhere(a != b)
        print("pruned != brute force at", d[:10].tolist(), a[a != b][:10], b[a != b][:10])
        raise SystemExit(1)
with capi.Phaserot(n_channels=2, blksiz=8192, oversample=4) as h:
    h.sweep(x)
    h.peaks()
with capi.Phaserot(n_channels=3, blksiz=32768) as h:
    h.sweep(O.harmonic(192000, 0.3, 3))
    h.peaks()
# dense mode: a sine long enough to overflow the 1 M-point list
t = np.arange(int(100 * 48000)) / 48000.0
s = np.stack([0.5 * np.sin(2 * np.pi * 440 * t), 0.5 * np.sin(2 * np.pi * 440 * t + 1.0)], axis=1).astype(np.float32)
with capi.Phaserot(n_channels=2, blksiz=8192, subsample=4) as h:
    h.sweep(s)
    h.peaks()
    print("dense repeats", h.stats()["dense_repeats"])
    # the handle is dense now: two tones go through the walk beyond the first window (sector hopping, wide list)
    tt = np.stack([0.5 * np.sin(2 * np.pi * 110 * t + c) + 0.25 * np.sin(2 * np.pi * 1760.3 * t) for c in (0.0, 1.0)], axis=1).astype(np.float32)
    h.reset()
    h.sweep(tt)
    h.peaks()
with capi.Phaserot(n_channels=1, blksiz=8192, subsample=50) as h:   # 0.02 degree grid: tables beyond shared memory, sector-window kernel
    m = np.ascontiguousarray(np.concatenate([s[:, :1], s[:, :1], s[:, :1]]))
    h.sweep(m)
    h.peaks()
    print("dense repeats (0.02 degree grid)", h.stats()["dense_repeats"])
with capi.PhaserotGroup(2, [0, 0], n_channels=2, blksiz=8192) as g:
    g.sweep(x)
    assert np.array_equal(g.peaks(), a)
with capi.Phaserot(mode=capi.MODE_PLUGIN, n_channels=2, sample_rate=48000.0) as h:
    for _ in range(6):
        h.process_levels(x[:1024].T.copy(), [90.0, -45.0])
    h.process(x[:40000].T.copy(), [90.0, -45.0])
print("sanitize_smoke: all paths ran")
