"""How the sweep's cost depends on the input (VERDICT r01 "the headline rate is a property of the input").

Device-resident stereo 48 kHz audio of `--seconds`, 0.1 degree grid, one line of JSON per material:
  programme      bench.py's synthetic programme (16 partials x AM + noise)
  two_sine       BASELINE config 1 material (110 Hz + 1760.3 Hz, phase offset between channels)
  sine440        constant envelope: every sample lies on the hull of the (x_d, H) point set
  programme_np   programme with PHASEROT_FLAG_NO_PRUNE (brute force: every sample at every angle)
bench.py imports `materials()` / `run_leg()` for its `inputs` object.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def tone_chunks(torch, dev, frames, sr, parts, chunk=1 << 22):
    """sum_k a_k sin(2 pi (f_k t / sr) + ph_kc), phases accumulated in float64 so that an hour is as clean as a second.
    parts: list of (amp, freq, [phase per channel in radians])."""
    C = len(parts[0][2])
    x = torch.empty((frames, C), device=dev, dtype=torch.float32)
    for lo in range(0, frames, chunk):
        hi = min(frames, lo + chunk)
        t = torch.arange(lo, hi, device=dev, dtype=torch.float64) / sr
        for c in range(C):
            acc = torch.zeros(hi - lo, device=dev, dtype=torch.float32)
            for amp, f, ph in parts:
                fr = torch.frac(t * float(f) + float(ph[c]) / (2.0 * np.pi))
                acc += float(amp) * torch.sin((fr * (2.0 * np.pi)).to(torch.float32))
            x[lo:hi, c] = acc
    return x


def materials(torch, dev, frames, sr, programme):
    """name -> (tensor [frames, 2] on dev, flags).  `programme` is the caller's programme tensor (>= frames)."""
    from phaserotate.lv2_b200 import capi
    return {
        "programme": (programme[:frames], 0),
        "two_sine": (tone_chunks(torch, dev, frames, sr, [(0.5, 110.0, [0.0, 1.0]), (0.25, 1760.3, [0.0, 0.0])]), 0),
        "sine440": (tone_chunks(torch, dev, frames, sr, [(0.5, 440.0, [0.0, 1.0])]), 0),
        "programme_no_prune": (programme[:frames], capi.FLAG_NO_PRUNE),
    }


def run_leg(torch, capi, x, frames, blksiz, subsample, flags, steps=3, device=0):
    C = x.shape[1]
    h = capi.Phaserot(mode=capi.MODE_CLI, n_channels=C, blksiz=blksiz, subsample=subsample, device=device, flags=flags)
    h.set_stream(torch.cuda.current_stream().cuda_stream)
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    h.sweep_device(x.data_ptr(), frames)      # first call on a fresh handle: pays the dense-mode repeat if the material needs one
    pk = h.peaks()
    c1.record()
    torch.cuda.synchronize()
    first_ms, first_rep = c0.elapsed_time(c1), h.stats()["dense_repeats"]
    h.reset()
    h.sweep_device(x.data_ptr(), frames)
    pk = h.peaks()
    h.reset_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        h.reset()
        h.sweep_device(x.data_ptr(), frames)
        pk = h.peaks()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    st = h.stats()
    h.set_profiling(True)
    h.reset()
    h.sweep_device(x.data_ptr(), frames)
    h.peaks()
    kt = h.kernel_times()
    h.close()
    A = 180 * subsample
    return {"ms_per_step": ms, "value": float(frames) * C * A / (ms * 1e-3) / 1e9, "unit": "Gsample-angles/s",
            "survivor_fraction": st["points_evaluated"] / max(1, st["points_total"]),
            "kernels_ms": {k: round(v["ms"], 4) for k, v in kt.items() if v["launches"]},
            "launches_per_step": st["kernel_launches"] // steps,
            "first_call_ms": first_ms, "first_call_dense_repeats": first_rep, "dense_repeats_steady": st["dense_repeats"]}, pk


if __name__ == "__main__":
    import torch
    import bench
    from phaserotate.lv2_b200 import capi
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=600.0)
    ap.add_argument("--subsample", type=int, default=10)
    ap.add_argument("--check", action="store_true", help="compare every pruned table with brute force (bit for bit)")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))  # an explicit stream: handle 0 means "private stream" to phaserot_set_stream
    frames = int(a.seconds * bench.SR)
    frames -= frames % (32768 - bench.BLKSIZ)
    n_chunks = (frames + bench.GEN_CHUNK - 1) // bench.GEN_CHUNK
    prog = torch.cat([bench.gen_chunk_torch(torch, k, dev) for k in range(n_chunks)])[:frames].contiguous()
    for name, (x, flags) in materials(torch, dev, frames, bench.SR, prog).items():
        x = x.contiguous()
        r, pk = run_leg(torch, capi, x, frames, bench.BLKSIZ, a.subsample, flags)
        r["material"] = name
        r["seconds"] = frames / bench.SR
        if a.check and not flags:
            _, pb = run_leg(torch, capi, x, frames, bench.BLKSIZ, a.subsample, capi.FLAG_NO_PRUNE, steps=1)
            r["equals_brute_force"] = bool(np.array_equal(pk, pb))
        print(json.dumps(r), flush=True)
