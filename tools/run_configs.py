"""BASELINE.json's five configs at FULL size on one B200 (plus the reference CLI's wall time for
config 1): one JSON object per config on stdout.  A measurement aid next to bench.py - bench.py's
line stays the contract; this shows every named configuration running through the same C ABI.

    python tools/run_configs.py [1 2 3 4 5]
"""
import ctypes
import json
import os
import struct
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from phaserotate.lv2_b200 import build, capi  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_stream(torch.cuda.Stream(device=dev))  # an explicit stream: handle 0 means "private stream" to phaserot_set_stream
GEN = 1 << 21


def gen(frames, channels, sr, seed):
    """bench.py's programme recipe (16 partials x slow AM + noise) for any rate / channel count."""
    x = torch.empty((frames, channels), device=dev, dtype=torch.float32)
    rng = np.random.default_rng(seed)
    f = np.exp(rng.uniform(np.log(50.0), np.log(15000.0), (channels, 16)))
    ph = rng.uniform(0, 1.0, (channels, 16))
    amp = 50.0 / f
    amp /= amp.sum(axis=1, keepdims=True)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    for k0 in range(0, frames, GEN):
        n = min(GEN, frames - k0)
        t = torch.arange(k0, k0 + n, device=dev, dtype=torch.float64) / sr
        for c in range(channels):
            acc = torch.zeros(n, device=dev, dtype=torch.float32)
            for k in range(16):
                acc += float(amp[c, k]) * torch.sin(torch.frac(t * float(f[c, k]) + float(ph[c, k])).to(torch.float32) * (2.0 * np.pi))
            env = 0.6 + 0.4 * torch.sin((torch.frac(t * 0.37) * (2.0 * np.pi)).to(torch.float32) + float(c))
            x[k0:k0 + n, c] = 0.8 * acc * env + 0.02 * torch.randn(n, device=dev, dtype=torch.float32, generator=g)
    torch.cuda.synchronize()
    return x


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps


def write_wav_f32(path, x, sr):
    x = np.ascontiguousarray(x, "<f4")
    n, c = x.shape
    data = x.tobytes()
    with open(path, "wb") as fo:
        fo.write(b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 3, c, sr, sr * c * 4, c * 4, 32) + b"data" + struct.pack("<I", len(data)))
        fo.write(data)


def config1():
    sr, secs = 48000, 60
    t = np.arange(sr * secs, dtype=np.float64) / sr
    x = np.stack([0.5 * np.sin(2 * np.pi * 110 * t + p) + 0.25 * np.sin(2 * np.pi * 1760.3 * t) for p in (0.0, 1.0)], 1).astype(np.float32)
    wav = "/tmp/config1.wav"
    write_wav_f32(wav, x, sr)
    exe = os.path.join(build.BIN_DIR, "phase-rotate")
    ref = os.path.join(ROOT, "oracle", "_ref", "phase-rotate-f32")
    out = {"config": 1, "workload": "phase-rotate CLI, stereo 48 kHz 60 s two-sine float WAV, digital peak; whole process wall time incl. file read, CUDA context, H2D"}
    for tag, argv in (("s2_coarse_refine", ["-s", "2"]), ("s1_full_grid", ["-s", "1"])):
        subprocess.run([exe] + argv + [wav], capture_output=True)
        t0 = time.perf_counter()
        r = subprocess.run([exe] + argv + [wav], capture_output=True, text=True)
        dt = time.perf_counter() - t0
        o = {"b200_cli_wall_s": dt, "stdout": r.stdout.strip().splitlines()}
        if os.path.exists(ref):
            t0 = time.perf_counter()
            rr = subprocess.run([ref] + argv + [wav], capture_output=True, text=True)
            o["reference_cli_wall_s"] = time.perf_counter() - t0
            o["reference_note"] = "unmodified reference sources, stand-in float FFT, 2 threads (one per channel)"
            o["same_report"] = rr.stdout == r.stdout
        out[tag] = o
    # the analysis itself through the ABI, host buffer (pinned) -> table
    with capi.Phaserot(n_channels=2, blksiz=8192) as h:
        xp = torch.from_numpy(x).pin_memory()
        dt = timed(lambda: (h.reset(), h.sweep((xp.data_ptr(), x.shape[0])), h.peaks()), 20, 3)
    out["abi_sweep_host_to_table_ms"] = 1e3 * dt
    out["gsample_angles_per_s"] = x.size * 360 / dt / 1e9
    return out


def config2():
    sr, secs = 48000, 600
    rng = np.random.default_rng(42)
    w = rng.standard_normal(sr * secs).astype(np.float32)
    # Paul Kellet's economy pink filter
    b = np.zeros(3)
    x = np.empty_like(w)
    bb0 = bb1 = bb2 = 0.0
    # vectorised approximation: three one-pole sections
    from scipy.signal import lfilter
    x = (lfilter([0.0990460], [1, -0.99765], w) + lfilter([0.2965164], [1, -0.96300], w) + lfilter([1.0526913], [1, -0.57000], w) + 0.1848 * w).astype(np.float32)
    x *= 0.5 / np.abs(x).max()
    y = np.zeros_like(x)
    ang = np.array([90.0], np.float32)
    out = {"config": 2, "workload": "LV2 plugin run(), mono 48 kHz 600 s pink noise, angle port 90 deg from the first call (ramp included), host buffers"}
    with capi.Phaserot(mode=capi.MODE_PLUGIN, n_channels=1, sample_rate=48000.0) as hp:
        ins, outs = (ctypes.c_void_p * 1)(), (ctypes.c_void_p * 1)()
        blk = 1024
        n_calls = x.size // blk
        t0 = time.perf_counter()
        for k in range(n_calls):
            ins[0] = x.ctypes.data + 4 * k * blk
            outs[0] = y.ctypes.data + 4 * k * blk
            hp.process_raw(ins, outs, blk, ang)
        dt = time.perf_counter() - t0
        out["calls_1024"] = {"calls": n_calls, "wall_s": dt, "us_per_call": 1e6 * dt / n_calls, "msamples_per_s": n_calls * blk / dt / 1e6,
                             "realtime_factor": secs / dt}
        hp.reset()
        ins[0], outs[0] = x.ctypes.data, y.ctypes.data
        hp.process_raw(ins, outs, x.size, ang)
        hp.reset()
        t0 = time.perf_counter()
        hp.process_raw(ins, outs, x.size, ang)
        dt = time.perf_counter() - t0
        out["bulk_call"] = {"wall_s": dt, "msamples_per_s": x.size / dt / 1e6}
    return out


def config3():
    sr, secs, L, S = 96000, 3600, 16384, 10
    frames = sr * secs
    frames -= frames % L
    x = gen(frames, 2, sr, 43)
    out = {"config": 3, "workload": "CLI min-peak sweep with 4x oversampled true-peak, stereo 96 kHz 1 h synthetic programme, 0.1 deg (1800 angles), blksiz 16384, device resident"}
    for tag, os_ in (("true_peak_4x", 4), ("digital_peak", 0)):
        with capi.Phaserot(n_channels=2, blksiz=L, subsample=S, oversample=os_) as h:
            dt = timed(lambda: (h.reset(), h.sweep_device(x.data_ptr(), frames), h.peaks()), 3, 2)
            pk = h.peaks()
            st = h.stats()
        out[tag] = {"ms_per_pass": 1e3 * dt, "gsample_angles_per_s": float(frames) * 2 * 180 * S / dt / 1e9,
                    "hbm_gbs_algorithmic": 4.0 * frames * 2 / dt / 1e9, "argmin_index": pk[:, 1:].argmin(1).tolist(),
                    "survivor_fraction": st["points_evaluated"] / max(1, st["points_total"])}
    return out


def config4():
    sr, secs, L, ntr = 48000, 180, 8192, 1024
    frames = sr * secs
    frames -= frames % L
    tracks = [gen(frames, 2, sr, 1000 + i) for i in range(8)]  # 8 distinct tracks reused cyclically (1024 x 69 MB would only cost generator time)
    y = torch.empty(((frames // L + 1) * L, 2), device=dev, dtype=torch.float32)
    out = {"config": 4, "workload": "batch render of 1024 stereo 48 kHz 3-min tracks at per-track theta_i = (37 i mod 360) * 0.5 deg, device resident, one GPU (sharded by track: 128 per GPU on 8)"}
    with capi.Phaserot(n_channels=2, blksiz=L) as h:
        h.set_stream(torch.cuda.current_stream().cuda_stream)

        def run():
            for i in range(ntr):
                a = (i * 37) % 360
                h.render_device(tracks[i % 8].data_ptr(), frames, [a, a], 1, y.data_ptr())
        dt = timed(run, 1, 1)
    out.update({"wall_s": dt, "tracks_per_s": ntr / dt, "rotated_msamples_per_s": ntr * frames * 2 / dt / 1e6,
                "hbm_gbs_algorithmic": 8.0 * ntr * frames * 2 / dt / 1e9})
    return out


def config5():
    sr, secs, L, S, C = 192000, 3600, 32768, 100, 8
    frames = sr * secs
    frames -= frames % L
    x = gen(frames, C, sr, 45)
    out = {"config": 5, "workload": "dense sweep 0.01 deg (18000 angles) over 8-channel 192 kHz 1 h synthetic audio (22 GB resident), blksiz 32768, digital peak, one GPU"}
    with capi.Phaserot(n_channels=C, blksiz=L, subsample=S) as h:
        dt = timed(lambda: (h.reset(), h.sweep_device(x.data_ptr(), frames), h.peaks()), 2, 1)
        pk = h.peaks()
        st = h.stats()
    out.update({"ms_per_pass": 1e3 * dt, "gsample_angles_per_s": float(frames) * C * 180 * S / dt / 1e9, "hbm_gbs_algorithmic": 4.0 * frames * C / dt / 1e9,
                "argmin_index": pk[:, 1:].argmin(1).tolist(), "survivor_fraction": st["points_evaluated"] / max(1, st["points_total"])})
    return out


if __name__ == "__main__":
    which = [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4, 5]
    for k in which:
        try:
            print(json.dumps({1: config1, 2: config2, 3: config3, 4: config4, 5: config5}[k]()), flush=True)
        except Exception as ex:  # keep going: one config must not hide the others
            print(json.dumps({"config": k, "error": repr(ex)}), flush=True)
        torch.cuda.empty_cache()
