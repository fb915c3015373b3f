#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 phase-rotation backend.

Metric (BASELINE.json): Gsample-angles/s of the min-peak theta sweep.
Workload at N=1 (north_star target): stereo 48 kHz, 1 hour of synthetic
programme material, 0.1 degree grid (subsample 10 -> 1800 angles on [0, 180)),
digital peak, CLI block size 8192 — one "step" is one whole-file analysis pass
(the reference's analyze_file(), cli/phase-rotate.cc:565-587).
sample-angles per step = frames x channels x angles.

  value : input already resident in HBM (interleaved float32), CUDA events on
          the launching stream, K steps, max over ranks
  e2e   : the same pass through the C ABI with a pinned HOST buffer
          (phaserot_sweep): H2D of the whole file and D2H of the peak table
          inside the timed region
  roofline : the dominant kernel (FFT convolution + filter) timed live with
          CUDA events around every launch; achieved = 4 B/sample algorithmic
          bytes / kernel time, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline : the reference's own analysis code (oracle/_ref, unmodified
          sources + stand-in FFT) on the host cores, bounded sample

N > 1 (torchrun): weak scaling by sample range — every rank sweeps its own
1-hour shard of an N-hour stream (halo = one block of history), then one NCCL
max all-reduce over the [channels x angles] peak table.

`--impl reference` runs only the reference CPU arm with the same JSON shape.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 48000
CHANNELS = 2
BLKSIZ = 8192
SUBSAMPLE = 10
GEN_CHUNK = 1 << 21  # frames per generator chunk (absolute-index aligned)
N_PARTIALS = 16


# ---------------------------------------------------------------------------
# synthetic programme material: 16 random-phase partials (1/f), slow AM, noise
# ---------------------------------------------------------------------------

def _partials(seed):
    rng = np.random.default_rng(seed)
    f = np.exp(rng.uniform(np.log(50.0), np.log(15000.0), (CHANNELS, N_PARTIALS)))
    ph = rng.uniform(0, 1.0, (CHANNELS, N_PARTIALS))
    amp = 50.0 / f
    amp /= amp.sum(axis=1, keepdims=True)
    return f, ph, amp


def gen_chunk_torch(torch, chunk_id, device, seed=43):
    """Frames [chunk_id*GEN_CHUNK, (chunk_id+1)*GEN_CHUNK) -> [GEN_CHUNK, CHANNELS] float32 on `device`."""
    f, ph, amp = _partials(seed)
    t = torch.arange(chunk_id * GEN_CHUNK, (chunk_id + 1) * GEN_CHUNK, device=device, dtype=torch.float64) / SR
    out = torch.empty((GEN_CHUNK, CHANNELS), device=device, dtype=torch.float32)
    g = torch.Generator(device=device)
    g.manual_seed(seed * 1000003 + chunk_id)
    for c in range(CHANNELS):
        acc = torch.zeros(GEN_CHUNK, device=device, dtype=torch.float32)
        for k in range(N_PARTIALS):
            frac = torch.frac(t * float(f[c, k]) + float(ph[c, k])).to(torch.float32)
            acc += float(amp[c, k]) * torch.sin(frac * (2.0 * np.pi))
        env = 0.6 + 0.4 * torch.sin((torch.frac(t * 0.37) * (2.0 * np.pi)).to(torch.float32) + float(c))
        noise = torch.randn(GEN_CHUNK, device=device, dtype=torch.float32, generator=g)
        out[:, c] = 0.8 * acc * env + 0.02 * noise
    return out


def gen_numpy(n_frames, seed=43):
    """CPU twin of the generator for the reference arm (same recipe; noise stream differs)."""
    f, ph, amp = _partials(seed)
    t = np.arange(n_frames, dtype=np.float64) / SR
    rng = np.random.default_rng(seed)
    out = np.empty((n_frames, CHANNELS), np.float32)
    for c in range(CHANNELS):
        acc = np.zeros(n_frames)
        for k in range(N_PARTIALS):
            acc += amp[c, k] * np.sin(2 * np.pi * ((t * f[c, k] + ph[c, k]) % 1.0))
        env = 0.6 + 0.4 * np.sin(2 * np.pi * ((t * 0.37) % 1.0) + c)
        out[:, c] = 0.8 * acc * env + 0.02 * rng.standard_normal(n_frames)
    return out


# ---------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ---------------------------------------------------------------------------

class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.  NVML in-process (a step is ~1.4 ms, the
    timed region some 15 ms: nvidia-smi's loop mode does not even start that fast), sampled every
    0.5 ms on a thread; nvidia-smi -lms as the fallback when pynvml is not importable."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.nvml = None
        self.samples = []
        self.stop_flag = False
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                import torch
                uuid = "GPU-" + str(torch.cuda.get_device_properties(gpu_index).uuid)
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                try:
                    rs = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    rs = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.samples.append((sm, rs))
            except Exception:
                break
            time.sleep(0.0005)

    def start(self):
        if self.nvml is not None:
            self.stop_flag = False
            self.thr = threading.Thread(target=self._poll, daemon=True)
            self.thr.start()
            return
        self._start_smi()

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.thr.join(timeout=2)
            n = self.nvml
            try:
                mx = float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM))
            except Exception:
                mx = None
            sm = [float(a) for a, _ in self.samples]
            reasons = set()
            for _, rs in self.samples:
                for bit, nm in self.BITS.items():
                    if rs & bit:
                        reasons.add(nm)
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons),
                    "source": "nvml, 0.5 ms period, timed region only"}
        return self._stop_smi()

    def _start_smi(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._pump, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def _stop_smi(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [s.strip() for s in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for nm, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------
# reference CPU arm
# ---------------------------------------------------------------------------

def run_reference(sample_seconds, steps, warmup, threads_total=None):
    """Times the reference's analyze_file() (unmodified source, oracle/_ref) on the host cores.

    The reference uses one thread per channel (cli/phase-rotate.cc:437-443); to
    occupy the box, nproc // channels independent instances run side by side,
    each over the whole sample.  Grid: the reference's full 0.5 degree grid
    (`-s 1`: 360 indices), the finest it supports.
    """
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    if not O.have_ref():
        raise RuntimeError("oracle/_ref is missing (built from /root/reference in the authoring container)")
    n_frames = int(sample_seconds * SR)
    x = gen_numpy(n_frames)
    nproc = os.cpu_count() or 1
    inst = max(1, (threads_total or nproc) // CHANNELS)
    lib = O.ref_cli(f32=True)
    peaks = [np.zeros((CHANNELS, 360), np.float32) for _ in range(inst)]

    def one(i):
        lib.ref_cli_analyze(x, n_frames, CHANNELS, BLKSIZ, 0, 360, 1, -1, peaks[i])

    def step():
        th = [threading.Thread(target=one, args=(i,)) for i in range(inst)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        return time.perf_counter() - t0

    for _ in range(warmup):
        step()
    times = [step() for _ in range(steps)]
    sa_per_step = float(inst) * n_frames * CHANNELS * 360
    total = sum(times)
    value = sa_per_step * steps / total / 1e9
    best = sa_per_step / min(times) / 1e9
    return {
        "value": value, "best": best, "unit": "Gsample-angles/s", "cores": inst * CHANNELS, "kind": "reference",
        "ms_per_step": 1e3 * total / steps,
        "sample": f"{sample_seconds:g} s of the same stereo 48 kHz programme, reference grid 0.5 deg (360 indices, -s 1), "
                  f"{inst} concurrent instances x {CHANNELS} threads; reference sources unmodified, FFTW replaced by the stand-in float FFT",
    }


def reference_main(args):
    r = run_reference(args.ref_seconds, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "min-peak theta sweep throughput", "value": r["value"], "unit": "Gsample-angles/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, ref=True),
        "cpu_baseline": {"value": r["value"], "unit": r["unit"], "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "Gsample-angles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(args, ref=False):
    cfg = {
        "workload": f"CLI min-peak sweep: stereo 48 kHz, {args.seconds:g} s per GPU, synthetic programme (16 partials x AM + noise), "
                    f"{1.0 / args.subsample:g} deg grid ({180 * args.subsample} angles), digital peak, blksiz {BLKSIZ}",
        "frames_per_gpu": int(args.seconds * SR), "channels": CHANNELS, "angles": 180 * args.subsample,
        "l2": "input (1.38 GB per GPU at 1 h) is larger than the 126 MB L2; no explicit flush",
        "sharding": "sample-range, one shard per rank, NCCL max all-reduce of the peak table",
    }
    if ref:
        cfg["reference_arm"] = (f"bounded sample: {args.ref_seconds:g} s of the same material on the reference's own grid "
                                "(0.5 deg, 360 indices: the reference cannot run finer grids); rate in the same unit")
    return cfg


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------

def secondary_legs(torch, capi, dev, local, x, frames, hbm_peak):
    """BASELINE.json's metric also names 'rotated Msamples/s': the CLI render (cli/phase-rotate.cc:950-1003,
    config 4 style) and the plugin run() (src/phaserotate.c:774-852, config 2) measured on the same box,
    plus the true-peak variant of the sweep (config 3).  Bounded to a few seconds."""
    out = {}
    ev = lambda: torch.cuda.Event(enable_timing=True)
    # -- CLI render, device resident: 10 min stereo per call, fixed per-channel angles
    rf = min(frames, 600 * SR)
    rf -= rf % BLKSIZ
    y = torch.empty(((rf // BLKSIZ + 1) * BLKSIZ, CHANNELS), device=dev, dtype=torch.float32)
    with capi.Phaserot(mode=capi.MODE_CLI, n_channels=CHANNELS, blksiz=BLKSIZ, device=local) as hr:
        hr.set_stream(torch.cuda.current_stream().cuda_stream)
        for _ in range(2):
            hr.render_device(x.data_ptr(), rf, [37, 181], 1, y.data_ptr())
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        e0.record()
        n_r = 5
        for _ in range(n_r):
            hr.render_device(x.data_ptr(), rf, [37, 181], 1, y.data_ptr())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n_r
        hr.set_profiling(True)
        hr.render_device(x.data_ptr(), rf, [37, 181], 1, y.data_ptr())
        hr.sync()
        kt = hr.kernel_times()
        gbs = 8.0 * rf * CHANNELS / (ms * 1e-3) / 1e9
        out["render"] = {"value": rf * CHANNELS / (ms * 1e-3) / 1e6, "unit": "rotated Msamples/s", "ms_per_call": ms,
                         "workload": f"CLI render, stereo 48 kHz, {rf / SR:g} s per call, device resident, angles 18.5/90.5 deg",
                         "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                                      "algorithmic_bytes_per_sample": 8},
                         "kernels_ms": {k: round(v["ms"], 4) for k, v in kt.items() if v["launches"]}}
    del y
    # -- plugin run(): mono 48 kHz, angle port at 90 deg from the first call (ramp included), 1024-frame calls
    rng = np.random.default_rng(42)
    n_calls, blk = 2000, 1024
    xin = (0.25 * rng.standard_normal(n_calls * blk)).astype(np.float32)
    yout = np.zeros_like(xin)
    ang = np.array([90.0], np.float32)
    with capi.Phaserot(mode=capi.MODE_PLUGIN, n_channels=1, sample_rate=48000.0, device=local) as hp:
        ins = (ctypes.c_void_p * 1)()
        outs = (ctypes.c_void_p * 1)()
        def run(lo, hi):
            for k in range(lo, hi):
                ins[0] = xin.ctypes.data + 4 * k * blk
                outs[0] = yout.ctypes.data + 4 * k * blk
                hp.process_raw(ins, outs, blk, ang)
        run(0, 200)
        t0 = time.perf_counter()
        run(200, n_calls)
        dt = time.perf_counter() - t0
        out["plugin"] = {"value": (n_calls - 200) * blk / dt / 1e6, "unit": "rotated Msamples/s", "us_per_call": 1e6 * dt / (n_calls - 200),
                         "workload": "LV2 run(): mono 48 kHz, 1024-frame calls, angle 90 deg, host buffers in and out (synchronous round trip per call)"}
        # bulk: 10 min in one call (host buffers; H2D + D2H inside)
        nb = 600 * SR
        xb = (0.25 * rng.standard_normal(nb)).astype(np.float32)
        yb = np.zeros_like(xb)
        hp.reset()
        ins[0], outs[0] = xb.ctypes.data, yb.ctypes.data
        hp.process_raw(ins, outs, nb, ang)
        hp.reset()
        t0 = time.perf_counter()
        hp.process_raw(ins, outs, nb, ang)
        dt = time.perf_counter() - t0
        out["plugin_bulk"] = {"value": nb / dt / 1e6, "unit": "rotated Msamples/s", "ms_per_call": 1e3 * dt,
                              "workload": "LV2 run(): mono 48 kHz, one 600 s call, pageable host buffers in and out"}
    # -- end to end from a 16-bit PCM file image: the headline workload quantised to int16 in
    #    pinned host memory, through phaserot_sweep_pcm (H2D of 2 bytes per sample + widening on the device)
    q = torch.empty((frames, CHANNELS), dtype=torch.int16, pin_memory=True)
    q.copy_((x * 32768.0).round().clamp_(-32768, 32767).to(torch.int16))
    torch.cuda.synchronize()
    with capi.Phaserot(mode=capi.MODE_CLI, n_channels=CHANNELS, blksiz=BLKSIZ, subsample=SUBSAMPLE, device=local) as hq:
        for _ in range(2):
            hq.reset()
            hq.sweep_pcm((q.data_ptr(), frames, np.int16))
        n_q = 3
        t0 = time.perf_counter()
        for _ in range(n_q):
            hq.reset()
            hq.sweep_pcm((q.data_ptr(), frames, np.int16))
            hq.peaks()
        dt = (time.perf_counter() - t0) / n_q
        out["e2e_pcm16"] = {"value": float(frames) * CHANNELS * 180 * SUBSAMPLE / dt / 1e9, "unit": "Gsample-angles/s", "ms_per_step": 1e3 * dt,
                            "h2d_bytes_per_step": int(q.numel() * 2),
                            "workload": "the headline sweep from a 16-bit PCM image in pinned host memory (phaserot_sweep_pcm): integers over PCIe, widened on the device"}
    del q
    # -- true-peak sweep (4x), same 1800-angle grid, 10 min stereo device resident
    tf = min(frames, 600 * SR)
    tf -= tf % (32768 - BLKSIZ)
    with capi.Phaserot(mode=capi.MODE_CLI, n_channels=CHANNELS, blksiz=BLKSIZ, subsample=SUBSAMPLE, device=local, oversample=4) as ht:
        ht.set_stream(torch.cuda.current_stream().cuda_stream)
        for _ in range(2):
            ht.reset()
            ht.sweep_device(x.data_ptr(), tf)
            ht.peaks()
        e0, e1 = ev(), ev()
        e0.record()
        n_t = 3
        for _ in range(n_t):
            ht.reset()
            ht.sweep_device(x.data_ptr(), tf)
            ht.peaks()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n_t
        ht.set_profiling(True)
        ht.reset()
        ht.sweep_device(x.data_ptr(), tf)
        ht.peaks()
        kt = ht.kernel_times()
        out["true_peak"] = {"value": float(tf) * CHANNELS * 180 * SUBSAMPLE / (ms * 1e-3) / 1e9, "unit": "Gsample-angles/s", "ms_per_step": ms,
                            "workload": f"4x oversampled true-peak sweep (new capability), stereo 48 kHz, {tf / SR:g} s, 1800 angles, device resident",
                            "hbm_frac_algorithmic": 4.0 * tf * CHANNELS / (ms * 1e-3) / 1e9 / hbm_peak,
                            "kernels_ms": {k: round(v["ms"], 4) for k, v in kt.items() if v["launches"]}}
    return out


def bind_to_gpu_numa_node(torch, local):
    """Run this rank on the CPUs NVML reports as local to its GPU (what `numactl --cpunodebind` does for a
    production launch), BEFORE any page-locked buffer is allocated: with 8 ranks uploading at once, host
    buffers that sit on the other socket share the inter-socket link and the end-to-end rate of those
    ranks drops to a third.  Returns the CPU list, or None when nothing was changed."""
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = "GPU-" + str(torch.cuda.get_device_properties(local).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
        ncpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {i for i in range(ncpu) if (mask[i // 64] >> (i % 64)) & 1} & set(os.sched_getaffinity(0))
        if cpus and cpus != set(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, cpus)
            return sorted(cpus)
    except Exception:
        pass
    return None


def gpu_main(args):
    import torch
    import torch.distributed as dist
    from phaserotate.lv2_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(torch, local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    frames = int(args.seconds * SR)
    frames -= frames % (32768 - BLKSIZ)  # whole blocks, cut on the FFT segment grid (phaserot_shard_align)
    A = 180 * args.subsample
    n_chunks = (frames + GEN_CHUNK - 1) // GEN_CHUNK

    # ---- synthesise this rank's shard on the device (absolute stream position = rank * frames)
    chunk0 = rank * ((frames + GEN_CHUNK - 1) // GEN_CHUNK)
    x = torch.empty((n_chunks * GEN_CHUNK, CHANNELS), device=dev, dtype=torch.float32)
    for k in range(n_chunks):
        x[k * GEN_CHUNK:(k + 1) * GEN_CHUNK] = gen_chunk_torch(torch, chunk0 + k, dev)
    x = x[:frames].contiguous()
    hist = None
    if rank > 0:
        prev = gen_chunk_torch(torch, chunk0 - 1, dev)
        # the previous rank's shard ends at frame `frames` of its own chunk range
        prev_tail_end = frames - (n_chunks - 1) * GEN_CHUNK
        hist = prev[prev_tail_end - BLKSIZ:prev_tail_end].contiguous()  # stays on the device, read in place
    hist_ptr = hist.data_ptr() if hist is not None else None
    torch.cuda.synchronize()

    h = capi.Phaserot(mode=capi.MODE_CLI, n_channels=CHANNELS, blksiz=BLKSIZ, subsample=args.subsample, device=local,
                      flags=capi.FLAG_NO_PRUNE if args.no_prune else 0)
    stream = torch.cuda.current_stream()
    h.set_stream(stream.cuda_stream)
    tables = {}

    def combine_shards(handle):
        """NCCL max all-reduce, in place, on the handle's device-resident table of the pending
        sweep (per-angle maxima + raw peaks; phaserot_pending_table): no host round trip."""
        ptr, nc, na = handle.pending_table()
        n = nc * na + nc
        t = tables.get((ptr, n))
        if t is None:
            class _Dev:
                __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}
            t = tables[(ptr, n)] = torch.as_tensor(_Dev(), device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)

    def step_device():
        h.reset()
        h.sweep_shard_device(x.data_ptr(), frames, hist_ptr, rank == 0, rank == world - 1)
        if world > 1:
            combine_shards(h)
        return h.peaks()  # sync + D2H of the (combined) table

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        wall = time.perf_counter() - w0
        ms = e0.elapsed_time(e1)
        # the table read-back is a host sync inside the step; wall and event time agree, keep the larger
        t = torch.tensor([max(ms / 1e3, wall if args.wall else 0.0)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(max(args.warmup, 3)):
        step_device()
    h.reset_stats()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t_dev = timed(step_device, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    st = h.stats()
    launches = torch.tensor([st["kernel_launches"]], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(launches)
    sa_step = float(frames) * CHANNELS * A * world
    value = sa_step * args.steps / t_dev / 1e9
    surv = st["points_evaluated"] / max(1, st["points_total"])

    # ---- e2e: pinned host buffer through phaserot_sweep (H2D + table D2H inside the timed region)
    xh = torch.empty((frames, CHANNELS), dtype=torch.float32, pin_memory=True)
    xh.copy_(x)
    torch.cuda.synchronize()
    he = capi.Phaserot(mode=capi.MODE_CLI, n_channels=CHANNELS, blksiz=BLKSIZ, subsample=args.subsample, device=local,
                       flags=capi.FLAG_NO_PRUNE if args.no_prune else 0)
    if world > 1:
        he.set_stream(stream.cuda_stream)  # the upload, the sweep and the all-reduce are ordered on one stream

    def step_e2e():
        he.reset()
        if world == 1:
            he.sweep((xh.data_ptr(), frames))
        else:
            # shard semantics need history: upload then shard call (H2D still inside the step)
            x.copy_(xh, non_blocking=True)
            he.sweep_shard_device(x.data_ptr(), frames, hist_ptr, rank == 0, rank == world - 1)
            combine_shards(he)
        he.peaks()

    for _ in range(2):
        step_e2e()
    e2e_steps = max(2, min(args.steps, 5))
    barrier()
    w0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    t_e2e = torch.tensor([time.perf_counter() - w0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = sa_step * e2e_steps / float(t_e2e.item()) / 1e9
    he.close()
    del xh

    # ---- roofline leg: per-kernel CUDA-event times of one more step (not part of `value`)
    h.set_profiling(True)
    step_device()
    kt = h.kernel_times()
    h.set_profiling(False)
    conv = kt["fftconv_filter"]
    alg_bytes = 4.0 * frames * CHANNELS  # per rank, per pass (4 B per input sample)
    peak, peak_src = measured_hbm_peak()
    achieved = alg_bytes / (conv["ms"] * 1e-3) / 1e9 if conv["ms"] > 0 else 0.0
    step_ms = 1e3 * t_dev / args.steps
    kshare = {k: round(v["ms"], 4) for k, v in kt.items() if v["launches"]}

    # ---- secondary legs (rank 0, N=1): the other two callers of the path.  Not part of `value`.
    extra = {}
    if world == 1 and not args.no_extra:
        extra = secondary_legs(torch, capi, dev, local, x, frames, peak)

    line = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            try:
                r = run_reference(args.ref_seconds, 2, 1)
                cpu = {"value": r["value"], "unit": r["unit"], "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
            except Exception as ex:  # the oracle build is test infrastructure; never fatal for the GPU line
                cpu = {"value": None, "unit": "Gsample-angles/s", "cores": 0, "kind": "reference", "sample": f"unavailable: {ex}"}
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("fftconv_filter_dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": "min-peak theta sweep throughput", "value": value, "unit": "Gsample-angles/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Gsample-angles/s", "h2d_bytes_per_step": int(frames * CHANNELS * 4 * world),
                    "host_numa_binding": ("rank bound to %d GPU-local CPUs" % len(numa)) if numa else "none",
                    "d2h_bytes_per_step": int((CHANNELS * A + CHANNELS) * 4 * world), "steps": e2e_steps,
                    "ms_per_step": 1e3 * float(t_e2e.item()) / e2e_steps},
            "gpu_launches": int(launches.item()),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": "fftconv_kernel<EPI_POINTS>", "peak_source": peak_src,
                         "algorithmic_bytes_per_step_per_gpu": alg_bytes, "kernel_ms_per_step": conv["ms"], "launches_per_step": conv["launches"],
                         "kernel_share_of_step": conv["ms"] / step_ms if step_ms else None,
                         "whole_step_frac": alg_bytes / (step_ms * 1e-3) / 1e9 / peak,
                         "all_kernels_ms": kshare},
            "pruning": {"enabled": not args.no_prune, "survivor_fraction": surv, "points_per_step_per_gpu": st["points_total"] // max(1, args.steps)},
            "cpu_baseline": cpu,
        }
        line.update(extra)
    h.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--seconds", type=float, default=3600.0, help="audio seconds per GPU (default: the 1-hour headline workload)")
    ap.add_argument("--subsample", type=int, default=SUBSAMPLE)
    ap.add_argument("--ref-seconds", type=float, default=240.0, help="bounded CPU sample for the reference arm")
    ap.add_argument("--no-prune", action="store_true", help="evaluate every sample at every angle")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the render / plugin / true-peak secondary legs")
    ap.add_argument("--wall", action="store_true", help="use max(wall, events) as the step time")
    args = ap.parse_args()
    if args.impl == "reference":
        # under torchrun only rank 0 works; the other ranks exit quietly
        if int(os.environ.get("RANK", "0")) != 0:
            return
        reference_main(args)
        return
    gpu_main(args)


if __name__ == "__main__":
    main()
