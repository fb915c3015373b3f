"""ctypes access to the oracle (C restatement) and to the reference build.

TEST INFRASTRUCTURE.  `oracle/_build/liboracle.so` is the CPU restatement
(oracle/phaserot_oracle.c); `oracle/_ref/*` are the reference's own sources
compiled unmodified (oracle/Makefile, target `ref`).  Neither is ever used by
the product path.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
BUILD_DIR = os.path.join(ORACLE_DIR, "_build")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")

f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build_oracle():
    """(Re)build the C restatement and the LV2 harness; cheap, CPU only."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "oracle"], check=True)


def build_ref():
    """Build oracle/_ref from /root/reference when that tree is present."""
    if os.path.isdir("/root/reference/cli"):
        subprocess.run(["make", "-s", "-C", ORACLE_DIR, "ref"], check=True)


_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        path = os.path.join(BUILD_DIR, "liboracle.so")
        if not os.path.exists(path):
            build_oracle()
        lib = C.CDLL(path)
        lib.pro_fir_taps.argtypes = [C.c_int, f32p]
        lib.pro_fir_taps_plugin.argtypes = [C.c_int, f32p]
        lib.pro_sincos_lut.argtypes = [C.c_int, f32p, f32p]
        lib.pro_hilbert_fir.argtypes = [f32p, C.c_int64, f32p, C.c_int, f32p, C.c_int64]
        lib.pro_cli_analyze.argtypes = [f32p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, f32p]
        lib.pro_cli_analyze_shard.argtypes = [f32p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, f32p]
        lib.pro_cli_analyze_tp_shard.argtypes = [f32p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, f32p]
        lib.pro_cli_apply.argtypes = [f32p, C.c_int64, C.c_int, C.c_int, C.c_int, i32p, C.c_int, f32p]
        lib.pro_cli_render_file.argtypes = [f32p, C.c_int64, C.c_int, C.c_int, C.c_int, i32p, f32p]
        lib.pro_cli_render_file.restype = C.c_int64
        lib.pro_plugin_run.argtypes = [C.c_double, f32p, f32p, C.c_int64, C.c_uint32, f32p]
        lib.pro_plugin_sizes.argtypes = [C.c_double] + [C.POINTER(C.c_uint32)] * 4
        _oracle = lib
    return _oracle


def have_ref():
    return os.path.exists(os.path.join(REF_DIR, "libref_cli.so"))


_ref = {}


def ref_cli(f32=False):
    """The reference CLI classes behind oracle/ref_cli_harness.cc."""
    key = "f32" if f32 else "f64"
    if key not in _ref:
        lib = C.CDLL(os.path.join(REF_DIR, "libref_cli_f32.so" if f32 else "libref_cli.so"))
        lib.ref_cli_analyze.argtypes = [f32p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, f32p]
        lib.ref_cli_analyze.restype = C.c_double
        lib.ref_cli_apply.argtypes = [f32p, C.c_int64, C.c_int, C.c_int, i32p, C.c_int, f32p]
        lib.ref_cli_apply.restype = C.c_double
        lib.ref_cli_lut.argtypes = [f32p, f32p]
        lib.ref_cli_taps.argtypes = [C.c_int, f32p]
        lib.ref_cli_maxsample.restype = C.c_int
        _ref[key] = lib
    return _ref[key]


_lv2h = None


def lv2_harness():
    global _lv2h
    if _lv2h is None:
        path = os.path.join(BUILD_DIR, "liblv2harness.so")
        if not os.path.exists(path):
            build_oracle()
        lib = C.CDLL(path)
        lib.lv2h_render.argtypes = [C.c_char_p, C.c_int, C.c_double, C.c_int, f32p, f32p, C.c_int64, C.c_uint32, f32p, C.c_int, C.POINTER(C.c_float)]
        lib.lv2h_render.restype = C.c_double
        _lv2h = lib
    return _lv2h


# ---------------------------------------------------------------------------
# convenience wrappers
# ---------------------------------------------------------------------------

def _inter(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    if x.ndim == 1:
        x = x[:, None]
    return x


def oracle_analyze(x, blksiz, subsample=2, ang_start=0, ang_end=None, stride=1, only_chn=-1, peaks=None):
    x = _inter(x)
    n, c = x.shape
    ms = 180 * subsample
    if ang_end is None:
        ang_end = ms
    if peaks is None:
        peaks = np.zeros((c, ms), np.float32)
    oracle().pro_cli_analyze(x, n, c, blksiz, subsample, ang_start, ang_end, stride, only_chn, peaks)
    return peaks


def oracle_analyze_shard(x, blksiz, hist, first, last, subsample=2, ang_start=0, ang_end=None, stride=1, only_chn=-1, peaks=None):
    x = _inter(x)
    n, c = x.shape
    ms = 180 * subsample
    if ang_end is None:
        ang_end = ms
    if peaks is None:
        peaks = np.zeros((c, ms), np.float32)
    hp = None
    if hist is not None:
        hist = np.ascontiguousarray(hist, np.float32)
        assert hist.shape == (blksiz, c)
        hp = hist.ctypes.data
    oracle().pro_cli_analyze_shard(x, n, c, blksiz, subsample, hp, int(first), int(last), ang_start, ang_end, stride, only_chn, peaks)
    return peaks


def oracle_analyze_tp(x, blksiz, oversample=4, hist=None, first=True, last=True, subsample=2, ang_start=0, ang_end=None, stride=1,
                      only_chn=-1, peaks=None):
    """Oversampled true-peak analysis (this library's own definition, include/phaserot_cuda.h; not a reference feature)."""
    x = _inter(x)
    n, c = x.shape
    ms = 180 * subsample
    if ang_end is None:
        ang_end = ms
    if peaks is None:
        peaks = np.zeros((c, ms), np.float32)
    hp = None
    if hist is not None:
        hist = np.ascontiguousarray(hist, np.float32)
        assert hist.shape == (blksiz, c)
        hp = hist.ctypes.data
    oracle().pro_cli_analyze_tp_shard(x, n, c, blksiz, subsample, oversample, hp, int(first), int(last), ang_start, ang_end, stride, only_chn, peaks)
    return peaks


def ref_analyze(x, blksiz, ang_start=0, ang_end=360, stride=1, only_chn=-1, f32=False):
    x = _inter(x)
    n, c = x.shape
    peaks = np.zeros((c, 360), np.float32)
    dt = ref_cli(f32).ref_cli_analyze(x, n, c, blksiz, ang_start, ang_end, stride, only_chn, peaks)
    assert dt >= 0
    return peaks, dt


def oracle_apply(x, blksiz, angles, flush_blocks=1, subsample=2):
    x = _inter(x)
    n, c = x.shape
    nblk = (n + blksiz - 1) // blksiz + flush_blocks
    out = np.zeros((nblk * blksiz, c), np.float32)
    oracle().pro_cli_apply(x, n, c, blksiz, subsample, np.asarray(angles, np.int32), flush_blocks, out)
    return out


def ref_apply(x, blksiz, angles, flush_blocks=1, f32=False):
    x = _inter(x)
    n, c = x.shape
    nblk = (n + blksiz - 1) // blksiz + flush_blocks
    out = np.zeros((nblk * blksiz, c), np.float32)
    dt = ref_cli(f32).ref_cli_apply(x, n, c, blksiz, np.asarray(angles, np.int32), flush_blocks, out)
    return out, dt


def oracle_render_file(x, blksiz, angles, subsample=2):
    x = _inter(x)
    n, c = x.shape
    out = np.zeros((n + 2 * blksiz, c), np.float32)
    w = oracle().pro_cli_render_file(x, n, c, blksiz, subsample, np.asarray(angles, np.int32), out)
    return out[:w].copy()


def oracle_plugin_run(x, rate, block, angles):
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros_like(x)
    oracle().pro_plugin_run(float(rate), x, out, x.shape[0], block, np.ascontiguousarray(angles, np.float32))
    return out


def lv2_render(so_path, x_planar, rate, block, angles, inplace=False):
    """x_planar: [n_chn][n]; angles: [n_calls][n_chn] degrees. Returns (out, latency, seconds)."""
    x = np.ascontiguousarray(x_planar, np.float32)
    if x.ndim == 1:
        x = x[None, :]
    c, n = x.shape
    ncalls = (n + block - 1) // block
    ang = np.ascontiguousarray(np.broadcast_to(np.asarray(angles, np.float32).reshape(-1, c) if np.ndim(angles) else np.full((ncalls, c), angles, np.float32), (ncalls, c)))
    out = np.zeros_like(x)
    lat = C.c_float(-1)
    dt = lv2_harness().lv2h_render(so_path.encode(), 0 if c == 1 else 1, float(rate), c, x, out, n, block, ang, int(inplace), C.byref(lat))
    if dt < 0:
        raise RuntimeError(f"lv2 harness failed ({dt}) for {so_path}")
    return out, lat.value, dt


# ---------------------------------------------------------------------------
# WAV helpers + synthetic signals shared by tests and bench
# ---------------------------------------------------------------------------

def write_wav_f32(path, x, sr):
    import struct
    x = _inter(x).astype("<f4")
    n, c = x.shape
    data = x.tobytes()
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 3, c, sr, sr * c * 4, c * 4, 32) + b"data" + struct.pack("<I", len(data)))
        f.write(data)


def write_wav_pcm(path, q, sr, bits):
    """q: integer samples [frames, channels] (int16 for 16 bit; int32 holding 24-bit or 32-bit values)."""
    import struct
    q = np.asarray(q)
    if q.ndim == 1:
        q = q[:, None]
    n, c = q.shape
    if bits == 16:
        data = q.astype("<i2").tobytes()
    elif bits == 24:
        b = q.astype("<i4").reshape(-1).view(np.uint8).reshape(-1, 4)[:, :3]
        data = np.ascontiguousarray(b).tobytes()
    else:
        data = q.astype("<i4").tobytes()
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, c, sr, sr * c * bits // 8, c * bits // 8, bits) + b"data" + struct.pack("<I", len(data)))
        f.write(data)


def read_wav_f32(path):
    import struct
    b = open(path, "rb").read()
    assert b[:4] == b"RIFF" and b[8:12] == b"WAVE"
    pos = 12
    fmt = None
    while pos + 8 <= len(b):
        cid, ln = b[pos:pos + 4], struct.unpack("<I", b[pos + 4:pos + 8])[0]
        if cid == b"fmt ":
            fmt = struct.unpack("<HHIIHH", b[pos + 8:pos + 24])
        elif cid == b"data":
            assert fmt and fmt[0] == 3 and fmt[5] == 32
            ln = min(ln, len(b) - pos - 8)
            return np.frombuffer(b, "<f4", ln // 4, pos + 8).reshape(-1, fmt[1]).copy(), fmt[2]
        pos += 8 + ln + (ln & 1)
    raise ValueError("no data chunk")


def two_sine(sr, seconds, channels=2):
    """0.5 sin(110 Hz + phi_c) + 0.25 sin(1760.3 Hz)."""
    n = int(sr * seconds)
    t = np.arange(n, dtype=np.float64) / sr
    phis = [0.0, 1.0, 2.0, 0.5, 1.5, 2.5, 0.25, 0.75]
    chans = [0.5 * np.sin(2 * np.pi * 110 * t + phis[c % 8]) + 0.25 * np.sin(2 * np.pi * 1760.3 * t) for c in range(channels)]
    return np.stack(chans, 1).astype(np.float32)


def harmonic(sr, seconds, channels=2, seed=5):
    """Harmonic-rich, asymmetric waveforms (band-limited saw / pulse-like sums with
    random start phases): the peak depends strongly on the rotation angle and has
    one clear minimum per channel, so angle selection is not decided by rounding."""
    rng = np.random.default_rng(seed)
    n = int(sr * seconds)
    t = np.arange(n, dtype=np.float64) / sr
    out = []
    for c in range(channels):
        f0 = [110.3, 82.4, 146.9, 65.4, 98.1, 123.5, 73.4, 55.2][c % 8]
        skew = rng.uniform(0.2, 1.2)
        y = np.zeros(n)
        for k in range(1, 13):
            y += (1.0 / k) * np.sin(2 * np.pi * k * f0 * t + skew * k + 0.1 * c)
        y *= 0.8 / np.max(np.abs(y))
        y += 0.01 * rng.standard_normal(n)
        out.append(y)
    return np.stack(out, 1).astype(np.float32)


def pink_noise(n, seed=42, peak=0.5):
    """Paul Kellet's economy pink filter over seeded uniform noise, peak normalised."""
    rng = np.random.default_rng(seed)
    w = rng.uniform(-1, 1, n)
    from scipy.signal import lfilter
    b = [0.049922035, -0.095993537, 0.050612699, -0.004408786]
    a = [1, -2.494956002, 2.017265875, -0.522189400]
    y = lfilter(b, a, w)
    y *= peak / np.max(np.abs(y))
    return y.astype(np.float32)


def programme(sr, seconds, channels=2, seed=43):
    """SURVEY 8(d) config 3 style: 32 random-phase partials (1/f) x slow AM + -20 dB pink."""
    rng = np.random.default_rng(seed)
    n = int(sr * seconds)
    t = np.arange(n, dtype=np.float64) / sr
    out = []
    for c in range(channels):
        f = np.exp(rng.uniform(np.log(50), np.log(15000), 32))
        ph = rng.uniform(0, 2 * np.pi, 32)
        y = np.zeros(n)
        for fi, pi_ in zip(f, ph):
            y += (50.0 / fi) * np.sin(2 * np.pi * fi * t + pi_)
        y *= 0.6 + 0.4 * np.sin(2 * np.pi * 0.37 * t + c)
        y /= np.max(np.abs(y))
        y = 0.7 * y + 0.1 * pink_noise(n, seed + 100 + c, 1.0)
        out.append(y)
    return np.stack(out, 1).astype(np.float32)
