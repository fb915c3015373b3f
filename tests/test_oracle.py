"""CPU tests of the oracle (oracle/phaserot_oracle.c): known answers, the golden
vectors generated from the reference build (tests/golden/make_golden.py), and —
when oracle/_ref is present — the reference build itself."""
import os

import numpy as np
import pytest

import oracle_lib as O

TOL = 2e-6  # oracle (exact convolution) vs reference build (fp32 overlap-add round trips)


def rel(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-30)))


# ---------------------------------------------------------------------------
# known answers
# ---------------------------------------------------------------------------

@pytest.mark.parametrize("L", [1024, 3072, 4096, 8192, 32768])
def test_fir_taps_closed_form(oracle_built, L):
    """taps = Hann(n) * (-(2/L) cot(pi (n - L/2) / L)) on odd n - L/2, zero elsewhere."""
    taps = np.zeros(L, np.float32)
    (O.oracle().pro_fir_taps_plugin if L == 3072 else O.oracle().pro_fir_taps)(L, taps)
    n = np.arange(L)
    m = n - L // 2
    exp = np.zeros(L)
    odd = (m % 2) != 0
    exp[odd] = -(2.0 / L) / np.tan(np.pi * m[odd] / L) * 0.5 * (1 - np.cos(2 * np.pi * n[odd] / L))
    assert np.all(taps[~odd] == 0)
    assert np.max(np.abs(taps - exp)) < 2e-7
    # antisymmetric about the centre L/2
    assert np.allclose(taps[1:], -taps[1:][::-1], atol=1e-9)


def test_fir_frequency_response(oracle_built):
    """Response is +j x delay(L/2): gain 1 from bin 2 up, 0.75 at bin 1, 0 at DC."""
    L = 4096
    taps = np.zeros(L, np.float32)
    O.oracle().pro_fir_taps(L, taps)
    H = np.fft.rfft(taps.astype(np.float64)) * np.exp(2j * np.pi * np.arange(L // 2 + 1) * (L // 2) / L)
    assert abs(H[0]) < 1e-6
    assert abs(H[1] - 0.75j) < 1e-5
    assert np.max(np.abs(H[2:L // 2 - 1] - 1j)) < 1e-4


def test_lut_reference_grid(oracle_built):
    s = np.zeros(360, np.float32)
    c = np.zeros(360, np.float32)
    O.oracle().pro_sincos_lut(2, s, c)
    assert s[0] == 0 and c[0] == 1
    # index 180 (90 deg) is sa = -1, ca = cosf(-1.5707964f)
    assert s[180] == -1.0 and c[180] == np.float32(-4.371139e-08)
    a = -np.arange(360) * 0.5 * np.pi / 180
    assert np.max(np.abs(s - np.sin(a))) < 2e-7 and np.max(np.abs(c - np.cos(a))) < 2e-7


def test_sine_rotated_by_90_degrees(oracle_built):
    """A sine through the render path at 90 deg comes out as the delayed -cos... i.e. a 90 degree
    phase shift with unchanged amplitude (README of the reference: rotation keeps the spectrum)."""
    sr, L, f = 48000, 8192, 1000.0
    n = 6 * L
    t = np.arange(n) / sr
    x = (0.5 * np.sin(2 * np.pi * f * t)).astype(np.float32)
    y = O.oracle_apply(x, L, [180], 1)[:, 0]          # index 180 = 90 deg
    D = L // 2
    tt = (np.arange(2 * L, 4 * L) - D) / sr
    # y(t) = ca x(t - D) + sa H(t), sa = sin(-90 deg) = -1, H = +j response => -(0.5 sin(w t + 90deg)) = -0.5 cos
    exp = -0.5 * np.cos(2 * np.pi * f * tt)
    assert np.max(np.abs(y[2 * L:4 * L] - exp)) < 2e-5
    y0 = O.oracle_apply(x, L, [0], 1)[:, 0]
    assert np.max(np.abs(y0[D:n] - x[:n - D])) < 1e-7  # angle 0 = pure delay of L/2


def test_impulse_reproduces_taps(oracle_built):
    L = 1024
    taps = np.zeros(L, np.float32)
    O.oracle().pro_fir_taps(L, taps)
    x = np.zeros(3 * L, np.float32)
    x[0] = 1.0
    y = O.oracle_apply(x, L, [180], 1)[:, 0]  # sa = -1, ca ~ -4e-8
    assert np.max(np.abs(y[:L] + taps)) < 1e-7


def test_plugin_latency_and_leading_zeros(oracle_built):
    import ctypes as C
    for rate, lat, P in [(44100, 1792, 256), (48000, 1792, 256), (96000, 2560, 512), (192000, 5120, 1024)]:
        v = [C.c_uint32() for _ in range(4)]
        O.oracle().pro_plugin_sizes(float(rate), *[C.byref(a) for a in v])
        assert v[3].value == lat and v[2].value == P
        x = O.pink_noise(8 * P, 3)
        y = O.oracle_plugin_run(x, rate, 2 * P, np.zeros(4, np.float32))
        assert np.all(y[:P] == 0)
        # angle 0: output is the input delayed by the reported latency
        assert np.max(np.abs(y[lat:] - x[:len(x) - lat])) < 1e-7


def test_first_block_and_raw_peak_rules(oracle_built):
    """Q1/Q2 of SURVEY 3.3 on a signal that is loud only in the first half block."""
    L = 1024
    x = np.zeros((4 * L, 1), np.float32)
    x[: L // 2, 0] = O.pink_noise(L // 2, 9, 0.9)
    pk = O.oracle_analyze(x, L)
    assert pk[0, 0] == np.max(np.abs(x))          # index 0 = raw input peak
    # the direct branch of the first half block is never examined: at ~0 deg (index 1) only sa*H contributes
    assert pk[0, 1] < 0.2 * pk[0, 0]


def test_empty_and_tiny_inputs(oracle_built):
    pk = O.oracle_analyze(np.zeros((0, 2), np.float32), 1024)
    assert pk.shape == (2, 360) and np.all(pk == 0)
    x = np.array([[0.5], [-0.25], [0.125]], np.float32)
    pk = O.oracle_analyze(x, 1024)
    assert pk[0, 0] == 0.5 and np.all(pk[0, 1:] > 0)


def test_angle_schedule_semantics(oracle_built):
    """thr_process's loop (cli:409-428): start..end step stride, leaving once angle >= end."""
    x = O.harmonic(48000, 0.2, 1)
    pk = O.oracle_analyze(x, 1024, 2, 0, 360, 24)
    assert set(np.nonzero(pk[0])[0]) == set(range(0, 360, 24))
    pk = O.oracle_analyze(x, 1024, 2, -12, 13, 1)
    assert set(np.nonzero(pk[0])[0]) == set(range(348, 360)) | set(range(0, 13))
    full = O.oracle_analyze(x, 1024)
    assert np.array_equal(pk[0, :13], full[0, :13]) and np.array_equal(pk[0, 348:], full[0, 348:])


# ---------------------------------------------------------------------------
# golden vectors from the reference build
# ---------------------------------------------------------------------------

def test_golden_cli_analyze(oracle_built, golden):
    g = golden["cli_analyze"]
    x, L = g["x"], int(g["blksiz"])
    s = np.zeros(360, np.float32)
    c = np.zeros(360, np.float32)
    O.oracle().pro_sincos_lut(2, s, c)
    assert np.array_equal(s, g["lut_sin"]) and np.array_equal(c, g["lut_cos"])
    taps = np.zeros(L, np.float32)
    O.oracle().pro_fir_taps(L, taps)
    assert np.max(np.abs(taps - g["taps"])) < 5e-8
    assert rel(O.oracle_analyze(x, L), g["full"]) < TOL
    assert rel(O.oracle_analyze(x, L, 2, 0, 360, 24), g["coarse"]) < TOL
    assert rel(O.oracle_analyze(x, L, 2, -12, 13, 1), g["refine"]) < TOL
    assert rel(O.oracle_analyze(x, L, 2, 36, 61, 1, 1), g["single"]) < TOL
    assert np.array_equal(O.oracle_analyze(x, L).argmin(1), g["full"].argmin(1))


def test_golden_cli_render(oracle_built, golden):
    g = golden["cli_render"]
    x, L, ang = g["x"], int(g["blksiz"]), g["angles"]
    assert np.max(np.abs(O.oracle_apply(x, L, ang, 1) - g["stream"])) < TOL
    # file loop incl. the float-offset (R1) and stale-tail (R2) quirks
    assert np.max(np.abs(O.oracle_render_file(x, L, ang) - g["file_stereo"])) < TOL
    assert np.max(np.abs(O.oracle_render_file(x[:, :1].copy(), L, ang[:1]) - g["file_mono"])) < TOL
    assert np.max(np.abs(O.oracle_render_file(x[:700], L, ang) - g["file_short"])) < TOL
    assert np.max(np.abs(O.oracle_render_file(g["x_exact"], L, ang) - g["file_exact"])) < TOL
    assert g["file_stereo"].shape[0] == x.shape[0]


def test_golden_plugin(oracle_built, golden):
    g = golden["plugin"]
    for rate, blk in [(48000, 256), (48000, 1000), (96000, 1024), (192000, 2048)]:
        k = f"r{rate}_b{blk}"
        y = O.oracle_plugin_run(g[k + "_x"], rate, blk, g[k + "_ang"])
        assert np.max(np.abs(y - g[k + "_y"])) < TOL, k
    for c in range(2):
        y = O.oracle_plugin_run(g["stereo_x"][c], 48000, 1000, g["stereo_ang"][:, c].copy())
        assert np.max(np.abs(y - g["stereo_y"][c])) < TOL


# ---------------------------------------------------------------------------
# against the reference build itself (present when built here or shipped to the GPU box)
# ---------------------------------------------------------------------------

needs_ref = pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built (needs /root/reference)")


@needs_ref
@pytest.mark.parametrize("sig,L", [("two_sine", 8192), ("pink", 4096), ("programme", 16384), ("harmonic", 1024)])
def test_oracle_vs_reference_peaks(oracle_built, sig, L):
    x = {"two_sine": lambda: O.two_sine(48000, 1.5, 2), "pink": lambda: O.pink_noise(60000, 1)[:, None],
         "programme": lambda: O.programme(96000, 1.0, 2), "harmonic": lambda: O.harmonic(48000, 0.7, 3)}[sig]()
    pr, _ = O.ref_analyze(x, L)
    po = O.oracle_analyze(x, L)
    assert rel(po, pr) < TOL
    pr32, _ = O.ref_analyze(x, L, f32=True)  # the timed-baseline build (float stand-in FFT) stays within the budget too
    assert rel(pr32, po) < 1e-5


@needs_ref
def test_oracle_vs_reference_render_and_plugin(oracle_built):
    x = O.programme(48000, 0.8, 2)
    yr, _ = O.ref_apply(x, 4096, [-45, 359], 1)
    assert np.max(np.abs(O.oracle_apply(x, 4096, [-45, 359], 1) - yr)) < TOL
    so = os.path.join(O.REF_DIR, "phaserotate_ref.so")
    xm = O.pink_noise(30000, 8)
    for blk in (64, 1024, 5000):
        ncalls = (len(xm) + blk - 1) // blk
        ang = np.linspace(-180, 180, ncalls).astype(np.float32)
        yr, lat, _ = O.lv2_render(so, xm, 48000, blk, ang[:, None])
        assert lat == 1792
        assert np.max(np.abs(O.oracle_plugin_run(xm, 48000, blk, ang) - yr[0])) < TOL


@needs_ref
def test_reference_cli_reads_pcm_files_like_float_files(oracle_built, tmp_path):
    """The stand-in libsndfile converts 16/24/32-bit PCM the way libsndfile does (sample / 2^15, / 2^31),
    so the unmodified reference CLI reports the same angles for a PCM file and for the float file
    holding the same sample values - the property phaserot_sweep_pcm is tested against on the GPU."""
    import subprocess
    x = O.two_sine(48000, 0.5, 2)
    exe = os.path.join(O.REF_DIR, "phase-rotate")
    for bits, scale in [(16, 32768.0), (24, 8388608.0), (32, 2147483648.0)]:
        q = np.clip(np.round(x.astype(np.float64) * scale), -scale, scale - 1).astype(np.int64)
        wf, wp = str(tmp_path / f"f{bits}.wav"), str(tmp_path / f"p{bits}.wav")
        O.write_wav_f32(wf, (q.astype(np.float64) / scale).astype(np.float32), 48000)
        O.write_wav_pcm(wp, q.astype(np.int32) if bits > 16 else q.astype(np.int16), 48000, bits)
        a = subprocess.run([exe, "-f", "1024", "-s", "4", wf], capture_output=True, text=True)
        b = subprocess.run([exe, "-f", "1024", "-s", "4", wp], capture_output=True, text=True)
        assert a.returncode == 0 and b.returncode == 0, (a.stderr, b.stderr)
        assert a.stdout == b.stdout and "Phase" in a.stdout, bits
