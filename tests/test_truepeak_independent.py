"""Oversampled true-peak (cfg.oversample = 2 / 4) pinned against an INDEPENDENT float64 computation.

True-peak is not a reference feature: round 1 checked the CUDA path only against this repository's
own CPU restatement (oracle/phaserot_oracle.c: pro_cli_analyze_tp_shard), i.e. the same author's reading of
the definition twice.  Here the definition in include/phaserot_cuda.h is written out a third time with nothing
shared: numpy/scipy in float64, the Hilbert FIR from its closed form, the interpolator coefficients typed from
ITU-R BS.1770-4 Annex 2 (table "Filter coefficients", four phases of the 48-tap FIR), scipy.signal.upfirdn for
the 4x interpolation.  The CPU test pins the oracle, the GPU test the library.  Tolerance 1e-5 relative
(north_star's bound for per-angle peaks; fp32 path against float64).
"""
import numpy as np
import pytest
import scipy.signal as sg

import oracle_lib as O

# ITU-R BS.1770-4, Annex 2, "Filter coefficients" for 4x oversampling: phase 0 .. phase 3, 12 taps each.
ITU_PHASES = np.array([
    [0.0017089843750, 0.0109863281250, -0.0196533203125, 0.0332031250000, -0.0594482421875, 0.1373291015625,
     0.9721679687500, -0.1022949218750, 0.0476074218750, -0.0266113281250, 0.0148925781250, -0.0083007812500],
    [-0.0291748046875, 0.0292968750000, -0.0517578125000, 0.0891113281250, -0.1665039062500, 0.4650878906250,
     0.7797851562500, -0.2003173828125, 0.1015625000000, -0.0582275390625, 0.0330810546875, -0.0189208984375],
    [-0.0189208984375, 0.0330810546875, -0.0582275390625, 0.1015625000000, -0.2003173828125, 0.7797851562500,
     0.4650878906250, -0.1665039062500, 0.0891113281250, -0.0517578125000, 0.0292968750000, -0.0291748046875],
    [-0.0083007812500, 0.0148925781250, -0.0266113281250, 0.0476074218750, -0.1022949218750, 0.9721679687500,
     0.1373291015625, -0.0594482421875, 0.0332031250000, -0.0196533203125, 0.0109863281250, 0.0017089843750],
])


def hilbert_taps(L):
    """cli/phase-rotate.cc:144-161 in closed form: -(2/L) cot(pi (n - L/2) / L) at odd n - L/2, Hann window 0.5 (1 - cos(2 pi n / L))."""
    n = np.arange(L)
    m = n - L // 2
    t = np.zeros(L)
    odd = (m & 1) == 1
    t[odd] = -2.0 / np.tan(np.pi * m[odd] / L) * (0.5 / L) * (1.0 - np.cos(2.0 * np.pi * n[odd] / L))
    return t


def interpolated(s, os_):
    """[1 + os_ phases... ] rows: the sample itself and s^[t, ph] = sum_k c[ph][k] s[t - k] (samples before 0 are zero).
    Done as a 4x polyphase up-sampling with the 48-tap prototype h[4 k + ph] = c[ph][k]: output 4 t + ph is phase ph at time t."""
    proto = ITU_PHASES.T.reshape(-1)                      # h[4k + ph]
    up = sg.upfirdn(proto, s, up=4)[:4 * len(s)]          # y[4 t + ph]
    ph = up.reshape(len(s), 4).T                          # [phase][t]
    use = [0, 1, 2, 3] if os_ == 4 else [0, 2]
    return np.vstack([s[None, :]] + [ph[p][None, :] for p in use])


def true_peak_table(x, L, S, os_):
    """peak[c][a] per the definition in include/phaserot_cuda.h, float64."""
    F, C = x.shape
    D = L // 2
    B = -(-F // L)
    T = (B + 1) * L                                        # B real blocks + one zero flush block (cli:572-586)
    MS = 180 * S
    taps = hilbert_taps(L)
    out = np.zeros((C, MS))
    ang = -np.arange(MS) * np.pi / MS                      # SinCosLut: sin / cos (-i / S degrees), cli:41-72
    for c in range(C):
        xc = np.zeros(T)
        xc[:F] = x[:, c].astype(np.float64)
        H = sg.fftconvolve(xc, taps)[:T]                   # H[t] = sum_k fir[k] x[t - k]
        xd = np.concatenate([np.zeros(D), xc])[:T]         # x[t - L/2]
        xd[:L] = 0.0                                       # first-block rule (cli:418-419): zero history under the first block
        xi, hi = interpolated(xd, os_), interpolated(H, os_)
        xi, hi = xi[:, D:], hi[:, D:]                      # samples the reference examines: t >= L/2
        for a in range(1, MS):
            out[c, a] = np.abs(np.cos(ang[a]) * xi + np.sin(ang[a]) * hi).max()
        out[c, 0] = np.abs(interpolated(xc[:B * L], os_)).max()   # un-wrapped angle 0: the same detector on the raw input (cli:413-414)
    return out


CASES = {
    "two_sine_L8192": (lambda: O.two_sine(48000, 1.5, 2), 8192),
    "hf_tone_L4096": (lambda: (0.6 * np.sin(2 * np.pi * 11025.0 / 48000.0 * np.arange(40000) + 0.7)).astype(np.float32)[:, None], 4096),
    "programme_L2048": (lambda: O.programme(48000, 0.8, 2), 2048),
}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("os_", [2, 4])
def test_oracle_true_peak_matches_independent_float64(name, os_):
    gen, L = CASES[name]
    x = gen()
    want = true_peak_table(x, L, 2, os_)
    got = O.oracle_analyze_tp(x, L, oversample=os_)
    assert np.max(np.abs(got - want) / want) <= 1e-5
    # a tone at fs / 4.35 peaks between the samples: the detector must see more than the digital peak
    if name == "hf_tone_L4096":
        dig = O.oracle_analyze(x, L)
        assert np.all(got[:, 1:] >= dig[:, 1:]) and got[0, 90] > 1.01 * dig[0, 90] if os_ == 4 else True


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("os_", [2, 4])
def test_gpu_true_peak_matches_independent_float64(name, os_):
    from phaserotate.lv2_b200 import capi
    gen, L = CASES[name]
    x = gen()
    want = true_peak_table(x, L, 2, os_)
    with capi.Phaserot(n_channels=x.shape[1], blksiz=L, oversample=os_) as h:
        h.sweep(x)
        got = h.peaks()
    assert np.max(np.abs(got - want) / want) <= 1e-5
