"""GPU tests of the round-2 additions to the C ABI (v4): host-memory shards, packed 24-bit ingest,
device groups (several GPUs in one process), the plugin's angle state, tables of odd size, handles on
two devices, and the CLI's opt-in long options (--subsample, --gpus, --fixed-write).
Run with `-m gpu` on a B200."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from phaserotate.lv2_b200 import build, capi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _built():
    build.build_library()
    build.build_host()
    O.build_oracle()


def _cli(*argv, env=None):
    exe = os.path.join(build.BIN_DIR, "phase-rotate")
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([exe] + list(argv), capture_output=True, text=True, env=e)


def _n_devices():
    import torch
    return torch.cuda.device_count()


def pack24(x):
    """float [-1, 1) -> (packed 24-bit little-endian bytes, the floats sf_readf_float would return)."""
    q = np.clip(np.round(x * 8388608.0), -8388608, 8388607).astype(np.int32)
    b = np.empty(q.shape + (3,), np.uint8)
    b[..., 0] = q & 0xff
    b[..., 1] = (q >> 8) & 0xff
    b[..., 2] = (q >> 16) & 0xff
    return b.reshape(q.shape[0], -1), (q.astype(np.float64) / 8388608.0).astype(np.float32)


def test_odd_sized_tables_read_back():
    """a table of odd length (mono, stride 8 -> 44 slots; a refine range through angle 0)
    used to put the 16-byte statistics copy 4 bytes past the pinned result buffer."""
    x = O.pink_noise(60000, 3)[:, None]
    po = O.oracle_analyze(x, 8192)
    with capi.Phaserot(n_channels=1, blksiz=8192) as h:
        h.sweep(x, 0, 360, 8)
        pk = h.peaks()
        idx = np.arange(0, 360, 8)
        assert np.max(np.abs(pk[0, idx] - po[0, idx]) / po[0, idx]) <= 1e-5
        h.reset()
        h.sweep(x, -3, 4, 1)          # mono refine window through index 0 (the loop stops at `end`): raw peak + 6 slots + 1
        pk = h.peaks()
        for a in (-3, -2, -1, 0, 1, 2, 3):
            assert abs(pk[0, a % 360] - po[0, a % 360]) <= 1e-5 * po[0, a % 360], a
        st = h.stats()
    assert st["d2h_bytes"] > 0


def test_host_shards_equal_device_shards_and_single_pass():
    """phaserot_sweep_shard (host memory, chunked overlapped upload) == phaserot_sweep_shard_device ==
    the single pass, bit for bit on the segment grid; float and int16 sources."""
    import torch
    x = O.programme(48000, 8.0, 2)
    q16 = np.clip(np.round(x * 32768.0), -32768, 32767).astype(np.int16)
    xf = (q16.astype(np.float32) / np.float32(32768.0)).astype(np.float32)
    with capi.Phaserot(n_channels=2, blksiz=8192, subsample=10) as h:
        h.sweep(xf)
        whole = h.peaks()
        al = h.shard_align()
        cut = al * ((xf.shape[0] // 2) // al)
        assert 0 < cut < xf.shape[0]
        for fmt, src in ((capi.PCM_F32, xf), (capi.PCM_S16, q16)):
            h.reset()
            h.sweep_shard(np.ascontiguousarray(src[:cut]), cut, None, True, False, fmt=fmt)
            a = h.peaks()
            h.reset()
            h.sweep_shard(np.ascontiguousarray(src[cut:]), src.shape[0] - cut, xf[cut - 8192:cut], False, True, fmt=fmt)
            b = h.peaks()
            assert np.array_equal(np.maximum(a, b), whole), fmt
        xd = torch.from_numpy(xf).cuda()
        h.reset()
        h.sweep_shard_device(xd[cut:].data_ptr(), xf.shape[0] - cut, xf[cut - 8192:cut], False, True)
        assert np.array_equal(h.peaks(), b)
        # a non-final shard must be whole blocks
        with pytest.raises(capi.PhaserotError):
            h.sweep_shard(xf[:1000], 1000, None, True, False)


def test_packed_24_bit_ingest_is_bit_identical():
    x = O.harmonic(48000, 2.5, 2)[:119997]
    b24, xf = pack24(x)
    with capi.Phaserot(n_channels=2, blksiz=8192) as h:
        h.sweep(xf)
        ref = h.peaks()
        h.reset()
        h.sweep_pcm(np.ascontiguousarray(b24).reshape(-1))
        assert np.array_equal(h.peaks(), ref)
    # ragged mono length (not a multiple of 4 samples)
    b1, x1 = pack24(O.pink_noise(30001, 9)[:, None])
    with capi.Phaserot(n_channels=1, blksiz=4096) as h:
        h.sweep(x1)
        ref = h.peaks()
        h.reset()
        h.sweep_pcm(np.ascontiguousarray(b1).reshape(-1))
        assert np.array_equal(h.peaks(), ref)


@pytest.mark.parametrize("ndev", [2, 3])
def test_device_group_equals_single_device(ndev):
    """phaserot_group_sweep: sample-range shards over the devices of one process, tables combined on
    the first device through peer memory.  On a one-GPU box the group is formed from handles on the same
    device, which exercises the same sharding / combine code."""
    have = _n_devices()
    devs = list(range(ndev)) if have >= ndev else [0] * ndev
    x = O.programme(48000, 9.0, 2)
    with capi.Phaserot(n_channels=2, blksiz=8192, subsample=10) as h:
        h.sweep(x)
        whole = h.peaks()
    with capi.PhaserotGroup(ndev, devs, n_channels=2, blksiz=8192, subsample=10) as g:
        assert g.size() == ndev
        g.sweep(x)
        assert np.array_equal(g.peaks(), whole)
        g.reset()
        q16 = np.clip(np.round(x * 32768.0), -32768, 32767).astype(np.int16)
        g.sweep(q16, fmt=capi.PCM_S16)
        got = g.peaks()
    with capi.Phaserot(n_channels=2, blksiz=8192, subsample=10) as h:
        h.sweep_pcm(q16)
        assert np.array_equal(got, h.peaks())
    # a file shorter than one shard per device uses fewer devices
    with capi.PhaserotGroup(ndev, devs, n_channels=2, blksiz=8192) as g:
        s = O.two_sine(48000, 0.2, 2)
        g.sweep(s)
        po = O.oracle_analyze(s, 8192)
        assert np.max(np.abs(g.peaks() - po) / po) <= 1e-5


def test_handles_on_two_devices_in_one_process():
    """cudaFuncSetAttribute is per device; a second handle on another device of the same
    process must be able to launch the 130 KB shared-memory kernels."""
    if _n_devices() < 2:
        pytest.skip("needs two GPUs")
    x = O.programme(48000, 3.0, 2)
    res = []
    for d in (0, 1):
        with capi.Phaserot(n_channels=2, blksiz=8192, device=d) as h:
            h.sweep(x)
            res.append(h.peaks())
    assert np.array_equal(res[0], res[1])
    with capi.Phaserot(mode=capi.MODE_PLUGIN, n_channels=1, sample_rate=48000.0, device=1) as hp:
        y = hp.process(O.pink_noise(4096, 1)[None, :], 90.0)
        assert np.isfinite(y).all()


def test_plugin_angle_state_follows_the_ramp():
    """phaserot_plugin_angle = Channel::angle (src/phaserotate.c:53): 0 after create, moving towards the
    target by at most P * 1e-6 turns per sample while ramping, equal to the target when arrived."""
    with capi.Phaserot(mode=capi.MODE_PLUGIN, n_channels=2, sample_rate=48000.0) as h:
        assert np.array_equal(h.plugin_angle(), [0.0, 0.0])
        x = np.zeros((2, 256), np.float32)
        h.process(x, [180.0, 0.0])
        a = h.plugin_angle()
        assert a[1] == 0.0 and -0.5 < a[0] < 0.0
        assert abs(a[0] + 256 * 256e-6) < 1e-6         # one partition at the slew limit (src:295, 688-693)
        for _ in range(12):
            h.process(x, [180.0, 0.0])
        assert h.plugin_angle()[0] == np.float32(-0.5)


def test_cli_subsample_option(tmp_path):
    """--subsample N: the report is computed on the 1/N degree grid (same text shape); N = 2 is the
    reference grid; the GPU table behind it is the library's (checked against the oracle at N = 10)."""
    x = O.harmonic(48000, 2.0, 2)
    wav = str(tmp_path / "h.wav")
    O.write_wav_f32(wav, x, 48000)
    a, b = _cli("-s", "1", wav), _cli("--subsample", "2", "-s", "1", wav)
    assert a.returncode == 0 and a.stdout == b.stdout
    r = _cli("--subsample", "10", "-s", "1", wav)
    assert r.returncode == 0, r.stderr
    po = O.oracle_analyze(x, 8192, subsample=10)
    lines = [l for l in r.stdout.splitlines() if l.startswith("Channel:")]
    assert len(lines) == 2
    for c, l in enumerate(lines):
        deg = float(l.split("Phase:")[1].split("deg")[0])
        idx = int(round(deg * 10)) % 1800
        # the reported angle is a minimum of the oracle's table within tolerance
        assert po[c, idx] <= po[c].min() * (1 + 1e-5), (c, deg)
    assert _cli("--subsample", "0", wav).returncode == 1


def test_cli_gpus_option(tmp_path):
    x = O.programme(48000, 6.0, 2)
    wav = str(tmp_path / "p.wav")
    O.write_wav_f32(wav, x, 48000)
    one = _cli("-s", "2", wav)
    env = {} if _n_devices() >= 2 else {"PHASEROT_GROUP_DEVICES": "0,0"}
    two = _cli("--gpus", "2", "-s", "2", wav, env=env)
    assert one.returncode == 0 and two.returncode == 0, two.stderr
    assert one.stdout == two.stdout
    q16 = np.clip(np.round(x * 32768.0), -32768, 32767).astype(np.int16)
    wp = str(tmp_path / "p16.wav")
    O.write_wav_pcm(wp, q16, 48000, 16)
    assert _cli("-s", "2", wp).stdout == _cli("--gpus", "2", "-s", "2", wp, env=env).stdout
    q24 = np.clip(np.round(x * 8388608.0), -8388608, 8388607).astype(np.int32)
    w24 = str(tmp_path / "p24.wav")
    O.write_wav_pcm(w24, q24, 48000, 24)
    wf = str(tmp_path / "p24f.wav")
    O.write_wav_f32(wf, (q24.astype(np.float64) / 8388608.0).astype(np.float32), 48000)
    assert _cli("-s", "1", w24).stdout == _cli("-s", "1", wf).stdout     # packed 24-bit ingest (sf_read_raw)


def test_cli_fixed_write(tmp_path):
    """--fixed-write: the output is y[t + L/2], t in [0, F) for every channel - no float-offset trim
    (cli:985), no stale tail (cli:973).  Mono output without the flag is already that, except for the tail."""
    L = 2048
    x = O.harmonic(48000, 0.75, 2)[:36000 - 700]      # short last block longer than the latency (R2 case)
    wav, out = str(tmp_path / "x.wav"), str(tmp_path / "y.wav")
    O.write_wav_f32(wav, x, 48000)
    r = _cli("--fixed-write", "-f", str(L), "-a", "18.5,90.5", wav, out)
    assert r.returncode == 0, r.stderr
    y, _ = O.read_wav_f32(out)
    assert y.shape == x.shape
    full = O.oracle_apply(x, L, [37, 181], 1)           # every block + one flush block, no trim
    want = full[L // 2:L // 2 + x.shape[0]]
    assert np.max(np.abs(y - want)) <= 1e-5 * float(np.abs(want).max())
    # the reference-compatible loop differs on the same input (stereo: first block offset in floats)
    r = _cli("-f", str(L), "-a", "18.5,90.5", wav, str(tmp_path / "q.wav"))
    yq, _ = O.read_wav_f32(str(tmp_path / "q.wav"))
    assert yq.shape != y.shape or np.max(np.abs(yq - y)) > 1e-3


def _tones(seconds, parts, sr=48000):
    t = np.arange(int(seconds * sr), dtype=np.float64) / sr
    x = np.zeros((t.size, 2), np.float32)
    for c in range(2):
        acc = np.zeros(t.size)
        for amp, f, ph in parts:
            acc += amp * np.sin(2 * np.pi * f * t + ph[c])
        x[:, c] = acc
    return x


def test_dense_mode_constant_envelope_equals_brute_force():
    """A pure sine: every sample lies on the hull of the (x_d, H) point set, the radius filter keeps
    them all and the survivor list (sized for programme material) overflows.  The pass is repeated in
    dense mode (sector thresholds + per-point angle windows) and the table still equals brute force bit
    for bit; the handle stays dense for the same material and returns to normal mode on programme."""
    import torch
    x = _tones(100.0, [(0.5, 440.0, [0.0, 1.0])])     # 4.8 M points per channel against a list of 1.8 M
    xd = torch.from_numpy(x).cuda()
    with capi.Phaserot(n_channels=2, blksiz=8192, subsample=10, flags=capi.FLAG_NO_PRUNE) as hb:
        hb.sweep_device(xd.data_ptr(), x.shape[0])
        brute = hb.peaks()
    with capi.Phaserot(n_channels=2, blksiz=8192, subsample=10) as h:
        h.sweep_device(xd.data_ptr(), x.shape[0])
        got = h.peaks()
        st = h.stats()
        assert st["dense_repeats"] == 1
        assert np.array_equal(got, brute)
        assert st["points_evaluated"] < 0.5 * st["points_total"]      # the bootstrap wave is swept at every angle (<= 1 M points per channel), the rest in windows
        h.reset()
        h.sweep_device(xd.data_ptr(), x.shape[0])                     # sticky: no second repeat
        assert np.array_equal(h.peaks(), brute)
        assert h.stats()["dense_repeats"] == 1
        # host source in dense mode, sharded, and the reference grid
        h.reset()
        al = h.shard_align()
        cut = al * ((x.shape[0] // 2) // al)
        h.sweep_shard(np.ascontiguousarray(x[:cut]), cut, None, True, False)
        a = h.peaks()
        h.reset()
        h.sweep_shard(np.ascontiguousarray(x[cut:]), x.shape[0] - cut, x[cut - 8192:cut], False, True)
        assert np.array_equal(np.maximum(a, h.peaks()), brute)
        # programme material: the handle leaves dense mode again
        prog = O.programme(48000, 60.0, 2)
        h.reset()
        h.sweep(prog)
        p1 = h.peaks()
        h.reset()
        h.sweep(prog)
        assert np.array_equal(h.peaks(), p1)
        assert h.stats()["dense_repeats"] == 1
    with capi.Phaserot(n_channels=2, blksiz=8192, subsample=10, flags=capi.FLAG_NO_PRUNE) as hb:
        hb.sweep(prog)
        assert np.array_equal(p1, hb.peaks())
    # reference grid, against the oracle (3 s)
    s3 = x[:144000]
    po = O.oracle_analyze(s3, 8192)
    with capi.Phaserot(n_channels=2, blksiz=8192) as h:
        h.sweep(s3)
        assert np.max(np.abs(h.peaks() - po) / po) <= 1e-5


@pytest.mark.parametrize("subsample,nch", [(1, 2), (4, 1), (10, 2), (20, 2), (50, 1)])
@pytest.mark.parametrize("material", ["faded_sine", "two_tone", "chirp"])
def test_dense_mode_grids_and_envelopes(subsample, nch, material):
    """Dense mode over the grids that select the different window kernels (0.05 degree and coarser: the
    walk kernel with a first window of 1 / 2 / 3 / 6 angles either side; 0.02 degree: tables too large for
    shared memory, the sector-window kernel) and over materials that stress different parts of it: a sine
    that fades in and out (flat peak table: every point within 1e-6 of every threshold, the first window is
    all there is), BASELINE config 1's two tones (interior points well above the smallest threshold: the
    walk from sector to sector), a slow chirp (constant envelope, directions not periodic).  Long enough to
    overflow the survivor list (mono needs more than the 148-segment first wave); every table equals brute
    force bit for bit."""
    import torch
    sr  = 48000
    secs = 100.0 if nch == 2 else 240.0
    n = int(sr * secs)
    t = np.arange(n, dtype=np.float64) / sr
    if material == "faded_sine":
        parts = [(0.5 * np.minimum(1.0, np.minimum(t, secs - t) / 2.0), 2 * np.pi * 440.0 * t, 0.9)]
    elif material == "two_tone":
        parts = [(0.5, 2 * np.pi * 110.0 * t, 1.0), (0.25, 2 * np.pi * 1760.3 * t, 0.0)]
    else:
        parts = [(0.5, 2 * np.pi * (300.0 * t + 0.5 * (200.0 / secs) * t * t), 0.9)]      # 300 Hz -> 500 Hz
    def render(parts):
        return np.stack([sum(a * np.sin(ph + dph * c) for a, ph, dph in parts).astype(np.float32) for c in range(nch)], axis=1)
    x = render(parts)
    xd = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    with capi.Phaserot(n_channels=nch, blksiz=8192, subsample=subsample, flags=capi.FLAG_NO_PRUNE) as hb:
        hb.sweep_device(xd.data_ptr(), n)
        brute = hb.peaks()
    with capi.Phaserot(n_channels=nch, blksiz=8192, subsample=subsample) as h:
        if material == "two_tone":
            # a tenth of its samples survive the radius filter: not enough to overflow the list by itself, plenty
            # to keep a handle dense that is (sticky until the lists hold < 0.1 % of the samples)
            sd = torch.from_numpy(render([(0.5, 2 * np.pi * 440.0 * t, 0.9)])).cuda()
            h.sweep_device(sd.data_ptr(), n)
            h.peaks()
            assert h.stats()["dense_repeats"] == 1
            h.reset()
        h.sweep_device(xd.data_ptr(), n)
        got = h.peaks()
        st = h.stats()
        assert np.array_equal(got, brute), st
        assert st["dense_repeats"] == 1, ("the material was meant to overflow the survivor list", st)
        h.reset()
        h.sweep_device(xd.data_ptr(), n)          # already dense: straight through the window kernels
        assert np.array_equal(h.peaks(), brute)
        assert h.stats()["dense_repeats"] == 1


def test_dense_mode_window_kernel_fallback_equals_walk_kernel(tmp_path):
    """PHASEROT_WALK=0 sends every dense-mode grid through sweep_window_kernel (the kernel finer grids and
    partial angle sets still use); the switch is read once per process, so the other setting runs in a child.
    Same table bit for bit from both kernels on a 0.1 degree grid."""
    import sys
    import torch
    x = _tones(100.0, [(0.5, 440.0, [0.0, 1.0])])
    xd = torch.from_numpy(x).cuda()
    with capi.Phaserot(n_channels=2, blksiz=8192, subsample=10) as h:
        h.sweep_device(xd.data_ptr(), x.shape[0])
        here = h.peaks()
        assert h.stats()["dense_repeats"] == 1
    out = tmp_path / "pk.npy"
    code = (
        "import sys, numpy as np, torch\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from phaserotate.lv2_b200 import capi\n"
        "import test_gpu_round2 as T\n"
        "x = T._tones(100.0, [(0.5, 440.0, [0.0, 1.0])])\n"
        "xd = torch.from_numpy(x).cuda()\n"
        "h = capi.Phaserot(n_channels=2, blksiz=8192, subsample=10)\n"
        "h.sweep_device(xd.data_ptr(), x.shape[0])\n"
        "np.save(%r, h.peaks())\n"
        "assert h.stats()['dense_repeats'] == 1\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)), str(out))
    env = dict(os.environ, PHASEROT_WALK="0")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert np.array_equal(np.load(out), here)


def test_dense_mode_two_tone_and_true_peak():
    """Config-1 material long enough to overflow the list (few-tone: a third of the samples survive the
    radius filter until the table has converged), digital and 4x true-peak: dense-mode tables equal
    brute force bit for bit."""
    import torch
    x = _tones(60.0, [(0.5, 110.0, [0.0, 1.0]), (0.25, 1760.3, [0.0, 0.0])])
    xd = torch.from_numpy(x).cuda()
    for os_ in (0, 4):
        with capi.Phaserot(n_channels=2, blksiz=8192, subsample=10, flags=capi.FLAG_NO_PRUNE, oversample=os_) as hb:
            hb.sweep_device(xd.data_ptr(), x.shape[0])
            brute = hb.peaks()
        with capi.Phaserot(n_channels=2, blksiz=8192, subsample=10, oversample=os_) as h:
            h.sweep_device(xd.data_ptr(), x.shape[0])
            assert np.array_equal(h.peaks(), brute), os_


def test_batch_render_per_track_angles():
    """BASELINE config 4 in small: a batch of stereo tracks, each rendered at its own per-channel angles
    (theta_i = (i * 37 mod 360) half degrees and its mirror) through one handle that is reset between tracks -
    host buffers and device buffers - against the oracle's render of every track."""
    import torch
    L = 8192
    tracks = [O.programme(48000, 0.7 + 0.1 * i, 2, seed=1000 + i) for i in range(5)]
    with capi.Phaserot(n_channels=2, blksiz=L) as h:
        for i, x in enumerate(tracks):
            ang = [(i * 37) % 360, (360 - i * 37) % 360]
            want = O.oracle_apply(x, L, ang, 1)
            h.reset()
            y = h.render(x, ang, 1)
            assert y.shape == want.shape
            assert np.max(np.abs(y - want)) <= 1e-5 * float(np.abs(want).max()), i
            xd = torch.from_numpy(np.ascontiguousarray(x)).cuda()
            yd = torch.empty((want.shape[0], 2), device="cuda", dtype=torch.float32)
            h.reset()
            h.render_device(xd.data_ptr(), x.shape[0], ang, 1, yd.data_ptr())
            torch.cuda.synchronize()
            assert np.array_equal(yd.cpu().numpy(), y), i


def test_sharded_overflow_protocol_e_again():
    """Two 'ranks' (two handles) sweep the halves of a pure sine; their pending device tables are combined by
    an element-wise max over the WHOLE buffers (what the NCCL max all-reduce does in place).  The overflow
    flag of either shard is part of the buffer: both ranks get PHASEROT_E_AGAIN from peaks() after having
    re-enqueued their shard in dense mode; one more combine gives the brute-force table on both."""
    import torch
    x = _tones(240.0, [(0.5, 440.0, [0.0, 1.0])])    # two shards of 5.8 M points per channel against lists of 1.8 M
    xd = torch.from_numpy(x).cuda()
    with capi.Phaserot(n_channels=2, blksiz=8192, subsample=10, flags=capi.FLAG_NO_PRUNE) as hb:
        hb.sweep_device(xd.data_ptr(), x.shape[0])
        brute = hb.peaks()
    ranks = [capi.Phaserot(n_channels=2, blksiz=8192, subsample=10) for _ in range(2)]
    try:
        al = ranks[0].shard_align()
        cut = al * ((x.shape[0] // 2) // al)
        hist = xd[cut - 8192:cut].contiguous()

        def views():
            out = []
            for h in ranks:
                ptr, nc, na = h.pending_table()
                n = nc * na + nc + 1

                class _Dev:
                    __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}
                out.append(torch.as_tensor(_Dev(), device="cuda"))
            return out

        def all_reduce_max():
            torch.cuda.synchronize()          # the handles run on their own streams
            a, b = views()
            m = torch.maximum(a, b)
            a.copy_(m)
            b.copy_(m)
            torch.cuda.synchronize()

        ranks[0].sweep_shard_device(xd.data_ptr(), cut, None, True, False)
        ranks[1].sweep_shard_device(xd[cut:].data_ptr(), x.shape[0] - cut, hist.data_ptr(), False, True)
        all_reduce_max()
        again = 0
        for h in ranks:
            with pytest.raises(capi.PhaserotError) as ei:
                h.peaks()
            assert ei.value.code == capi.E_AGAIN
            again += 1
        assert again == 2
        all_reduce_max()
        for h in ranks:
            assert np.array_equal(h.peaks(), brute)
            assert h.stats()["dense_repeats"] == 1
    finally:
        for h in ranks:
            h.close()


def test_two_phase_sharded_sweep_equals_single_pass():
    """phaserot_sweep_shard_boot_device + phaserot_sweep_shard_resume: the ranks' bootstrap waves are combined
    (max over the pending tables) before the contiguous passes, so every rank prunes with the thresholds of the
    whole stream's sample.  Same table as the single pass, bit for bit; fewer survivors than shards that each
    bootstrap on their own."""
    import torch
    x = O.programme(48000, 300.0, 2)
    x[: x.shape[0] // 2] *= 0.25                      # a quiet first half: its own bootstrap would prune poorly
    xd = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    with capi.Phaserot(n_channels=2, blksiz=8192, subsample=10) as h:
        h.sweep_device(xd.data_ptr(), x.shape[0])
        whole = h.peaks()
        al = h.shard_align()
    cut = al * ((x.shape[0] // 2) // al)
    hist = xd[cut - 8192:cut].contiguous()
    shards = [(xd.data_ptr(), cut, None, True, False), (xd[cut:].data_ptr(), x.shape[0] - cut, hist.data_ptr(), False, True)]

    def run(two_phase):
        ranks = [capi.Phaserot(n_channels=2, blksiz=8192, subsample=10) for _ in range(2)]
        try:
            def all_reduce_max():
                torch.cuda.synchronize()
                v = []
                for h in ranks:
                    ptr, nc, na = h.pending_table()

                    class _Dev:
                        __cuda_array_interface__ = {"shape": (nc * na + nc + 1,), "typestr": "<f4", "data": (ptr, False), "version": 2}
                    v.append(torch.as_tensor(_Dev(), device="cuda"))
                m = torch.maximum(v[0], v[1])
                v[0].copy_(m)
                v[1].copy_(m)
                torch.cuda.synchronize()
            if two_phase:
                for h, s in zip(ranks, shards):
                    h.sweep_shard_boot_device(*s)
                all_reduce_max()
                for h in ranks:
                    h.sweep_shard_resume()
            else:
                for h, s in zip(ranks, shards):
                    h.sweep_shard_device(*s)
            all_reduce_max()
            tabs = [h.peaks() for h in ranks]
            ev = sum(h.stats()["points_evaluated"] for h in ranks)
            return tabs, ev
        finally:
            for h in ranks:
                h.close()

    t1, ev1 = run(False)
    t2, ev2 = run(True)
    for t in t1 + t2:
        assert np.array_equal(t, whole)
    assert ev2 <= ev1
