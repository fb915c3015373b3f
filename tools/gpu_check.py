"""GPU diagnostic: print parity numbers of the CUDA path against the oracle for
many cases in one run (development aid; the asserted versions live in tests/)."""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402
from phaserotate.lv2_b200 import capi  # noqa: E402


def rel(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-30)))


def section(name):
    print(f"\n=== {name} ===", flush=True)


def run(name, fn):
    try:
        fn()
    except Exception:
        print(f"!! {name} FAILED")
        traceback.print_exc()


def sweep_cases():
    section("sweep parity vs oracle")
    for nm, x, L in [
        ("two_sine 2ch 3s L8192", O.two_sine(48000, 3.0, 2), 8192),
        ("pink mono 2.5s L8192", O.pink_noise(120000, 7)[:, None], 8192),
        ("programme 2ch 2s L4096", O.programme(48000, 2.0, 2), 4096),
        ("short 1000 frames 2ch L8192", O.two_sine(48000, 1000 / 48000, 2), 8192),
        ("two_sine 3ch 1s L16384 (96k style)", O.two_sine(96000, 1.0, 3), 16384),
        ("odd length 2ch L1024", O.two_sine(48000, 0.5, 2)[:23999], 1024),
    ]:
        po = O.oracle_analyze(x, L)
        for flags, fn in [(0, "pruned"), (capi.FLAG_NO_PRUNE, "brute ")]:
            with capi.Phaserot(n_channels=x.shape[1], blksiz=L, flags=flags) as h:
                t0 = time.time()
                h.sweep(x)
                pg = h.peaks()
                dt = time.time() - t0
                st = h.stats()
            print(f"{nm:38s} {fn} max rel {rel(pg, po):.3e} argmin gpu {pg.argmin(1)} oracle {po.argmin(1)}"
                  f" surv {st['points_evaluated']}/{st['points_total']} launches {st['kernel_launches']} {dt*1e3:.1f} ms")
            if rel(pg, po) > 1e-5:
                bad = np.argwhere(np.abs(pg - po) / np.maximum(po, 1e-30) > 1e-5)
                print("   first bad (c,a):", bad[:8].tolist(), "gpu", pg[tuple(bad[0])], "oracle", po[tuple(bad[0])])
    section("sweep: reference coarse/refine call patterns")
    x = O.two_sine(48000, 2.0, 2)
    with capi.Phaserot(n_channels=2, blksiz=8192) as h:
        for (s, e, st, ch) in [(0, 360, 24, -1), (-12, 13, 1, -1), (36, 61, 1, 1), (348, 373, 1, 0)]:
            h.reset()
            h.sweep(x, s, e, st, ch)
            pg = h.peaks()
            po = O.oracle_analyze(x, 8192, 2, s, e, st, ch)
            print(f"range ({s},{e},{st},chn {ch}): max rel {rel(pg, po):.3e} nonzero gpu {np.count_nonzero(pg)} oracle {np.count_nonzero(po)}")
    section("sweep: subsample 10 (0.1 deg) vs oracle")
    x = O.programme(48000, 1.0, 2)
    po = O.oracle_analyze(x, 8192, 10)
    with capi.Phaserot(n_channels=2, blksiz=8192, subsample=10) as h:
        h.sweep(x)
        pg = h.peaks()
        print("S=10 max rel", rel(pg, po), "argmin", pg.argmin(1), po.argmin(1), h.stats())
    section("streaming analyze() drop-in")
    x = O.two_sine(48000, 1.0, 2)
    L = 8192
    nblk = (x.shape[0] + L - 1) // L
    xp = np.zeros(((nblk + 1) * L, 2), np.float32)
    xp[: x.shape[0]] = x
    with capi.Phaserot(n_channels=2, blksiz=L) as h:
        for b in range(nblk + 1):
            h.analyze(xp[b * L:(b + 1) * L], 0, 360, 1, -1, b == 0)
        pg = h.peaks()
    print("analyze stream max rel", rel(pg, O.oracle_analyze(x, L)))


def render_cases():
    section("render parity vs oracle")
    for nm, x, L, ang in [
        ("two_sine 2ch 1.3s L8192", O.two_sine(48000, 1.3, 2), 8192, [37, 181]),
        ("pink mono L8192", O.pink_noise(100000, 3)[:, None], 8192, [180]),
        ("programme 2ch L2048", O.programme(48000, 0.7, 2), 2048, [-45, 359]),
    ]:
        yo = O.oracle_apply(x, L, ang, 1)
        with capi.Phaserot(n_channels=x.shape[1], blksiz=L) as h:
            yg = h.render(x, ang, 1)
            print(f"{nm:30s} bulk max abs diff {np.abs(yg - yo).max():.3e} (scale {np.abs(yo).max():.3f})")
            # block streaming apply()
            nblk = yo.shape[0] // L
            xp = np.zeros((nblk * L, x.shape[1]), np.float32)
            xp[: x.shape[0]] = x
            ys = np.concatenate([h.apply(xp[b * L:(b + 1) * L].copy(), ang) for b in range(min(nblk, 6))])
            print(f"{'':30s} apply() stream max abs diff {np.abs(ys - yo[: ys.shape[0]]).max():.3e}")


def plugin_cases():
    section("plugin process parity vs oracle")
    for rate, blk, n in [(48000, 1024, 40000), (48000, 333, 20000), (96000, 1024, 40000), (192000, 4096, 60000), (48000, 64, 6000)]:
        x = O.pink_noise(n, 42)
        ncalls = (n + blk - 1) // blk
        ang = np.full(ncalls, 90.0, np.float32)
        ang[ncalls // 2:] = -135.0
        yo = O.oracle_plugin_run(x, rate, blk, ang)
        with capi.Phaserot(mode=capi.MODE_PLUGIN, n_channels=1, sample_rate=rate) as h:
            lat = h.latency()
            t0 = time.time()
            yg = np.concatenate([h.process(x[None, i * blk:(i + 1) * blk], ang[i])[0] for i in range(ncalls)])
            dt = time.time() - t0
        print(f"rate {rate} block {blk}: latency {lat} max abs diff {np.abs(yg - yo).max():.3e} ({ncalls} calls, {dt / ncalls * 1e6:.1f} us/call)")
    # bulk call (FFT path) and stereo
    for rate, n in [(48000, 300000), (96000, 200000)]:
        x = np.stack([O.pink_noise(n, 1), O.pink_noise(n, 2)])
        yo = np.stack([O.oracle_plugin_run(x[c], rate, n, np.array([90.0 if c == 0 else -30.0], np.float32)) for c in range(2)])
        with capi.Phaserot(mode=capi.MODE_PLUGIN, n_channels=2, sample_rate=rate) as h:
            yg = h.process(x, [90.0, -30.0])
        print(f"bulk stereo rate {rate} n {n}: max abs diff {np.abs(yg - yo).max():.3e}")
    # bulk then small continuation
    rate, n1, blk = 48000, 100000, 1024
    x = O.pink_noise(n1 + 10 * blk, 5)
    calls = [n1] + [blk] * 10
    angs = np.array([45.0] + [45.0] * 5 + [-90.0] * 5, np.float32)
    # oracle needs a fixed block: emulate by per-call angle list over variable blocks -> use the LV2 reference if present
    ref_so = os.path.join(O.REF_DIR, "phaserotate_ref.so")
    with capi.Phaserot(mode=capi.MODE_PLUGIN, n_channels=1, sample_rate=rate) as h:
        outs, pos = [], 0
        for k, nn in enumerate(calls):
            outs.append(h.process(x[None, pos:pos + nn], angs[k])[0])
            pos += nn
        yg = np.concatenate(outs)
    # oracle with block = 1024 after the first bulk part is equivalent iff the bulk part is a multiple of 1024? no: emulate with gcd block
    g = 32
    per = np.concatenate([np.full(nn // g, angs[k], np.float32) for k, nn in enumerate(calls)])
    assert all(nn % g == 0 for nn in calls)
    yo = O.oracle_plugin_run(x, rate, g, per)
    print(f"bulk+small mixed: max abs diff {np.abs(yg - yo).max():.3e}  (note: oracle emulated with {g}-frame calls)")


if __name__ == "__main__":
    which = sys.argv[1:] or ["sweep", "render", "plugin"]
    O.build_oracle()
    t0 = time.time()
    if "sweep" in which:
        run("sweep", sweep_cases)
    if "render" in which:
        run("render", render_cases)
    if "plugin" in which:
        run("plugin", plugin_cases)
    print(f"\ntotal {time.time() - t0:.1f}s")
