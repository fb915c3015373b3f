"""Generate tests/golden/*.npz from the REFERENCE build (oracle/_ref: the
reference's own sources compiled unmodified, see oracle/Makefile target `ref`).

Run in the authoring container only (needs /root/reference to build oracle/_ref):
    python tests/golden/make_golden.py
The fixtures pin the oracle (tests/test_oracle.py) and the CUDA path
(tests/test_gpu_parity.py) to reference outputs on machines where
/root/reference does not exist.
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402


def main():
    O.build_oracle()
    O.build_ref()
    assert O.have_ref(), "oracle/_ref missing"
    rng = np.random.default_rng(2024)

    # --- CLI analysis: stereo, 0.5 s @ 48 kHz, blksiz 2048, full 360-index table + coarse/refine style passes
    n = 24000
    t = np.arange(n) / 48000.0
    x = np.stack([
        0.5 * np.sin(2 * np.pi * 110 * t) + 0.25 * np.sin(2 * np.pi * 1760.3 * t) + 0.05 * rng.standard_normal(n),
        0.4 * np.sin(2 * np.pi * 220 * t + 1.0) + 0.3 * np.sin(2 * np.pi * 3300.7 * t) + 0.05 * rng.standard_normal(n),
    ], 1).astype(np.float32)
    L = 2048
    full, _ = O.ref_analyze(x, L, 0, 360, 1)
    coarse, _ = O.ref_analyze(x, L, 0, 360, 24)
    refine, _ = O.ref_analyze(x, L, -12, 13, 1)
    single, _ = O.ref_analyze(x, L, 36, 61, 1, only_chn=1)
    s = np.zeros(360, np.float32)
    c = np.zeros(360, np.float32)
    O.ref_cli().ref_cli_lut(s, c)
    taps = np.zeros(L, np.float32)
    O.ref_cli().ref_cli_taps(L, taps)
    np.savez_compressed(os.path.join(HERE, "cli_analyze.npz"), x=x, blksiz=L, full=full, coarse=coarse, refine=refine, single=single,
                        lut_sin=s, lut_cos=c, taps=taps)

    # --- CLI render: PhaseRotate::apply stream and the file loop (binary) incl. R1/R2 quirks
    xr = x[:9000]
    angles = np.array([37, 181], np.int32)
    stream, _ = O.ref_apply(xr, L, angles, 1)
    outs = {}
    for nm, sig, ang in [("stereo", xr, "18.5,90.5"), ("mono", xr[:, :1].copy(), "18.5"), ("short", xr[:700], "18.5,90.5"), ("exact", x[:4 * L], "18.5,90.5")]:
        O.write_wav_f32("/tmp/_g_in.wav", sig, 48000)
        subprocess.run([os.path.join(O.REF_DIR, "phase-rotate"), "-f", str(L), "-a", ang, "/tmp/_g_in.wav", "/tmp/_g_out.wav"], check=True)
        y, _ = O.read_wav_f32("/tmp/_g_out.wav")
        outs["file_" + nm] = y
    np.savez_compressed(os.path.join(HERE, "cli_render.npz"), x=xr, x_exact=x[:4 * L], blksiz=L, angles=angles, stream=stream, **outs)

    # --- CLI text output (search logic) on a 1 s stereo file, default options and a few variants
    xs = O.harmonic(48000, 1.0, 2)
    O.write_wav_f32("/tmp/_g_cli.wav", xs, 48000)
    texts = {}
    for key, argv in {"default": [], "stride2": ["-s", "2"], "stride1": ["-s", "1"], "link": ["-l"], "f4096_s6": ["-f", "4096", "-s", "6"]}.items():
        r = subprocess.run([os.path.join(O.REF_DIR, "phase-rotate")] + argv + ["/tmp/_g_cli.wav"], check=True, capture_output=True, text=True)
        texts[key] = r.stdout
    np.savez_compressed(os.path.join(HERE, "cli_text.npz"), x=xs, **{k: np.array(v) for k, v in texts.items()})

    # --- plugin: mono pink, angle schedule with two ramps, three rates
    ref_so = os.path.join(O.REF_DIR, "phaserotate_ref.so")
    plug = {}
    for rate, blk, n in [(48000, 256, 8000), (48000, 1000, 8000), (96000, 1024, 10000), (192000, 2048, 16000)]:
        xm = O.pink_noise(n, 11)
        ncalls = (n + blk - 1) // blk
        ang = np.full((ncalls, 1), 90.0, np.float32)
        ang[ncalls // 2:] = -135.0
        y, lat, _ = O.lv2_render(ref_so, xm, rate, blk, ang)
        key = f"r{rate}_b{blk}"
        plug[key + "_x"] = xm
        plug[key + "_ang"] = ang[:, 0].copy()
        plug[key + "_y"] = y[0]
        plug[key + "_lat"] = np.float32(lat)
    # stereo instance
    xs2 = np.stack([O.pink_noise(8000, 21), O.pink_noise(8000, 22)])
    ang2 = np.tile(np.array([[45.0, -90.0]], np.float32), (8, 1))
    y2, lat2, _ = O.lv2_render(ref_so, xs2, 48000, 1000, ang2)
    plug["stereo_x"], plug["stereo_ang"], plug["stereo_y"] = xs2, ang2, y2
    np.savez_compressed(os.path.join(HERE, "plugin.npz"), **plug)
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
