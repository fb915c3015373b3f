"""Write a synthetic test WAV: python tools/make_wav.py out.wav seconds [kind=two_sine|programme] [bits=32f|16|24] [sr=48000] [channels=2]
two_sine is BASELINE config 1's material (110 Hz + 1760.3 Hz, 1 rad between the channels' low tones)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402

path, seconds = sys.argv[1], float(sys.argv[2])
kind = sys.argv[3] if len(sys.argv) > 3 else "two_sine"
bits = sys.argv[4] if len(sys.argv) > 4 else "32f"
sr = int(sys.argv[5]) if len(sys.argv) > 5 else 48000
ch = int(sys.argv[6]) if len(sys.argv) > 6 else 2
n = int(seconds * sr)
t = np.arange(n, dtype=np.float64) / sr
if kind == "two_sine":
    x = np.stack([0.5 * np.sin(2 * np.pi * 110.0 * t + (1.0 if c % 2 else 0.0)) + 0.25 * np.sin(2 * np.pi * 1760.3 * t) for c in range(ch)], axis=1).astype(np.float32)
else:
    sys.path.insert(0, ROOT)
    import bench
    x = bench.gen_numpy(n)[:, :ch] if ch <= 2 else np.tile(bench.gen_numpy(n), (1, (ch + 1) // 2))[:, :ch]
if bits == "32f":
    O.write_wav_f32(path, x, sr)
else:
    b = int(bits)
    q = np.clip(np.round(x.astype(np.float64) * 2 ** (b - 1)), -2 ** (b - 1), 2 ** (b - 1) - 1).astype(np.int32 if b > 16 else np.int16)
    O.write_wav_pcm(path, q, sr, b)
print(path, n, "frames", ch, "ch", bits)
