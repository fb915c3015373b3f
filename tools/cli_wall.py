"""Wall clock of the phase-rotate CLI against the reference build (oracle/_ref/phase-rotate: the reference's own
sources with the stand-in FFT / libsndfile), on the GPU box.  Output: gpurun_out/cli_wall.txt"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
OURS = "phaserotate/lv2_b200/bin/phase-rotate"
REF = "oracle/_ref/phase-rotate"
TMP = "/tmp/cliw"
os.makedirs(TMP, exist_ok=True)
os.makedirs("gpurun_out", exist_ok=True)
lines = []


def say(s):
    print(s, flush=True)
    lines.append(s)


def wall(label, cmd, runs=3, env=None, limit=900):
    for k in range(runs):
        t0 = time.perf_counter()
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=limit, env=env)
            out, rc = (r.stdout + r.stderr).strip().splitlines(), r.returncode
        except subprocess.TimeoutExpired:
            out, rc = ["timeout"], -1
        dt = time.perf_counter() - t0
        say(f"{label:24s} run {k + 1}: {dt:7.3f} s rc={rc} | {(out[-1] if out else '')[:110]}")
    return out


for args in (["c1.wav", "30", "two_sine", "32f"], ["h16.wav", "3600", "programme", "16"]):
    subprocess.run([sys.executable, "tools/make_wav.py", os.path.join(TMP, args[0])] + args[1:], check=True, capture_output=True)
c1, h16 = os.path.join(TMP, "c1.wav"), os.path.join(TMP, "h16.wav")
tenv = dict(os.environ, PHASEROT_CLI_TIMING="1")
say("== config 1 (30 s stereo float WAV), -s 1; nothing else holds the GPU (no persistence daemon on the box: every start re-initialises the device)")
wall("cuda -s 1, cold device", [OURS, "-s", "1", c1])
# what nvidia-persistenced does on a production host: keep the device initialised between processes
holder = subprocess.Popen([sys.executable, "-c", "import torch,time,sys; torch.zeros(1, device='cuda'); print('ready', flush=True); time.sleep(1200)"], stdout=subprocess.PIPE, text=True)
holder.stdout.readline()
say("== from here on a second process holds a CUDA context open (the state a persistence daemon keeps)")
wall("cuda -s 1", [OURS, "-s", "1", c1])
if os.access(REF, os.X_OK):
    wall("reference -s 1", [REF, "-s", "1", c1])
say("== config 1, stage timing of the CUDA CLI (stderr)")
r = subprocess.run([OURS, "-s", "1", c1], capture_output=True, text=True, env=tenv)
for ln in r.stderr.strip().splitlines()[:14]:
    say("   " + ln[:140])
say("== config 1, --subsample 10 (0.1 degree grid; the reference has no such grid)")
wall("cuda --subsample 10", [OURS, "--subsample", "10", "-s", "1", c1])
say("== 1 h stereo 16-bit WAV (691 MB), -s 1")
wall("cuda 1 h 16-bit", [OURS, "-s", "1", h16])
r = subprocess.run([OURS, "-s", "1", h16], capture_output=True, text=True, env=tenv)
for ln in r.stderr.strip().splitlines()[:14]:
    say("   " + ln[:140])
if os.access(REF, os.X_OK) and "--ref-1h" in sys.argv:
    wall("reference 1 h 16-bit", [REF, "-s", "1", h16], runs=1)
holder.kill()
open("gpurun_out/cli_wall.txt", "w").write("\n".join(lines) + "\n")
