"""Packed fp32 warp instructions (FADD2 + FMUL2 + FFMA2) one FFT-convolution segment executes, from an
`ncu --set full --import-source on` capture of ONE fftconv_kernel launch; written into profiles/ncu_traffic.json
for bench.py's roofline.fp32_issue_frac.
  python tools/fp2_per_segment.py gpurun_out/ncu_<tag>_conv.ncu-rep <segments in the captured launch>
(1 h stereo 48 kHz, third launch of a step: 2 x (7032 - 592) = 12880 segments)"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, segs = sys.argv[1], int(sys.argv[2])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
start = [i for i, l in enumerate(src) if l.startswith('"Address"')][0]
fp2 = tot = 0
for x in csv.DictReader(src[start:]):
    t = x["Source"].strip().split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    n = int(x["Instructions Executed"])
    tot += n
    if op in ("FADD2", "FMUL2", "FFMA2"):
        fp2 += n
path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
d = json.load(open(path)) if os.path.exists(path) else {}
d["fftconv_fp2_warp_inst_per_segment"] = fp2 / segs
d["fftconv_warp_inst_per_segment"] = tot / segs
d["fp2_source"] = f"{os.path.basename(rep)}: {fp2} packed fp32 of {tot} warp instructions over {segs} segments"
json.dump(d, open(path, "w"), indent=1)
print(json.dumps(d, indent=1))
