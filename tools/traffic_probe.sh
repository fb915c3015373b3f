#!/bin/bash
# DRAM traffic of the FFT-convolution launches under runtime switches (GPU box).  One CSV per setting
# in gpurun_out/traffic_<tag>_<label>.csv and a summary line each (bytes per step / algorithmic).
# Usage: tools/traffic_probe.sh <tag> "label|ENV=... ENV=..." ...
set -u
TAG=$1; shift
mkdir -p gpurun_out
for spec in "$@"; do
  IFS='|' read -r label envs <<< "$spec"
  env $envs ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:fftconv_kernel -c 200 --csv \
      --log-file gpurun_out/traffic_${TAG}_${label}.csv python tools/prof_sweep.py --seconds 3600 --steps 2 > gpurun_out/traffic_${TAG}_${label}.log 2>&1
  python - gpurun_out/traffic_${TAG}_${label}.csv "$label" <<'PY'
import csv, sys
rows = list(csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"')))
by = {}
for r in rows:
    v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"].lower()
    m = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}.get(u, 1)
    by.setdefault(int(r["ID"]), {})[r["Metric Name"]] = v * m
n = len(by)
rd = sum(x.get("dram__bytes_read.sum", 0) for x in by.values()); wr = sum(x.get("dram__bytes_write.sum", 0) for x in by.values())
t = sum(x.get("gpu__time_duration.sum", 0) for x in by.values())
steps = 2
print(f"{sys.argv[2]:28s} launches/step {n / steps:5.1f}  read {rd / steps / 1e6:8.1f} MB  write {wr / steps / 1e6:6.1f} MB  ratio {(rd + wr) / steps / 1382350848:5.3f}  kernel time/step {t / steps * 1e3:7.4f} ms (under ncu)")
PY
done
