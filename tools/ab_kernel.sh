#!/bin/bash
# A/B of library builds and runtime switches on the headline workload (GPU box, through gpurun).
# Usage: tools/ab_kernel.sh <tag> "<NAME=path-or-empty ENV=... >" ...   each argument: label|lib path|env assignments
# Output: gpurun_out/ab_<tag>.jsonl (one bench line per variant, key "variant" added)
set -u
TAG=$1; shift
mkdir -p gpurun_out
: > gpurun_out/ab_${TAG}.jsonl
for spec in "$@"; do
  IFS='|' read -r label lib envs <<< "$spec"
  line=$(env PHASEROT_LIB="$lib" $envs python bench.py --steps 10 --warmup 3 --no-cpu --no-extra 2>gpurun_out/ab_${TAG}_${label}.err | tail -1)
  python - "$label" "$line" >> gpurun_out/ab_${TAG}.jsonl <<'PY'
import json, sys
try:
    d = json.loads(sys.argv[2])
    r = d["roofline"]
    print(json.dumps({"variant": sys.argv[1], "ms_per_step": d["ms_per_step"], "kernel_ms": r["kernel_ms_per_step"], "frac": r["frac"],
                      "all_kernels_ms": r["all_kernels_ms"], "e2e_ms": d["e2e"]["ms_per_step"], "surv": d["pruning"]["survivor_fraction"], "clocks": d["clocks"]}))
except Exception as ex:
    print(json.dumps({"variant": sys.argv[1], "error": str(ex), "raw": sys.argv[2][:300]}))
PY
done
cat gpurun_out/ab_${TAG}.jsonl
