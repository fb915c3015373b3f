for spec in "default|" "steps6|PHASEROT_WALK_STEPS=6" "steps24|PHASEROT_WALK_STEPS=24" "steps3|PHASEROT_WALK_STEPS=3" "boot32k|PHASEROT_BOOT_BRUTE=32768" "boot512k|PHASEROT_BOOT_BRUTE=524288" "rad4|PHASEROT_WALK_RAD=4e-3"; do
  IFS='|' read -r label envs <<< "$spec"
  for m in sine two_sine; do
    echo -n "$label $m: "; env $envs timeout -k 5 120 python tools/dense_probe.py 3600 $m 2>&1 | grep "^{'ms_per_step" | python -c "
import sys, ast
for l in sys.stdin:
    d=ast.literal_eval(l); print(round(d['ms_per_step'],3), d['kernels_ms'])
"
  done
done
