"""profiles/ncu_traffic.json from the DRAM-traffic pass of tools/ncu_round.sh:
python tools/traffic_json.py profiles/r01/ncu_traffic_vNN.csv <launches_per_step>"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path, per_step = sys.argv[1], int(sys.argv[2])
rows = list(csv.DictReader(l for l in open(path) if l.startswith('"')))
by_id = {}
for r in rows:
    if "fftconv_kernel<0" not in r["Kernel Name"].replace("(int)", ""):
        continue
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"].lower()
    mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    by_id.setdefault(r["ID"], {})[r["Metric Name"]] = v * mult
ids = sorted(by_id, key=int)
ids = ids[:len(ids) - len(ids) % per_step]  # whole steps only
tot = sum(by_id[i].get("dram__bytes_read.sum", 0) + by_id[i].get("dram__bytes_write.sum", 0) for i in ids)
out = {
    "fftconv_filter_dram_bytes_per_launch": tot / len(ids),
    "fftconv_filter_dram_bytes_per_step": tot / (len(ids) // per_step),
    "launches_per_step": per_step,
    "launches_captured": len(ids),
    "source": f"{os.path.relpath(path, ROOT)}: dram__bytes_read.sum + dram__bytes_write.sum of the fftconv_kernel<EPI_POINTS> launches of "
              f"{len(ids) // per_step} whole sweep step(s) ({per_step} launches per step: bootstrap wave, then 8 and up to 128 segments per CTA), "
              "1 h stereo 48 kHz; algorithmic bytes on the same basis = 1382350848 per step",
}
json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
