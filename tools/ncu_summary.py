"""Summarise an .ncu-rep (one kernel) into text: key raw metrics, opcode mix, stall
samples per phase between barriers.  Usage: python tools/ncu_summary.py rep.ncu-rep > profiles/rNN/x.txt"""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(raw))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"]
for k, vals in enumerate(rows[2:]):
    print(f"## launch {k}")
    for h, u, v in zip(hdr, units, vals):
        if h in want:
            print(f"{h:88s} {u:16s} {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
start = [i for i, l in enumerate(src) if l.startswith('"Address"')]
if start:
    r = list(csv.DictReader(src[start[0]:]))
    tot = sum(int(x["Instructions Executed"]) for x in r)
    ts = max(1, sum(int(x["# Samples"]) for x in r))
    byop, samp = defaultdict(int), defaultdict(int)
    for x in r:
        t = x["Source"].strip().split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        byop[op] += int(x["Instructions Executed"])
        samp[op] += int(x["# Samples"])
    print(f"\n## opcode mix (warp instructions executed, total {tot}; SASS lines {len(r)})")
    for k, v in sorted(byop.items(), key=lambda kv: -kv[1])[:20]:
        print(f"{k:10s} {v:12d} {100 * v / tot:5.1f}%   stall samples {100 * samp[k] / ts:5.1f}%")
    print("\n## phases between BAR.SYNC: inst%  fp-inst  lds/sts  ldg | samples%  long_sb short_sb wait not_selected barrier")
    cur = defaultdict(int)
    reg = []
    for x in r:
        t = x["Source"].strip().split()
        op = t[1] if t[0].startswith("@") else t[0]
        n = int(x["Instructions Executed"])
        cur["inst"] += n
        cur["samp"] += int(x["# Samples"])
        for k in ("stall_long_sb", "stall_short_sb", "stall_wait", "stall_not_selected", "stall_barrier"):
            cur[k] += int(x[k])
        if op.split(".")[0] in ("FADD", "FMUL", "FFMA"):
            cur["fp"] += n
        if op.startswith("LDS") or op.startswith("STS"):
            cur["lds"] += n
        if op.startswith("LDG"):
            cur["ldg"] += n
        if op.startswith("BAR"):
            reg.append(cur)
            cur = defaultdict(int)
    reg.append(cur)
    for i, x in enumerate(reg):
        print(f"phase {i}: {100 * x['inst'] / tot:5.1f}% {x['fp']:9d} {x['lds']:8d} {x['ldg']:7d} | {100 * x['samp'] / ts:5.1f}%  "
              f"{x['stall_long_sb']:5d} {x['stall_short_sb']:5d} {x['stall_wait']:5d} {x['stall_not_selected']:5d} {x['stall_barrier']:5d}")
