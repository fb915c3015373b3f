"""Per-kernel SASS opcode counts of a built library: python tools/sass_count.py lib.so [substring]
(static counts; proves FADD2/FMUL2/FFMA2, LDTM/STTM, UBLKPF etc. are in the binary and lets two builds be compared)."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else "fftconv"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, data = None, collections.OrderedDict()
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        data[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m and cur:
        data[cur][m.group(1)] += 1
ops = ["FADD2", "FMUL2", "FFMA2", "FADD", "FMUL", "FFMA", "FSEL", "MOV", "LDS", "STS", "LDG", "LDTM", "STTM", "SHFL", "BAR", "FMNMX3", "LOP3", "IADD3"]
for f, c in data.items():
    if want not in f:
        continue
    tot = sum(c.values())
    print(f, "total", tot, " ".join(f"{o}={c[o]}" for o in ops if c[o]))
