import json,sys
for f in sys.argv[1:]:
    try:
        d=json.load(open(f)); print(f, d["ms_per_step"], d["roofline"]["all_kernels_ms"], round(d["roofline"]["frac"],4), d["e2e"]["ms_per_step"])
    except Exception as e: print(f,"ERR",e)
