#!/usr/bin/env python
"""Print the interesting parts of a bench.py JSON line: python scripts/show_bench.py <file>"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value=%.3e ms/step=%.4f frac=%.3f dram_frac=%s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"].get("dram_frac")))
for k in ("e2e", "e2e_prepacked", "e2e_wire", "scatter", "recompute_only", "cpu_baseline"):
    if d.get(k):
        print(k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in d[k].items() if a not in ("api", "sample")})
