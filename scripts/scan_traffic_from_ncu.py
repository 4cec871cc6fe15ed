#!/usr/bin/env python
"""profiles/scan_traffic.json from a committed ncu summary: dram__bytes_read.sum + dram__bytes_write.sum of the scan's
launches of ONE batch (k_scan sweep 1 + sweep 2 + k_merge), which is what bench.py reports as `roofline.traffic`
next to the algorithmic bytes of the same launches.  The capture is one `ncu --set full` pass of
`bench.py --steps 1 --warmup 0 --no-cpu --latency-ticks 0` (scripts/gpu_r2_measure6.sh), summarised into
profiles/<tag>_batch_full_summary.csv.

Usage: python scripts/scan_traffic_from_ncu.py [profiles/r2f_batch_full_summary.csv]"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2f_batch_full_summary.csv")
rows = {r[0]: r[1:] for r in csv.reader(open(src))}
names = rows["Kernel Name"][1:]
unit = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
total, parts = 0.0, {}
for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
    u = unit[rows[key][0]]
    for name, v in zip(names, rows[key][1:]):
        if name.startswith(("void k_scan", "k_scan", "k_merge")):
            total += float(v) * u
            parts[f"{name.split('(')[0]} {key}"] = float(v) * u
out = {"traffic_bytes_per_launch": total, "source": os.path.relpath(src, ROOT), "launches": "k_scan<0,1> + k_scan<0,2> + k_merge of one batch",
       "parts": parts}
with open(os.path.join(ROOT, "profiles", "scan_traffic.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))
