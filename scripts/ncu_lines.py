#!/usr/bin/env python
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump by CUDA source line."""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
cur_file = None
hdr = None
data = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or r[0] == "Function Name":
        continue
    if r[2] != "-":       # SASS row
        continue
    try:
        smp = float(r[hdr.index("# Samples")] or 0)
        ins = float(r[hdr.index("Instructions Executed")] or 0)
    except ValueError:
        continue
    stalls = {h: float(v or 0) for h, v in zip(hdr, r) if h.startswith("stall_") and "Not Issued" not in h}
    data.append((smp, ins, cur_file, r[0], r[1].strip()[:100], stalls))
ts = sum(d[0] for d in data) or 1
ti = sum(d[1] for d in data) or 1
print(f"total samples {ts:.0f} instructions {ti:.0f}")
agg = {}
for d in data:
    for k, v in d[5].items():
        agg[k] = agg.get(k, 0) + v
print("stall mix:", {k: f"{100*v/ts:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for d in sorted(data, key=lambda d: -d[0])[:top]:
    st = sorted(d[5].items(), key=lambda kv: -kv[1])[:2]
    print(f"{100*d[0]/ts:5.1f}% smp {100*d[1]/ti:5.1f}% ins {d[2]}:{d[3]:>4s} {d[4]}  [{', '.join(f'{k[6:]} {v:.0f}' for k, v in st)}]")
