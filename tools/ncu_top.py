#!/usr/bin/env python3
"""ncu_top.py <source-page csv> [n]: rank SASS instructions of one kernel by warp-stall samples"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    try:
        s = int(r[isamp])
    except Exception:
        continue
    data.append((s, r))
tot = sum(s for s, _ in data)
print("kernel:", rows[0][1], "| total samples", tot, "| instructions", len(data))
agg = {}
for s, r in data:
    for i, h in stall_cols:
        agg[h] = agg.get(h, 0) + int(r[i] or 0)
print("stall totals:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
data.sort(key=lambda x: -x[0])
for s, r in data[:n]:
    st = sorted([(int(r[i] or 0), h) for i, h in stall_cols], reverse=True)[:2]
    print(f"{s:7d} {100 * s / tot:5.1f}% ex={r[iex]:>8} {r[isrc][:84]:84s} {st}")
