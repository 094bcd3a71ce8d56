#!/usr/bin/env python3
"""Summarises tools/sweep_gemm_modes.sh: mean time per GEMM label for each tile mode and batch size."""
import collections
import csv
import glob
import os
import re
import sys

root = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/sweep"
table = collections.defaultdict(dict)
for path in sorted(glob.glob(os.path.join(root, "prof_*_m*_b*.csv"))):
    m = re.search(r"prof_(\w+?)_m(\d)_b(\d+)\.csv", os.path.basename(path))
    model, mode, B = m.group(1), int(m.group(2)), int(m.group(3))
    acc = collections.defaultdict(list)
    with open(path) as f:
        for row in csv.reader(f):
            if not row or not row[0].startswith("gemm"):
                continue
            kind, label = row[0].split(":", 1)
            label = re.sub(r"blk\d+\.", "blk.", label)
            acc[label].append((float(row[1]), kind))
    for label, v in acc.items():
        table[(model, B, label)][mode] = (sum(t for t, _ in v), len(v), v[0][1])
print("model,B,label,launches," + ",".join(f"ms_mode{m},kind{m}" for m in range(4)) + ",best")
for (model, B, label), modes in sorted(table.items()):
    cells = []
    best = min((modes[m][0], m) for m in modes if m != 0)[1] if any(m != 0 for m in modes) else -1
    for m in range(4):
        cells += [f"{modes[m][0]:.4f}", modes[m][2]] if m in modes else ["", ""]
    n = next(iter(modes.values()))[1]
    print(f"{model},{B},{label},{n}," + ",".join(cells) + f",{best}")
