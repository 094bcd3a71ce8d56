#!/usr/bin/env python3
"""Pins the parity gates of the -m gpu tests at <= 1.5 x the value measured on B200.
usage: python tools/make_parity_gates.py [gpurun_out/parity_gates.jsonl]   ->  tests/golden/parity_gates.json"""
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
log = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "parity_gates.jsonl")
out = os.path.join(ROOT, "tests", "golden", "parity_gates.json")
FACTOR = 1.5


def round_up(x, digits=2):
    if x <= 0:
        return 1e-12
    e = math.floor(math.log10(x)) - (digits - 1)
    return math.ceil(x / 10 ** e) * 10 ** e


worst = {}
for line in open(log):
    r = json.loads(line)
    worst[r["name"]] = max(worst.get(r["name"], 0.0), r["value"])
gates = json.load(open(out)) if os.path.exists(out) else {}
for name, v in sorted(worst.items()):
    gates[name] = {"measured": float(f"{v:.4g}"), "limit": float(f"{round_up(FACTOR * v):.4g}")}
json.dump(dict(sorted(gates.items())), open(out, "w"), indent=1)
print(f"{len(worst)} gates measured, {len(gates)} in {out}")
