#!/usr/bin/env python3
"""Per-kernel SASS opcode evidence for libdpt_b200.so (B200_PROFILING.md "What proves a Blackwell-native kernel"):
counts of UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG (TMA loads), UTCBAR-style commits, plus the legacy
tensor mnemonics that must NOT appear (HMMA = mma.sync, HGMMA = wgmma). Runs without a GPU.
usage: python tools/sass_histogram.py [lib.so] > profiles/r2_sass_opcode_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "muggled_dpt_b200", "lib", "libdpt_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
WATCH = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCATOMSWS", "MUFU.EX2", "FFMA2", "SYNCS", "HMMA", "HGMMA", "LDGSTS"]
kernels = collections.OrderedDict()
name = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        name = re.sub(r"\(dpt::\w+\)$", "", name).replace("void dpt::", "")
        kernels[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        op = m.group(1)
        kernels[name]["_total"] += 1
        for w in WATCH:
            if op == w or op.startswith(w + ".") or (w == "UTCHMMA.2CTA" and op.startswith("UTCHMMA") and ".2CTA" in op):
                kernels[name][w] += 1
tot = collections.Counter()
print(f"# SASS opcode counts per kernel of {os.path.relpath(lib, ROOT)} (cuobjdump -sass, sm_100a); {len(kernels)} kernels")
print("# UTCHMMA = tcgen05.mma kind::f16, LDTM/STTM = tcgen05.ld/st, UTMALDG / UTMASTG = cp.async.bulk.tensor loads / stores (TMA), LDGSTS = cp.async;")
print("# HMMA (mma.sync) / HGMMA (wgmma) would be legacy tensor paths: none expected")
cols = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "LDGSTS", "MUFU.EX2", "FFMA2", "HMMA", "HGMMA", "_total"]
print("kernel," + ",".join(cols))
for k, c in kernels.items():
    if c["UTCHMMA"] or c["UTMALDG"] or c["LDTM"]:
        print('"' + k + '",' + ",".join(str(c[x]) for x in cols))
    tot.update(c)
print('"ALL KERNELS (incl. the memory-bound helpers not listed above)",' + ",".join(str(tot[x]) for x in cols))
