#!/usr/bin/env python3
"""Per-phase clock64 timeline of the softmax warps of the FIRST-GENERATION attention kernel (attn_tc.cuh, selected with
DPT_ATTN_V1=1; trace build of the library, -DATT_TRACE). The stamps are invasive (global stores per phase): use the
relative phase lengths, not the absolute step time. The shipped kernel (attn64_tc.cuh) is profiled with ncu instead.
usage (GPU box): python tools/attn_trace.py [B]     builds muggled_dpt_b200/lib/libdpt_b200_trace.so first"""
import ctypes
import os
import subprocess
import sys

os.environ["DPT_ATTN_V1"] = "1"

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
lib = os.path.join(ROOT, "muggled_dpt_b200", "lib", "libdpt_b200_trace.so")
csrc = os.path.join(ROOT, "muggled_dpt_b200", "csrc")
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-DATT_TRACE",
                       "-Xcompiler", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), "-I", csrc, "-o", lib,
                       os.path.join(csrc, "dpt_api.cu")])
import muggled_dpt_b200._native as native  # noqa: E402

native.LIB_PATH = lib
import torch  # noqa: E402
from gpu_util import attention  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
qkv = torch.randn(B, 1297, 3072, device="cuda").to(torch.bfloat16)
for _ in range(3):
    attention(qkv, 16, 0.125)
torch.cuda.synchronize()
handle = ctypes.CDLL(lib)
buf = (ctypes.c_longlong * (4 * 16 * 10))()
assert handle.dpt_debug_attn_trace(buf) == 0
names = ["wait s_full", "tmem ld S", "max+decide", "ffma sweep", "ex2 sweep", "sum+pack", "wait o_full", "tmem st P"]
for w in range(4):
    print(f"warp {w}: per-step phase durations in clocks (steps 1..9)")
    tot = [0] * 8
    n = 0
    for j in range(1, 10):
        t = [buf[(w * 16 + j) * 10 + ph] for ph in range(9)]
        d = [t[i + 1] - t[i] for i in range(8)]
        nxt = buf[(w * 16 + j + 1) * 10 + 0] - t[8]
        print(f"  j={j}: " + " ".join(f"{x:5d}" for x in d) + f" | step {t[8] - t[0]:5d} gap {nxt}")
        tot = [a + b for a, b in zip(tot, d)]
        n += 1
    print("  mean: " + ", ".join(f"{nm} {x / n:.0f}" for nm, x in zip(names, tot)) + f" | sum {sum(tot) / n:.0f}")
