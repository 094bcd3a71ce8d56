#!/usr/bin/env python3
"""Clock64 timeline of the spatial GEMM's epilogue warps and MMA thread (trace build of the library, -DGEMM_TRACE): per
tile, when the epilogue got its operands, how long it waited for the accumulator, how long each 32-column unit took
(TMEM load wait | bias / LayerNorm fold / activation / pack / staging | coalesced global IO), and how long the MMA thread
waited for a free accumulator stage. The stamps are global stores by one lane per warp - cheap, but not free: read the
phase lengths relative to each other.
usage (GPU box): python tools/gemm_trace.py [M K N act [f32]]     (re)builds muggled_dpt_b200/lib/libdpt_b200_trace.so when stale"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
lib = os.path.join(ROOT, "muggled_dpt_b200", "lib", os.environ.get("GEMM_TRACE_LIB", "libdpt_b200_trace.so"))
csrc = os.path.join(ROOT, "muggled_dpt_b200", "csrc")
if not os.path.exists(lib) or (not os.environ.get("GEMM_TRACE_NOBUILD") and os.path.getmtime(lib) < max(os.path.getmtime(os.path.join(csrc, f)) for f in os.listdir(csrc))):
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-DGEMM_TRACE",
                           "-Xcompiler", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), "-I", csrc, "-o", lib,
                           os.path.join(csrc, "dpt_api.cu")])
import muggled_dpt_b200._native as native  # noqa: E402

native.LIB_PATH = lib
import torch  # noqa: E402
from gpu_util import conv_gemm  # noqa: E402
from muggled_dpt_b200.weights import pack_linear  # noqa: E402

M, K, N, act = (int(v) for v in sys.argv[1:5]) if len(sys.argv) >= 5 else (10376, 768, 3072, 1)
f32_residual = len(sys.argv) >= 6 and sys.argv[5] == "f32"  # fp32 output with an in-place fp32 residual (proj / fc2)
A = torch.randn(1, 1, M, K, device="cuda").to(torch.bfloat16)
W = pack_linear((torch.randn(N, K) * K**-0.5).to(torch.bfloat16)).cuda()
bias = torch.randn(N, device="cuda")
x = torch.randn(M, N, device="cuda")
for _ in range(3):
    if f32_residual:
        rc = native.lib().dpt_op_conv_gemm(A.data_ptr(), W.data_ptr(), bias.data_ptr(), x.data_ptr(), x.data_ptr(), None, None,
                                           1, 1, M, K, N, 1, 0, 0, 1, native.DPT_BF16,
                                           ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        native.check(rc, None, "gemm")
    else:
        conv_gemm(A, W, bias, act=act)
torch.cuda.synchronize()
handle = ctypes.CDLL(lib)
buf = (ctypes.c_longlong * (9 * 8 * 16))()
assert handle.dpt_debug_gemm_trace(buf) == 0


def g(role, it, k):
    return buf[(role * 8 + it) * 16 + k]


print(f"GEMM M={M} K={K} N={N} act={act}{' fp32 out + in-place residual' if f32_residual else ''}: CTA 4, clocks")
t00 = g(8, 0, 0)
print("MMA thread: per tile  [wait acc stage -> first operands ready -> all MMAs issued]")
for it in range(8):
    if g(8, it, 2) == 0:
        break
    nxt = g(8, it + 1, 0) - g(8, it, 2) if it + 1 < 8 and g(8, it + 1, 0) else 0
    print(f"  tile {it}: start {g(8, it, 0) - t00:7d}  operands +{g(8, it, 1) - g(8, it, 0):5d}  issue {g(8, it, 2) - g(8, it, 1):6d}"
          f"  wait for next stage {nxt:6d}")
for w in (0, 5):
    print(f"epilogue warp {w}: per tile  [operands+barriers | wait accumulator | units: ld-wait / math+staging / global IO]")
    for it in range(8):
        if g(w, it, 15) == 0:
            break
        s = [g(w, it, k) for k in range(16)]
        units = []
        prev = s[2]
        for u in range(4):
            a, b, c = s[3 + 3 * u], s[4 + 3 * u], s[5 + 3 * u]
            if a == 0:
                break
            seg = f"{a - prev}/{b - a}"
            prev = b
            if c >= b and c != 0:
                seg += f"/{c - b}"
                prev = c
            units.append(seg)
        print(f"  tile {it}: start {s[0] - t00:7d}  prologue {s[1] - s[0]:5d}  wait acc {s[2] - s[1]:6d}  " + "  ".join(units) +
              f"  | tile {s[15] - s[0]:6d}")
