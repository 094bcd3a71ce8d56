#!/usr/bin/env python3
"""Runs single operators at ViT-L shapes once each (for `ncu --set full`). usage: prof_ops.py [gemm|attn|all] [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from gpu_util import attention, conv_gemm  # noqa: E402
from muggled_dpt_b200.weights import pack_linear  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
N = 1297
M = B * N
dt = torch.bfloat16
torch.manual_seed(0)
if what in ("gemm", "all"):
    A = torch.randn(1, 1, M, 1024, device="cuda").to(dt)
    Wq = pack_linear((torch.randn(3072, 1024, device="cuda") / 32).to(dt))
    bq = torch.randn(3072, device="cuda")
    for _ in range(2):
        conv_gemm(A, Wq, bq)                 # qkv shape, 16-bit out
    W1 = pack_linear((torch.randn(4096, 1024, device="cuda") / 32).to(dt))
    b1 = torch.randn(4096, device="cuda")
    for _ in range(2):
        conv_gemm(A, W1, b1, act=1)          # fc1 shape, GELU
    Wp = pack_linear((torch.randn(1024, 1024, device="cuda") / 32).to(dt))
    bp = torch.randn(1024, device="cuda")
    x = torch.randn(1, 1, M, 1024, device="cuda")
    for _ in range(2):
        conv_gemm(A, Wp, bp, add1=x, out_f32=True)  # proj shape, fp32 residual
if what in ("attn", "all"):
    qkv = torch.randn(B, N, 3072, device="cuda").to(dt)
    for _ in range(2):
        attention(qkv, 16, 0.125)
torch.cuda.synchronize()
print("done")
