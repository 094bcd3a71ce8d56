#!/usr/bin/env python3
"""Times dpt_op_attention alone (CUDA events, L2 flushed between launches). DPT_LIB selects the library.
usage: python tools/time_attention.py [tag] [B N heads head_dim dtype]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import muggled_dpt_b200._native as native  # noqa: E402

if os.environ.get("DPT_LIB"):
    native.LIB_PATH = os.path.join(ROOT, os.environ["DPT_LIB"])
import torch  # noqa: E402
from gpu_util import attention  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else ""
B, N, H, D = (int(x) for x in sys.argv[2:6]) if len(sys.argv) > 5 else (32, 1297, 16, 64)
dtype = torch.float16 if (len(sys.argv) > 6 and sys.argv[6] == "fp16") else torch.bfloat16
torch.manual_seed(0)
qkv = torch.randn(B, N, 3 * H * D, device="cuda").to(dtype)
bias = None
if D == 32:
    bias = torch.randn(1, H, N, N, device="cuda").to(dtype)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    attention(qkv, H, 0.125, bias=bias, head_dim=D)
ts = []
for _ in range(10):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    attention(qkv, H, 0.125, bias=bias, head_dim=D)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
ms = ts[len(ts) // 2]
fl = 4.0 * B * H * N * N * D
print(f"{tag} B={B} N={N} H={H} d={D} {dtype}: median {ms:.4f} ms  {fl / ms / 1e9:.1f} TFLOP/s (min {ts[0]:.4f})", flush=True)
