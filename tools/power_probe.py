#!/usr/bin/env python3
"""Sustained per-operator rate, SM clock and board power: each operator is launched back to back for ~2.5 s while
NVML is sampled every 20 ms. Tells which kernels run power-capped (and at what energy per launch) versus
time-bound at full clock. usage: power_probe.py [B]"""
import ctypes as C
import os
import statistics
import sys
import threading
import time

import pynvml
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from muggled_dpt_b200 import _native as N  # noqa: E402
from muggled_dpt_b200.weights import pack_linear  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
SECS = float(os.environ.get("PROBE_SECS", "2.5"))
Ntok, F = 1297, 1024
M = B * Ntok
dt = torch.bfloat16
L = N.lib()
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)


def p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.stop = False
        self.clk, self.pw = [], []

    def run(self):
        while not self.stop:
            self.clk.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.pw.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
            time.sleep(0.02)


OPS = os.environ.get("PROBE_OPS", "")


def probe(name, fn, flop=0.0, nbytes=0.0):
    if OPS and not any(o in name for o in OPS.split(",")):
        return
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    s = Sampler()
    s.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    t0 = time.time()
    e0.record()
    while time.time() - t0 < SECS:
        for _ in range(20):
            fn()
        n += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    s.stop = True
    s.join()
    ms = e0.elapsed_time(e1) / n
    k = len(s.clk) // 3  # skip the ramp
    clk, pw = statistics.median(s.clk[k:]), statistics.mean(s.pw[k:])
    print(f"{name:28s} {ms:8.4f} ms  {flop / ms / 1e9:8.1f} TF/s {nbytes / ms / 1e6:8.0f} GB/s  clk {clk:6.0f} MHz  {pw:6.0f} W  {pw * ms / 1000:8.3f} J/launch",
          flush=True)


torch.manual_seed(0)
A = torch.randn(1, 1, M, F, device="cuda").to(dt)
A4 = torch.randn(1, 1, M, 4 * F, device="cuda").to(dt)
x32 = torch.randn(1, 1, M, F, device="cuda")


def gemm(a, n, k, act=0, f32=False, add=None):
    W = pack_linear((torch.randn(n, k, device="cuda") / 32).to(dt))
    b = torch.randn(n, device="cuda")
    out = torch.empty((1, 1, M, n), device="cuda", dtype=torch.float32 if f32 else dt)

    def fn():
        rc = L.dpt_op_conv_gemm(p(a), p(W), p(b), p(out), p(add if add is not None else None), None, None, 1, 1, M, k, n, 1, 0, act,
                                int(f32), N.DPT_BF16, st())
        assert rc == 0, L.dpt_op_last_error()
    return fn, 2.0 * M * n * k


fn, fl = gemm(A, 3 * F, F)
probe("gemm qkv", fn, fl)
fn, fl = gemm(A, 4 * F, F, act=1)
probe("gemm fc1+gelu", fn, fl)
fn, fl = gemm(A4, F, 4 * F, f32=True, add=x32)
probe("gemm fc2+residual", fn, fl)
fn, fl = gemm(A, F, F, f32=True, add=x32)
probe("gemm proj+residual", fn, fl)

qkv = torch.randn(B, Ntok, 3 * F, device="cuda").to(dt)
ao = torch.empty((B, Ntok, F), device="cuda", dtype=dt)


def attn():
    rc = L.dpt_op_attention(p(qkv), None, 0, 1, p(ao), B, Ntok, 16, 64, 0.125, N.DPT_BF16, st())
    assert rc == 0


probe("attention", attn, 4.0 * B * 16 * Ntok * Ntok * 64)

w = torch.ones(F, device="cuda")
bb = torch.zeros(F, device="cuda")
y = torch.empty((M, F), device="cuda", dtype=dt)
x2 = x32.view(M, F)


def ln():
    rc = L.dpt_op_layernorm(p(x2), p(w), p(bb), p(y), M, F, 1e-6, N.DPT_BF16, st())
    assert rc == 0


probe("layernorm", ln, 0.0, M * F * 6.0)

# cuBLAS reference points at the same shapes
Wt = torch.randn(4 * F, F, device="cuda").to(dt)
A2 = A.view(M, F)
probe("torch.matmul fc1 shape", lambda: torch.matmul(A2, Wt.t()), 2.0 * M * 4 * F * F)
a8 = torch.randn(8192, 8192, device="cuda").to(dt)
b8 = torch.randn(8192, 8192, device="cuda").to(dt)
probe("torch.matmul 8192^3", lambda: torch.matmul(a8, b8), 2.0 * 8192 ** 3)
