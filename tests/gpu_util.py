"""helpers for the -m gpu tests: call the C ABI operators on torch CUDA tensors"""
import ctypes as C
import json
import os

import torch

from muggled_dpt_b200 import _native as N

DT = {torch.float16: N.DPT_F16, torch.bfloat16: N.DPT_BF16}


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def conv_gemm(A, Wt, bias=None, add1=None, add2=None, want_relu=False, taps=1, xoff=0, act=0, out_f32=False, N_out=None):
    """A [B,H,W,C] 16-bit, Wt [N, taps*kpad] 16-bit -> out [B,H,W-xoff,N]"""
    B, H, W, Cc = A.shape
    n = N_out if N_out is not None else Wt.shape[0]
    n_cols = n // 2 if act == 4 else n  # act 4 = SwiGLU gate in the epilogue: half as many output columns
    out = torch.empty((B, H, W - xoff, n_cols), device=A.device, dtype=torch.float32 if out_f32 else A.dtype)
    out_relu = torch.empty_like(out) if want_relu else None
    rc = N.lib().dpt_op_conv_gemm(_p(A), _p(Wt), _p(bias), _p(out), _p(add1), _p(add2), _p(out_relu), B, H, W, Cc, n,
                                  taps, xoff, act, int(out_f32), DT[A.dtype], _stream())
    N.check(rc, None, "dpt_op_conv_gemm")
    torch.cuda.synchronize()
    return (out, out_relu) if want_relu else out


def attention(qkv, heads, scale, bias=None, head_dim=64, bias_wmod=1):
    """bias: [heads, n, n] or [bias_wmod, heads, n, n]"""
    B, n, F3 = qkv.shape
    out = torch.empty((B, n, F3 // 3), device=qkv.device, dtype=qkv.dtype)
    ld = 0
    if bias is not None:  # pad rows to a multiple of 128 columns (kernel contract)
        ld = (n + 127) // 128 * 128
        padded = torch.zeros(tuple(bias.shape[:-1]) + (ld,), device=bias.device, dtype=bias.dtype)
        padded[..., :n] = bias
        bias = padded
    rc = N.lib().dpt_op_attention(_p(qkv), _p(bias), ld, bias_wmod, _p(out), B, n, heads, head_dim, float(scale),
                                  DT[qkv.dtype], _stream())
    N.check(rc, None, "dpt_op_attention")
    torch.cuda.synchronize()
    return out


def layernorm(x, w, b, eps, dtype):
    M, F = x.shape
    y = torch.empty((M, F), device=x.device, dtype=dtype)
    rc = N.lib().dpt_op_layernorm(_p(x), _p(w), _p(b), _p(y), M, F, float(eps), DT[dtype], _stream())
    N.check(rc, None, "dpt_op_layernorm")
    torch.cuda.synchronize()
    return y


def resize(x, OH, OW):
    B, IH, IW, Cc = x.shape
    y = torch.empty((B, OH, OW, Cc), device=x.device, dtype=x.dtype)
    rc = N.lib().dpt_op_resize_bilinear(_p(x), _p(y), B, IH, IW, OH, OW, Cc, DT[x.dtype], _stream())
    N.check(rc, None, "dpt_op_resize_bilinear")
    torch.cuda.synchronize()
    return y


def rel_err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item(), (a - b).abs().max().item()


# ---------------------------------------------------------------------------------------------------------------------
# parity gates: every tolerance of the -m gpu parity tests lives in tests/golden/parity_gates.json as
# {name: {"measured": value on B200, "limit": <= 1.5 x measured}} (tools/make_parity_gates.py writes it from the log a
# GPU run leaves in gpurun_out/parity_gates.jsonl). A name without an entry falls back to the loose default passed by
# the test and is reported, so a new test can be measured once before its gate is pinned.
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_GATES_PATH = os.path.join(_ROOT, "tests", "golden", "parity_gates.json")
_GATES = json.load(open(_GATES_PATH)) if os.path.exists(_GATES_PATH) else {}


def gate(name: str, value: float, default_limit: float) -> float:
    """assert value < gate(name); logs (name, value, limit) for tools/make_parity_gates.py"""
    entry = _GATES.get(name)
    limit = entry["limit"] if entry else default_limit
    try:
        os.makedirs(os.path.join(_ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(_ROOT, "gpurun_out", "parity_gates.jsonl"), "a") as f:
            f.write(json.dumps({"name": name, "value": value, "limit": limit, "pinned": entry is not None}) + "\n")
    except OSError:
        pass
    print(f"gate {name}: {value:.3e} (limit {limit:.3e}{'' if entry else ', UNPINNED default'})")
    assert value < limit, f"{name}: {value:.3e} is not below its gate {limit:.3e}"
    return value
