"""Kernel-level parity (-m gpu): each sm_100a kernel, called through the C ABI, against a plain PyTorch fp32
reference of the same op on the same (16-bit-rounded) inputs. Tolerances are stated per test: the kernels accumulate
in fp32, so the only differences are the final 16-bit rounding (bf16: 2^-9 = 2e-3 relative, fp16: 2^-11 = 5e-4) and
summation order."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from muggled_dpt_b200.weights import pack_conv, pack_linear  # noqa: E402


def _tol(dtype):
    return 6e-3 if dtype == torch.bfloat16 else 1.5e-3


def _mk(shape, dtype, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).cuda()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M,K,N", [(128, 64, 32), (256, 128, 64), (1000, 384, 1152), (300, 1024, 256), (77, 192, 48),
                                   (4099, 256, 1024)])
def test_gemm_plain(M, K, N, dtype):
    from gpu_util import conv_gemm, rel_err

    A = _mk((1, 1, M, K), dtype, 1)
    W = _mk((N, K), dtype, 2, K**-0.5)
    bias = _mk((N,), torch.float32, 3)
    out = conv_gemm(A, pack_linear(W), bias)
    ref = A.float().reshape(M, K) @ W.float().t() + bias
    rel, mx = rel_err(out.reshape(M, N), ref)
    assert rel < _tol(dtype), (rel, mx)
    assert torch.isfinite(out.float()).all()


@pytest.mark.parametrize("act", [1, 2])
def test_gemm_activation(act):
    from gpu_util import conv_gemm, rel_err

    M, K, N = 513, 256, 512
    A = _mk((1, 1, M, K), torch.bfloat16, 4)
    W = _mk((N, K), torch.bfloat16, 5, K**-0.5 * 2)
    bias = _mk((N,), torch.float32, 6)
    out = conv_gemm(A, pack_linear(W), bias, act=act)
    pre = A.float().reshape(M, K) @ W.float().t() + bias
    ref = F.gelu(pre) if act == 1 else F.relu(pre)
    rel, mx = rel_err(out.reshape(M, N), ref)
    assert rel < 6e-3, (rel, mx)


def test_gemm_f32_residual_inplace():
    from gpu_util import rel_err
    import ctypes as C
    from muggled_dpt_b200 import _native as NN

    M, K, N = 700, 512, 384
    A = _mk((1, 1, M, K), torch.bfloat16, 7)
    W = _mk((N, K), torch.bfloat16, 8, K**-0.5)
    bias = _mk((N,), torch.float32, 9)
    x = _mk((M, N), torch.float32, 10)
    ref = x + A.float().reshape(M, K) @ W.float().t() + bias
    Wp = pack_linear(W)
    rc = NN.lib().dpt_op_conv_gemm(A.data_ptr(), Wp.data_ptr(), bias.data_ptr(), x.data_ptr(), x.data_ptr(), None, None,
                                   1, 1, M, K, N, 1, 0, 0, 1, NN.DPT_BF16,
                                   C.c_void_p(torch.cuda.current_stream().cuda_stream))
    NN.check(rc, None, "gemm")
    torch.cuda.synchronize()
    rel, mx = rel_err(x, ref)
    assert rel < 1e-3, (rel, mx)  # fp32 output: only bf16 input rounding (already in ref) + accumulation order


def test_gemm_token_mode_skips_cls():
    from gpu_util import conv_gemm, rel_err

    B, Ntok, K, N = 3, 1 + 36, 128, 96
    A = _mk((B, 1, Ntok, K), torch.bfloat16, 11)
    W = _mk((N, K), torch.bfloat16, 12, K**-0.5)
    bias = _mk((N,), torch.float32, 13)
    out = conv_gemm(A, pack_linear(W), bias, xoff=1)
    ref = A.float()[:, 0, 1:, :] @ W.float().t() + bias
    rel, mx = rel_err(out.reshape(B, Ntok - 1, N), ref)
    assert rel < 6e-3, (rel, mx)


@pytest.mark.parametrize("B,H,W,C,N", [(2, 18, 18, 64, 64), (1, 36, 36, 32, 64), (2, 20, 12, 96, 32), (1, 72, 72, 256, 256),
                                       (1, 9, 144, 48, 128)])
def test_conv3x3(B, H, W, C, N):
    from gpu_util import conv_gemm, rel_err

    x = _mk((B, H, W, C), torch.bfloat16, 14)
    w = _mk((N, C, 3, 3), torch.bfloat16, 15, (9 * C) ** -0.5)
    bias = _mk((N,), torch.float32, 16)
    out = conv_gemm(x, pack_conv(w), bias, taps=9)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, padding=1).permute(0, 2, 3, 1)
    rel, mx = rel_err(out, ref)
    assert rel < 6e-3, (rel, mx)


def test_conv3x3_addends_and_relu_copy():
    from gpu_util import conv_gemm, rel_err

    B, H, W, C = 2, 36, 36, 64
    x = _mk((B, H, W, C), torch.bfloat16, 17)
    w = _mk((C, C, 3, 3), torch.bfloat16, 18, (9 * C) ** -0.5)
    bias = _mk((C,), torch.float32, 19)
    a1 = _mk((B, H, W, C), torch.bfloat16, 20)
    a2 = _mk((B, H, W, C), torch.bfloat16, 21)
    out, out_relu = conv_gemm(x, pack_conv(w), bias, add1=a1, add2=a2, want_relu=True, taps=9)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, padding=1).permute(0, 2, 3, 1) + a1.float() + a2.float()
    rel, mx = rel_err(out, ref)
    assert rel < 8e-3, (rel, mx)
    assert torch.equal(out_relu, torch.relu(out))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("B,N,heads", [(1, 128, 1), (2, 257, 2), (1, 1297, 3), (2, 100, 1), (1, 17, 1), (2, 64, 2),
                                       (3, 65, 1), (1, 193, 2)])
def test_attention(B, N, heads, dtype):
    from gpu_util import attention, rel_err

    Fd = heads * 64
    qkv = _mk((B, N, 3 * Fd), dtype, 22)
    out = attention(qkv, heads, 0.125)
    q, k, v = qkv.float().reshape(B, N, 3, heads, 64).permute(2, 0, 3, 1, 4).unbind(0)
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, N, Fd)
    rel, mx = rel_err(out, ref)
    assert rel < (1e-2 if dtype == torch.bfloat16 else 3e-3), (rel, mx)  # P is rounded to 16 bits before P@V


@pytest.mark.parametrize("N", [40, 333, 1297])
def test_attention_large_logits(N):
    """logits with a spread of tens of exp2 units: the stabiliser moves (lazy rescale) and the two split-KV halves of
    a row end up with very different maxima before they are merged"""
    from gpu_util import attention, rel_err

    B, heads = 2, 2
    Fd = heads * 64
    qkv = (_mk((B, N, 3 * Fd), torch.bfloat16, 31).float() * 3.0).to(torch.bfloat16)
    out = attention(qkv, heads, 0.125)
    q, k, v = qkv.float().reshape(B, N, 3, heads, 64).permute(2, 0, 3, 1, 4).unbind(0)
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, N, Fd)
    rel, mx = rel_err(out, ref)
    assert rel < 1e-2, (rel, mx)


def test_attention_with_bias():
    from gpu_util import attention, rel_err

    B, N, heads = 2, 577, 2
    Fd = heads * 64
    qkv = _mk((B, N, 3 * Fd), torch.bfloat16, 23)
    bias = _mk((heads, N, N), torch.bfloat16, 24)
    out = attention(qkv, heads, 0.125, bias=bias)
    q, k, v = qkv.float().reshape(B, N, 3, heads, 64).permute(2, 0, 3, 1, 4).unbind(0)
    a = (q * 0.125) @ k.transpose(-2, -1) + bias.float()[None]
    ref = (a.softmax(-1) @ v).transpose(1, 2).reshape(B, N, Fd)
    rel, mx = rel_err(out, ref)
    assert rel < 1e-2, (rel, mx)


@pytest.mark.parametrize("Fd", [128, 384, 768, 1024, 192])
def test_layernorm(Fd):
    from gpu_util import layernorm, rel_err

    M = 1297 * 2 + 3
    x = _mk((M, Fd), torch.float32, 25, 3.0) + 0.5
    w = 1 + 0.1 * _mk((Fd,), torch.float32, 26)
    b = 0.1 * _mk((Fd,), torch.float32, 27)
    y = layernorm(x, w, b, 1e-6, torch.bfloat16)
    ref = F.layer_norm(x, (Fd,), w, b, 1e-6)
    rel, mx = rel_err(y, ref)
    assert rel < 4e-3, (rel, mx)


@pytest.mark.parametrize("IH,IW,OH,OW", [(18, 18, 36, 36), (36, 20, 72, 40), (32, 32, 56, 56), (288, 288, 504, 504)])
def test_resize_bilinear_align_corners(IH, IW, OH, OW):
    from gpu_util import resize, rel_err

    x = _mk((2, IH, IW, 16), torch.bfloat16, 28)
    y = resize(x, OH, OW)
    ref = F.interpolate(x.float().permute(0, 3, 1, 2), size=(OH, OW), mode="bilinear", align_corners=True).permute(0, 2, 3, 1)
    rel, mx = rel_err(y, ref)
    assert rel < 4e-3, (rel, mx)


def test_attention_head_dim_32_with_window_bias():
    """SwinV2 shape: 32 features per head, per-window additive bias tables (b % n_windows)"""
    from gpu_util import attention, rel_err

    nW, B, N, heads = 4, 2, 144, 3
    Fd = heads * 32
    qkv = _mk((B * nW, N, 3 * Fd), torch.bfloat16, 31, 0.5)
    bias = _mk((nW, heads, N, N), torch.bfloat16, 32)
    out = attention(qkv, heads, 1.0, bias=bias, head_dim=32, bias_wmod=nW)
    q, k, v = qkv.float().reshape(B * nW, N, 3, heads, 32).permute(2, 0, 3, 1, 4).unbind(0)
    a = q @ k.transpose(-2, -1) + bias.float().repeat(B, 1, 1, 1)
    ref = (a.softmax(-1) @ v).transpose(1, 2).reshape(B * nW, N, Fd)
    rel, mx = rel_err(out, ref)
    assert rel < 1e-2, (rel, mx)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_attention_swinv2_large_window_24(dtype):
    """config W's real attention shape: 24 x 24 windows = 576 tokens (nine 64-column kv steps), 32 features per head,
    per-window additive tables (bias + shift mask, -100 on masked pairs, b % n_windows) - windowed_attention.py:65-123"""
    from gpu_util import attention, gate, rel_err

    nW, B, N, heads = 4, 2, 576, 6
    Fd = heads * 32
    g = torch.Generator(device="cpu").manual_seed(77)
    q = torch.nn.functional.normalize(torch.randn(B * nW, N, heads, 32, generator=g), dim=-1) * 10.0  # cosine logits * scale
    k = torch.nn.functional.normalize(torch.randn(B * nW, N, heads, 32, generator=g), dim=-1)
    v = torch.randn(B * nW, N, heads, 32, generator=g)
    qkv = torch.stack([q, k, v], dim=2).reshape(B * nW, N, 3 * Fd).to("cuda", dtype)
    bias = 16 * torch.sigmoid(torch.randn(nW, heads, N, N, generator=g))
    mask = (torch.rand(nW, 1, N, N, generator=g) < 0.3).float() * -100.0
    mask[0] = 0  # the unshifted window has no mask
    bias = (bias + mask).to("cuda", dtype)
    out = attention(qkv, heads, 1.0, bias=bias, head_dim=32, bias_wmod=nW)
    qf, kf, vf = qkv.float().reshape(B * nW, N, 3, heads, 32).permute(2, 0, 3, 1, 4).unbind(0)
    a = qf @ kf.transpose(-2, -1) + bias.float().repeat(B, 1, 1, 1)
    ref = (a.softmax(-1) @ vf).transpose(1, 2).reshape(B * nW, N, Fd)
    rel, mx = rel_err(out, ref)
    gate(f"op.attention.swin_w24.{'bf16' if dtype == torch.bfloat16 else 'fp16'}.rel_l2", rel, 1e-2 if dtype == torch.bfloat16 else 3e-3)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_attention_beit_large_shape_with_bias(dtype):
    """config E's attention shape: 577 tokens, 16 heads of 64, one additive bias table per head (image_encoder_model.py:335-354)"""
    from gpu_util import attention, gate, rel_err

    B, N, heads = 2, 577, 16
    Fd = heads * 64
    qkv = _mk((B, N, 3 * Fd), dtype, 53)
    bias = _mk((heads, N, N), dtype, 54, 2.0)
    out = attention(qkv, heads, 0.125, bias=bias)
    q, k, v = qkv.float().reshape(B, N, 3, heads, 64).permute(2, 0, 3, 1, 4).unbind(0)
    a = (q * 0.125) @ k.transpose(-2, -1) + bias.float()[None]
    ref = (a.softmax(-1) @ v).transpose(1, 2).reshape(B, N, Fd)
    rel, mx = rel_err(out, ref)
    gate(f"op.attention.beit_577.{'bf16' if dtype == torch.bfloat16 else 'fp16'}.rel_l2", rel, 1e-2 if dtype == torch.bfloat16 else 3e-3)


# ---- shapes large enough for the 2-CTA (cta_group::2, 256 x 256 tile) kernel: >= 74 tile pairs


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M,K,N,act", [(20001, 512, 768, 0), (12000, 1024, 1024, 1), (16300, 256, 512, 2)])
def test_gemm_two_cta_pairs(M, K, N, act, dtype):
    from gpu_util import conv_gemm, rel_err

    A = _mk((1, 1, M, K), dtype, 41)
    W = _mk((N, K), dtype, 42, K**-0.5)
    bias = _mk((N,), torch.float32, 43)
    out = conv_gemm(A, pack_linear(W), bias, act=act)
    pre = A.float().reshape(M, K) @ W.float().t() + bias
    ref = F.gelu(pre) if act == 1 else (F.relu(pre) if act == 2 else pre)
    rel, mx = rel_err(out.reshape(M, N), ref)
    assert rel < _tol(dtype), (rel, mx)
    assert torch.isfinite(out.float()).all()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M,K,h", [(300, 128, 344), (1297, 1536, 4096), (20752, 1536, 4096), (9000, 256, 96)])
def test_gemm_swiglu_epilogue(M, K, h, dtype):
    """act 4: the doubled inner Linear of the ViT-G FFN with interleaved rows (weights.interleave_swiglu); the epilogue
    writes silu(gate) * linear (misc_helpers.py:181-184) - single-CTA 128 / 256 tiles and CTA pairs, padded hidden"""
    from gpu_util import conv_gemm, rel_err
    from muggled_dpt_b200.weights import interleave_swiglu

    A = _mk((1, 1, M, K), dtype, 51)
    W12 = _mk((2 * h, K), dtype, 52, K**-0.5 * 1.5)
    b12 = _mk((2 * h,), torch.float32, 53, 0.3)
    Wp, bp = interleave_swiglu(W12, b12)
    hp = Wp.shape[0] // 2
    assert hp % 64 == 0 and hp >= h
    out = conv_gemm(A, pack_linear(Wp), bp, act=4).reshape(M, hp)
    pre = A.float().reshape(M, K) @ W12.float().t() + b12
    ref = F.silu(pre[:, :h]) * pre[:, h:]
    rel, mx = rel_err(out[:, :h], ref)
    assert rel < _tol(dtype), (rel, mx)
    assert torch.count_nonzero(out[:, h:]) == 0  # the padding columns are written, as zeros


def test_gemm_two_cta_f32_residual_inplace():
    from gpu_util import rel_err
    import ctypes as C
    from muggled_dpt_b200 import _native as NN

    M, K, N = 19999, 512, 1024
    A = _mk((1, 1, M, K), torch.bfloat16, 44)
    W = _mk((N, K), torch.bfloat16, 45, K**-0.5)
    bias = _mk((N,), torch.float32, 46)
    x = _mk((M, N), torch.float32, 47)
    ref = x + A.float().reshape(M, K) @ W.float().t() + bias
    Wp = pack_linear(W)
    rc = NN.lib().dpt_op_conv_gemm(A.data_ptr(), Wp.data_ptr(), bias.data_ptr(), x.data_ptr(), x.data_ptr(), None, None,
                                   1, 1, M, K, N, 1, 0, 0, 1, NN.DPT_BF16,
                                   C.c_void_p(torch.cuda.current_stream().cuda_stream))
    NN.check(rc, None, "gemm")
    torch.cuda.synchronize()
    rel, mx = rel_err(x, ref)
    assert rel < 1e-3, (rel, mx)


@pytest.mark.parametrize("M,K,N", [(5188, 1024, 1024), (5188, 4096, 1024), (6144, 512, 1000), (4800, 1536, 904)])
def test_gemm_wide_pair_tiles_f32_residual_inplace(M, K, N):
    """shapes the cost model gives to the 256 x 384 CTA-pair tiles (38..48 m-tiles, 897..1024 columns): short K (residual
    prefetch variant) and long K, a ragged last n-tile, an odd trailing m-tile"""
    from gpu_util import rel_err
    import ctypes as C
    from muggled_dpt_b200 import _native as NN

    A = _mk((1, 1, M, K), torch.bfloat16, 61)
    W = _mk((N, K), torch.bfloat16, 62, K**-0.5)
    bias = _mk((N,), torch.float32, 63)
    x = _mk((M, N), torch.float32, 64)
    ref = x + A.float().reshape(M, K) @ W.float().t() + bias
    Wp = pack_linear(W)
    rc = NN.lib().dpt_op_conv_gemm(A.data_ptr(), Wp.data_ptr(), bias.data_ptr(), x.data_ptr(), x.data_ptr(), None, None,
                                   1, 1, M, K, N, 1, 0, 0, 1, NN.DPT_BF16,
                                   C.c_void_p(torch.cuda.current_stream().cuda_stream))
    NN.check(rc, None, "gemm")
    torch.cuda.synchronize()
    rel, mx = rel_err(x, ref)
    assert rel < 1e-3, (rel, mx)


def test_conv3x3_two_cta_pairs_128_wide():
    """N = 128 convolution large enough (>= 296 tile pairs) for the 256 x 128 CTA-pair kernel (head c1 at full size)"""
    from gpu_util import conv_gemm, rel_err

    B, H, W, C, N = 2, 200, 200, 64, 128   # 2 * 13 * 25 = 650 M-tiles of 8 x 16 / 16 x 8 pixels
    x = _mk((B, H, W, C), torch.bfloat16, 58)
    w = _mk((N, C, 3, 3), torch.bfloat16, 59, (9 * C) ** -0.5)
    bias = _mk((N,), torch.float32, 60)
    for act in (0, 2):
        out = conv_gemm(x, pack_conv(w), bias, taps=9, act=act)
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, padding=1).permute(0, 2, 3, 1)
        ref = F.relu(ref) if act == 2 else ref
        rel, mx = rel_err(out, ref)
        assert rel < 6e-3, (act, rel, mx)


def test_conv3x3_two_cta_pairs():
    from gpu_util import conv_gemm, rel_err

    B, H, W, C, N = 3, 144, 144, 64, 256   # 3 * 162 = 486 M-tiles
    x = _mk((B, H, W, C), torch.bfloat16, 48)
    w = _mk((N, C, 3, 3), torch.bfloat16, 49, (9 * C) ** -0.5)
    bias = _mk((N,), torch.float32, 50)
    a1 = _mk((B, H, W, N), torch.bfloat16, 51)
    out, out_relu = conv_gemm(x, pack_conv(w), bias, add1=a1, want_relu=True, taps=9)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, padding=1).permute(0, 2, 3, 1) + a1.float()
    rel, mx = rel_err(out, ref)
    assert rel < 8e-3, (rel, mx)
    assert torch.equal(out_relu, torch.relu(out))

