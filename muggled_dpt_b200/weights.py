"""
Weight loader for the B200 depth path: upstream checkpoint -> (config dict, packed device tensors).

Behaviour mirrors the reference loader for Depth-Anything V2
(muggled_dpt/v2_depthanything/state_dict_conversion/config_from_original_state_dict.py:17-43 and
 convert_original_state_dict_keys.py:15-86,295-317): the model hyper-parameters are inferred from tensor shapes, the
position embedding is split into cls / patch parts, `pretrained.mask_token` and
`depth_head.scratch.refinenet4.resConfUnit1.*` are silently dropped. Instead of renaming keys into five nn.Module
state dicts, the tensors are packed straight into the layouts the sm_100a kernels consume:

  * every matmul weight is [N, taps * kpad] K-major 16-bit with kpad = roundup(Cin, 64) (zero padded), tap = ky*3+kx
    for 3x3 kernels (gemm_tc.cuh); ConvTranspose2d(k=s) becomes s*s stacked [Cout, kpad] matrices, one per (ky, kx);
  * LayerScale is folded into the preceding Linear: gamma*(Wx+b) = (gamma*W)x + gamma*b (transformer_block.py:58,63);
  * biases, LayerNorm parameters and the position embedding stay fp32.
"""

from __future__ import annotations

import math
import re

import torch

GEMM_K = 64


# ---------------------------------------------------------------------------------------------------------------------
# model-type sniffing (muggled_dpt/make_dpt.py:78-116)


def determine_model_type_from_state_dict(model_path: str, state_dict: dict) -> str:
    import os.path as osp

    keys = state_dict.keys()
    if "pretrained.model.layers.0.blocks.0.attn.logit_scale" in keys:
        return "swinv2"
    if "pretrained.model.blocks.0.attn.relative_position_bias_table" in keys:
        return "beit"
    if "pretrained.blocks.0.ls1.gamma" in keys:
        name = osp.basename(model_path).lower()
        is_v2 = "v2" in name
        is_v1 = (not is_v2) and (("anything_vit" in name) or ("v1" in name))
        if (not is_v1) and (not is_v2):
            print("", "WARNING: Unable to determine DepthAnything model version!", "-> Will assume v2",
                  "-> Will use v1 if the file name contains 'v1'", sep="\n")
        return "depthanythingv1" if is_v1 else "depthanythingv2"
    return "unknown"


# ---------------------------------------------------------------------------------------------------------------------
# config inference - same keys, same order as the reference's config dict


def get_model_config_from_state_dict(state_dict: dict, enable_cache: bool, enable_optimizations: bool) -> dict:
    def need(key):
        assert key in state_dict, f"Error determining model config! Couldn't find {key} key"
        return state_dict[key]

    pe = need("pretrained.patch_embed.proj.weight")
    features_per_token = int(pe.shape[0])
    patch_size_px = int(pe.shape[3])
    block_ids = [int(m.group(1)) for m in (re.match(r"pretrained\.blocks\.(\d+)\.", k) for k in state_dict) if m]
    assert block_ids and max(block_ids) > 0, "Error determining number of transformer blocks! Could not find any blocks"
    reasm = [int(need(f"depth_head.scratch.layer{i}_rn.weight").shape[1]) for i in (1, 2, 3, 4)]
    num_tokens = int(need("pretrained.pos_embed").shape[1]) - 1
    base = int(math.isqrt(num_tokens))
    return {
        "features_per_token": features_per_token,
        "num_blocks": 1 + max(block_ids),
        "num_heads": features_per_token // 64,
        "reassembly_features_list": reasm,
        "fusion_channels": int(need("depth_head.scratch.layer1_rn.weight").shape[0]),
        "patch_size_px": patch_size_px,
        "base_patch_grid_hw": (base, base),
        "is_giant": "pretrained.blocks.0.mlp.w12.weight" in state_dict,
        "is_metric": "is_metric" in state_dict,
        "enable_cache": enable_cache,
        "enable_optimizations": enable_optimizations,
    }


def get_model_config_from_v1_state_dict(state_dict: dict, enable_cache: bool, enable_optimizations: bool) -> dict:
    """v1_depthanything/state_dict_conversion/config_from_original_state_dict.py:17-37: the V2 inference without the
    is_giant / is_metric keys (same weight schema, same key order otherwise)"""
    cfg = get_model_config_from_state_dict(state_dict, enable_cache, enable_optimizations)
    cfg.pop("is_giant")
    cfg.pop("is_metric")
    return cfg


# ---------------------------------------------------------------------------------------------------------------------
# packing helpers (pure tensor reshapes - unit-tested on CPU in tests/test_weights.py)


def _roundup(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def pack_linear(w: torch.Tensor) -> torch.Tensor:
    """[N, K] -> [N, roundup(K, 64)] zero padded"""
    n, k = w.shape
    out = w.new_zeros(n, _roundup(k, GEMM_K))
    out[:, :k] = w
    return out


def fold_layernorm(w: torch.Tensor, b, ln_w: torch.Tensor, ln_b: torch.Tensor):
    """Linear(LN(x)) with LN(x) = xhat * ln_w + ln_b (transformer_block.py:53-65: the norm feeds exactly one Linear):
    W' = W * ln_w[None, :], b' = b + W @ ln_b. The GEMM then runs on the raw residual stream and its epilogue applies
    the per-row mean / rstd (csrc/gemm_tc.cuh "folded LayerNorm"). Returns (packed W', b'); kind "half_colsum" makes
    DPTModel.to() emit the column sums of the 16-bit W' next to it."""
    w = w.to(torch.float32)
    b2 = w @ ln_b.to(torch.float32)
    if b is not None:
        b2 = b2 + b.to(torch.float32)
    return pack_linear(w * ln_w.to(torch.float32)[None, :]), b2


def interleave_swiglu(w12: torch.Tensor, b12: torch.Tensor):
    """Doubled inner Linear of the SwiGLU FFN (components/misc_helpers.py:161-184: rows [0, h) are the gate half, rows
    [h, 2h) the linear half) -> rows in blocks of 64 = 32 gate rows followed by the 32 linear rows of the same features,
    h zero-padded to hp = roundup(h, 64): [2 hp, K] and [2 hp]. The GEMM epilogue (csrc/gemm_tc.cuh ACT_SWIGLU) then
    holds both halves of a feature in one thread and writes silu(gate) * linear directly; the zero rows produce the
    zero columns [h, hp) that the outer Linear's K padding expects."""
    h = w12.shape[0] // 2
    hp = _roundup(h, GEMM_K)
    j = torch.arange(h)
    dst_gate = (j // 32) * 64 + (j % 32)
    w = w12.new_zeros(2 * hp, w12.shape[1])
    b = b12.new_zeros(2 * hp)
    w[dst_gate], w[dst_gate + 32] = w12[:h], w12[h:]
    b[dst_gate], b[dst_gate + 32] = b12[:h], b12[h:]
    return w, b


def pack_conv(w: torch.Tensor) -> torch.Tensor:
    """Conv2d weight [Cout, Cin, kh, kw] -> [Cout, kh*kw*kpad], column = (ky*kw + kx)*kpad + ci"""
    co, ci, kh, kw = w.shape
    kpad = _roundup(ci, GEMM_K)
    out = w.new_zeros(co, kh * kw, kpad)
    out[:, :, :ci] = w.permute(0, 2, 3, 1).reshape(co, kh * kw, ci)
    return out.reshape(co, kh * kw * kpad)


CONVT_CH = 32  # ConvTranspose output channels are padded to whole 32-column GEMM tiles


def pack_conv_transpose(w: torch.Tensor) -> torch.Tensor:
    """ConvTranspose2d(k = stride = s) weight [Cin, Cout, s, s] -> [s*s*cop, kpad], cop = roundup(Cout, 32); block
    (ky*s + kx) holds W_sub[co, ci] = w[ci, co, ky, kx], so out[b, y*s+ky, x*s+kx, co] = sum_ci in[b,y,x,ci] *
    W_sub[co, ci] + bias[co] (reassembly_model.py:262-269). The zero rows [Cout, cop) make every sub-pixel block a whole
    number of GEMM n-tiles, so all s*s blocks run as ONE launch with pixel-shuffled stores for any channel count (ViT-S:
    48 -> 64); the padded output channels are zeros that the next convolution's zero K-padding ignores."""
    ci, co, s, s2 = w.shape
    assert s == s2
    kpad = _roundup(ci, GEMM_K)
    cop = _roundup(co, CONVT_CH)
    out = w.new_zeros(s * s, cop, kpad)
    out[:, :co, :ci] = w.permute(2, 3, 1, 0).reshape(s * s, co, ci)
    return out.reshape(s * s * cop, kpad)


def pad_conv_transpose_bias(b):
    """bias [Cout] -> [roundup(Cout, 32)] (zeros), matching pack_conv_transpose"""
    if b is None:
        return None
    out = b.new_zeros(_roundup(b.shape[0], CONVT_CH))
    out[: b.shape[0]] = b
    return out


def pack_patch_embed(w: torch.Tensor) -> torch.Tensor:
    """[F, 3, P, P] -> [F, roundup(3*P*P, 64)], column = c*P*P + ky*P + kx (patch_embed.py:92-97)"""
    return pack_linear(w.reshape(w.shape[0], -1))


# ---------------------------------------------------------------------------------------------------------------------
# strict loading: unexpected keys
#
# The reference converts a checkpoint key by key; keys that match none of its patterns never reach
# `load_state_dict` and are silently ignored, keys on its drop list return None, and every other converted key must
# exist in the module tree or `load_state_dict(strict=True)` raises "Unexpected key(s)". The packers below read the
# checkpoint through `_Tracked`, so a key under one of the reference's converted prefixes that no packer consumed is
# reported the same way (v2_depthanything/state_dict_conversion/convert_original_state_dict_keys.py:42-70,
# v31_beit/.../convert_midas_state_dict_keys.py:103-325, v31_swinv2/.../convert_midas_state_dict_keys.py:104-349).


class _Tracked:
    def __init__(self, sd: dict):
        self.sd = sd
        self.used = set()

    def __contains__(self, key):
        return key in self.sd

    def __getitem__(self, key):
        self.used.add(key)
        return self.sd[key]

    def unexpected(self, converted: tuple, dropped: tuple) -> list:
        out = []
        for k in self.sd:
            k = str(k)
            if k in self.used or not any(re.match(c, k) for c in converted) or any(re.search(d, k) for d in dropped):
                continue
            out.append(k)
        return out


def _raise_unexpected(keys: list):
    if keys:
        raise RuntimeError("Error(s) in loading state_dict: Unexpected key(s): " + ", ".join(keys[:8]) +
                           (" ..." if len(keys) > 8 else ""))


_DAV2_CONVERTED = (r"pretrained\.patch_embed", r"pretrained\.cls_token$", r"pretrained\.pos_embed$", r"pretrained\.norm",
                   r"pretrained\.blocks\.\d+", r"depth_head\.projects", r"depth_head\.resize_layers",
                   r"depth_head\.scratch\.layer\d+_rn", r"depth_head\.scratch\.refinenet", r"depth_head\.scratch\.output_conv")
_DAV2_DROPPED = (r"^depth_head\.scratch\.refinenet4\.resConfUnit1",)
_BEIT_CONVERTED = (r"pretrained\.model\.patch_embed", r"pretrained\.model\.cls_token$", r"pretrained\.model\.blocks\.\d+",
                   r"pretrained\.act_postprocess", r"scratch\.layer\d+_rn", r"scratch\.refinenet", r"scratch\.output_conv")
_BEIT_DROPPED = (r"relative_position_index", r"^scratch\.refinenet4\.resConfUnit1")
_SWIN_CONVERTED = (r"pretrained\.model\.patch_embed\.", r"pretrained\.model\.layers\.\d+\.blocks\.\d+",
                   r"pretrained\.model\.layers\.\d+\.downsample", r"pretrained\.act_postprocess", r"scratch\.layer\d+_rn",
                   r"scratch\.refinenet", r"scratch\.output_conv")
_SWIN_DROPPED = (r"attn_mask", r"^scratch\.refinenet4\.resConfUnit1")


def pack_depthanything_v2(sd: dict, cfg: dict, strict: bool = True) -> dict:
    """Returns {packed name: (fp32 cpu tensor, kind)} with kind in {"half", "half_colsum", "f32", "host"}."""
    F = cfg["features_per_token"]
    L = cfg["num_blocks"]
    missing = []
    sd = _Tracked(sd)

    def get(key, default_shape=None, fill=0.0):
        if key in sd:
            return sd[key].detach().to(torch.float32).cpu()
        missing.append(key)
        if strict or default_shape is None:
            return None
        return torch.full(default_shape, fill)

    out = {}

    def put(name, t, kind):
        if t is not None:
            out[name] = (t.contiguous(), kind)

    pw = get("pretrained.patch_embed.proj.weight")
    put("patch.w", pack_patch_embed(pw) if pw is not None else None, "half")
    put("patch.b", get("pretrained.patch_embed.proj.bias", (F,)), "f32")
    pos = get("pretrained.pos_embed")
    if pos is not None:
        put("pos.cls_emb", pos[0, 0, :].clone(), "f32")
        put("pos.base", pos[0, 1:, :].clone(), "f32")
    cls = get("pretrained.cls_token", (1, 1, F))
    put("pos.cls_tok", cls.reshape(-1) if cls is not None else None, "f32")
    for i in range(L):
        s, d = f"pretrained.blocks.{i}.", f"blk{i}."
        l1w, l1b = get(s + "norm1.weight", (F,), 1.0), get(s + "norm1.bias", (F,))
        l2w, l2b = get(s + "norm2.weight", (F,), 1.0), get(s + "norm2.bias", (F,))
        qw, qb = get(s + "attn.qkv.weight", (3 * F, F)), get(s + "attn.qkv.bias", (3 * F,))
        if all(t is not None for t in (l1w, l1b, qw, qb)):
            w_, b_ = fold_layernorm(qw, qb, l1w, l1b)
            put(d + "qkv.w", w_, "half_colsum")
            put(d + "qkv.b", b_, "f32")
        g1 = get(s + "ls1.gamma", (F,), 1.0)
        g2 = get(s + "ls2.gamma", (F,), 1.0)
        w, b = get(s + "attn.proj.weight", (F, F)), get(s + "attn.proj.bias", (F,))
        if all(t is not None for t in (g1, w, b)):
            put(d + "proj.w", pack_linear(g1[:, None] * w), "half")
            put(d + "proj.b", g1 * b, "f32")
        if cfg.get("is_giant", False):
            # ViT-G: SwiGLU FFN (components/misc_helpers.py:125-185) - w12 is the doubled inner Linear (gate half first),
            # w3 the outer Linear; they take the places of fc1 / fc2 (dpt_config.mlp_swiglu), fc1 with its rows
            # interleaved for the fused gate epilogue
            w1, b1 = get(s + "mlp.w12.weight"), get(s + "mlp.w12.bias")
            w, b = get(s + "mlp.w3.weight"), get(s + "mlp.w3.bias")
            if w1 is not None and b1 is not None:
                w1, b1 = interleave_swiglu(w1, b1)
        else:
            w1, b1 = get(s + "mlp.fc1.weight", (4 * F, F)), get(s + "mlp.fc1.bias", (4 * F,))
            w, b = get(s + "mlp.fc2.weight", (F, 4 * F)), get(s + "mlp.fc2.bias", (F,))
        if all(t is not None for t in (l2w, l2b, w1, b1)):
            w_, b_ = fold_layernorm(w1, b1, l2w, l2b)
            put(d + "fc1.w", w_, "half_colsum")
            put(d + "fc1.b", b_, "f32")
        if all(t is not None for t in (g2, w, b)):
            put(d + "fc2.w", pack_linear(g2[:, None] * w), "half")
            put(d + "fc2.b", g2 * b, "f32")
    put("outnorm.w", get("pretrained.norm.weight", (F,), 1.0), "f32")
    put("outnorm.b", get("pretrained.norm.bias", (F,)), "f32")

    for k in range(4):
        d = f"reasm{k}."
        w = get(f"depth_head.projects.{k}.weight")
        put(d + "proj.w", pack_linear(w.reshape(w.shape[0], -1)) if w is not None else None, "half")
        put(d + "proj.b", get(f"depth_head.projects.{k}.bias"), "f32")
        if k in (0, 1):
            w = get(f"depth_head.resize_layers.{k}.weight")
            put(d + "up.w", pack_conv_transpose(w) if w is not None else None, "half")
            put(d + "up.b", pad_conv_transpose_bias(get(f"depth_head.resize_layers.{k}.bias")), "f32")
        elif k == 3:
            w = get("depth_head.resize_layers.3.weight")
            put(d + "down.w", pack_conv(w) if w is not None else None, "half")
            put(d + "down.b", get("depth_head.resize_layers.3.bias"), "f32")
        w = get(f"depth_head.scratch.layer{k + 1}_rn.weight")
        put(d + "fuse.w", pack_conv(w) if w is not None else None, "half")

    for lvl in range(4):
        s, d = f"depth_head.scratch.refinenet{lvl + 1}.", f"fus{lvl}."
        units = (("rcu1", "resConfUnit1"), ("rcu2", "resConfUnit2")) if lvl < 3 else (("rcu2", "resConfUnit2"),)
        for dn, sn in units:  # refinenet4.resConfUnit1 is dropped (convert_original_state_dict_keys.py:230)
            for cv in (1, 2):
                w = get(f"{s}{sn}.conv{cv}.weight")
                put(f"{d}{dn}.c{cv}.w", pack_conv(w) if w is not None else None, "half")
                put(f"{d}{dn}.c{cv}.b", get(f"{s}{sn}.conv{cv}.bias"), "f32")
        w = get(s + "out_conv.weight")
        put(d + "out.w", pack_linear(w.reshape(w.shape[0], -1)) if w is not None else None, "half")
        put(d + "out.b", get(s + "out_conv.bias"), "f32")

    s = "depth_head.scratch."
    w = get(s + "output_conv1.weight")
    put("head.c1.w", pack_conv(w) if w is not None else None, "half")
    put("head.c1.b", get(s + "output_conv1.bias"), "f32")
    w = get(s + "output_conv2.0.weight")
    put("head.c2.w", pack_conv(w) if w is not None else None, "half")
    put("head.c2.b", get(s + "output_conv2.0.bias"), "f32")
    w = get(s + "output_conv2.2.weight")
    put("head.c3.w_host", w.reshape(-1) if w is not None else None, "host")
    put("head.c3.b_host", get(s + "output_conv2.2.bias"), "host")

    if missing and strict:
        raise RuntimeError("Error(s) in loading state_dict: Missing key(s): " + ", ".join(missing[:8]) +
                           (" ..." if len(missing) > 8 else ""))
    if strict:
        _raise_unexpected(sd.unexpected(_DAV2_CONVERTED, _DAV2_DROPPED))
    return out


# =====================================================================================================================
# MiDaS v3.1 BEiT (muggled_dpt/v31_beit/state_dict_conversion/*): config inference + packing
# =====================================================================================================================


def get_model_config_from_midas_beit_state_dict(state_dict: dict, enable_cache: bool, enable_optimizations: bool) -> dict:
    """config_from_midas_state_dict.py:17-40 - same keys, same order as the reference's BEiT config dict"""
    def need(key):
        assert key in state_dict, f"Error determining model config! Couldn't find {key} key"
        return state_dict[key]

    pe = need("pretrained.model.patch_embed.proj.weight")
    block_ids = [int(m.group(1)) for m in (re.match(r"pretrained\.model\.blocks\.(\d+)\.", k) for k in state_dict) if m]
    assert block_ids and max(block_ids) > 0, "Error determining number of transformer blocks! Could not find any blocks"
    table = need("pretrained.model.blocks.0.attn.relative_position_bias_table")
    side = int(math.isqrt(int(table.shape[0]) - 3))  # 2g - 1
    base = (side + 1) // 2
    return {
        "features_per_token": int(pe.shape[0]),
        "num_blocks": 1 + max(block_ids),
        "num_heads": int(table.shape[1]),
        "reassembly_features_list": [int(need(f"scratch.layer{i}_rn.weight").shape[1]) for i in (1, 2, 3, 4)],
        "fusion_channels": int(need("scratch.layer1_rn.weight").shape[0]),
        "patch_size_px": int(pe.shape[3]),
        "base_patch_grid_hw": (base, base),
        "enable_cache": enable_cache,
        "enable_optimizations": enable_optimizations,
    }


def pack_beit(sd: dict, cfg: dict, strict: bool = True) -> dict:
    """Returns {packed name: (fp32 cpu tensor, kind)}. Dropped like the reference does
    (convert_midas_state_dict_keys.py:160-162,264): `relative_position_index`, `scratch.refinenet4.resConfUnit1.*`.
    q_bias / v_bias become the bias of the fused QKV GEMM ([q_bias, 0, v_bias]; K has none - image_encoder_model.py:341)."""
    F = cfg["features_per_token"]
    L = cfg["num_blocks"]
    missing = []
    sd = _Tracked(sd)

    def get(key):
        if key in sd:
            return sd[key].detach().to(torch.float32).cpu()
        missing.append(key)
        return None

    out = {}

    def put(name, t, kind):
        if t is not None:
            out[name] = (t.contiguous(), kind)

    def lin(key):
        w = get(key)
        return pack_linear(w.reshape(w.shape[0], -1)) if w is not None else None

    def conv(key):
        w = get(key)
        return pack_conv(w) if w is not None else None

    pw = get("pretrained.model.patch_embed.proj.weight")
    put("patch.w", pack_patch_embed(pw) if pw is not None else None, "half")
    put("patch.b", get("pretrained.model.patch_embed.proj.bias"), "f32")
    cls = get("pretrained.model.cls_token")
    put("beit.cls", cls.reshape(-1) if cls is not None else None, "f32")
    for i in range(L):
        s, d = f"pretrained.model.blocks.{i}.", f"blk{i}."
        l1w, l1b = get(s + "norm1.weight"), get(s + "norm1.bias")
        l2w, l2b = get(s + "norm2.weight"), get(s + "norm2.bias")
        qw, qb, vb = get(s + "attn.qkv.weight"), get(s + "attn.q_bias"), get(s + "attn.v_bias")
        if all(t is not None for t in (l1w, l1b, qw, qb, vb)):
            w_, b_ = fold_layernorm(qw, torch.cat([qb.reshape(-1), torch.zeros(F), vb.reshape(-1)]), l1w, l1b)
            put(d + "qkv.w", w_, "half_colsum")
            put(d + "qkv.b", b_, "f32")
        put(d + "relpos.table", get(s + "attn.relative_position_bias_table"), "f32")
        g1, g2 = get(s + "gamma_1"), get(s + "gamma_2")
        w, b = get(s + "attn.proj.weight"), get(s + "attn.proj.bias")
        if all(t is not None for t in (g1, w, b)):
            put(d + "proj.w", pack_linear(g1[:, None] * w), "half")
            put(d + "proj.b", g1 * b, "f32")
        w1, b1 = get(s + "mlp.fc1.weight"), get(s + "mlp.fc1.bias")
        if all(t is not None for t in (l2w, l2b, w1, b1)):
            w_, b_ = fold_layernorm(w1, b1, l2w, l2b)
            put(d + "fc1.w", w_, "half_colsum")
            put(d + "fc1.b", b_, "f32")
        w, b = get(s + "mlp.fc2.weight"), get(s + "mlp.fc2.bias")
        if all(t is not None for t in (g2, w, b)):
            put(d + "fc2.w", pack_linear(g2[:, None] * w), "half")
            put(d + "fc2.b", g2 * b, "f32")

    for k in range(4):
        s, d = f"pretrained.act_postprocess{k + 1}.", f"reasm{k}."
        w = get(s + "0.project.0.weight")  # Linear(2F, F) on [patch, cls] (readout_projection.py:74-79)
        if w is not None:
            put(d + "readout.w1", pack_linear(w[:, :F]), "half")
            put(d + "readout.w2", w[:, F:].clone(), "f32")
        put(d + "readout.b", get(s + "0.project.0.bias"), "f32")
        put(d + "proj.w", lin(s + "3.weight"), "half")
        put(d + "proj.b", get(s + "3.bias"), "f32")
        if k in (0, 1):
            w = get(s + "4.weight")
            put(d + "up.w", pack_conv_transpose(w) if w is not None else None, "half")
            put(d + "up.b", pad_conv_transpose_bias(get(s + "4.bias")), "f32")
        elif k == 3:
            put(d + "down.w", conv(s + "4.weight"), "half")
            put(d + "down.b", get(s + "4.bias"), "f32")
        put(d + "fuse.w", conv(f"scratch.layer{k + 1}_rn.weight"), "half")

    for lvl in range(4):
        s, d = f"scratch.refinenet{lvl + 1}.", f"fus{lvl}."
        units = (("rcu1", "resConfUnit1"), ("rcu2", "resConfUnit2")) if lvl < 3 else (("rcu2", "resConfUnit2"),)
        for dn, sn in units:
            for cv in (1, 2):
                put(f"{d}{dn}.c{cv}.w", conv(f"{s}{sn}.conv{cv}.weight"), "half")
                put(f"{d}{dn}.c{cv}.b", get(f"{s}{sn}.conv{cv}.bias"), "f32")
        put(d + "out.w", lin(s + "out_conv.weight"), "half")
        put(d + "out.b", get(s + "out_conv.bias"), "f32")

    put("head.c1.w", conv("scratch.output_conv.0.weight"), "half")
    put("head.c1.b", get("scratch.output_conv.0.bias"), "f32")
    put("head.c2.w", conv("scratch.output_conv.2.weight"), "half")
    put("head.c2.b", get("scratch.output_conv.2.bias"), "f32")
    w = get("scratch.output_conv.4.weight")
    put("head.c3.w_host", w.reshape(-1) if w is not None else None, "host")
    put("head.c3.b_host", get("scratch.output_conv.4.bias"), "host")

    if missing:
        if strict:
            raise RuntimeError("Error(s) in loading state_dict: Missing key(s): " + ", ".join(missing[:8]) +
                               (" ..." if len(missing) > 8 else ""))
        raise RuntimeError("non-strict loading of BEiT checkpoints with missing keys is not supported: " + missing[0])
    if strict:
        _raise_unexpected(sd.unexpected(_BEIT_CONVERTED, _BEIT_DROPPED))
    return out


# =====================================================================================================================
# MiDaS v3.1 SwinV2 (muggled_dpt/v31_swinv2/state_dict_conversion/*): config inference + packing
# =====================================================================================================================


def get_model_config_from_midas_swinv2_state_dict(state_dict: dict, enable_cache: bool, enable_optimizations: bool) -> dict:
    """config_from_midas_state_dict.py:17-214 - same keys, same order as the reference's SwinV2 config dict. Window
    size and base grid come from the first stored `attn_mask` ([nW, A, A]); the pretrained-window LUT is the reference's."""
    pe_key = "pretrained.model.patch_embed.proj.weight"
    assert pe_key in state_dict, f"Error determining transformer features per token! Couldn't find {pe_key} key"
    f0, patch = int(state_dict[pe_key].shape[0]), int(state_dict[pe_key].shape[3])
    heads, layers = {}, {}
    for k in state_dict:
        m = re.match(r"pretrained\.model\.layers\.(\d+)\.blocks\.(\d+)\.", k)
        if m:
            st, bi = int(m.group(1)), int(m.group(2))
            layers[st] = max(layers.get(st, 0), bi + 1)
            if k.endswith("logit_scale"):
                heads[st] = int(state_dict[k].shape[0])
    assert len(heads) == 4, f"Expecting 4 stages in swinv2 dpt, got: {len(heads)}"
    assert len(layers) == 4, f"Expecting 4 stages in swinv2 dpt, got: {len(layers)}"
    mask_keys = sorted(k for k in state_dict if k.endswith("attn_mask"))
    assert mask_keys, "Error, couldn't find attn_mask key, can't determine window size!"
    num_windows, window_area = state_dict[mask_keys[0]].shape[0:2]
    win = int(math.sqrt(int(window_area)))
    base = int(math.sqrt(int(num_windows) * int(window_area)))
    pretrained = {16: [16, 16, 16, 8], 24: [12, 12, 12, 6]}.get(win, [None] * 4)
    lrn = "scratch.layer1_rn.weight"
    assert lrn in state_dict, f"Error determining fusion channel count! Couldn't find {lrn} key"
    return {
        "features_per_stage": [f0 * (2**i) for i in range(4)],
        "heads_per_stage": [heads[i] for i in range(4)],
        "layers_per_stage": [layers[i] for i in range(4)],
        "base_patch_grid_hw": (base, base),
        "window_size_hw": (win, win),
        "pretrained_window_sizes_per_stage": pretrained,
        "fusion_channels": int(state_dict[lrn].shape[0]),
        "patch_size_px": patch,
        "enable_cache": enable_cache,
        "enable_optimizations": enable_optimizations,
    }


def pack_swinv2(sd: dict, cfg: dict, strict: bool = True) -> dict:
    """Returns {packed name: (fp32 cpu tensor, kind)}. Mirrors convert_midas_state_dict_keys.py: stored `attn_mask`
    tensors are dropped (:179-181), `logit_scale` is clamped at ln(100) and exponentiated once (:115-131), q/v biases
    become the bias of the fused QKV GEMM."""
    missing = []
    sd = _Tracked(sd)

    def get(key):
        if key in sd:
            return sd[key].detach().to(torch.float32).cpu()
        missing.append(key)
        return None

    out = {}

    def put(name, t, kind):
        if t is not None:
            out[name] = (t.contiguous(), kind)

    def lin(key):
        w = get(key)
        return pack_linear(w.reshape(w.shape[0], -1)) if w is not None else None

    def conv(key):
        w = get(key)
        return pack_conv(w) if w is not None else None

    pw = get("pretrained.model.patch_embed.proj.weight")
    put("patch.w", pack_patch_embed(pw) if pw is not None else None, "half")
    put("patch.b", get("pretrained.model.patch_embed.proj.bias"), "f32")
    put("patch.ln.w", get("pretrained.model.patch_embed.norm.weight"), "f32")
    put("patch.ln.b", get("pretrained.model.patch_embed.norm.bias"), "f32")
    for st in range(4):
        Fs = cfg["features_per_stage"][st]
        for bi in range(cfg["layers_per_stage"][st]):
            s, d = f"pretrained.model.layers.{st}.blocks.{bi}.", f"sw{st}.{bi}."
            put(d + "qkv.w", lin(s + "attn.qkv.weight"), "half")
            qb, vb = get(s + "attn.q_bias"), get(s + "attn.v_bias")
            if qb is not None and vb is not None:
                put(d + "qkv.b", torch.cat([qb.reshape(-1), torch.zeros(Fs), vb.reshape(-1)]), "f32")
            ls = get(s + "attn.logit_scale")
            if ls is not None:
                put(d + "logit", torch.clamp(ls.reshape(-1), max=math.log(1.0 / 0.01)).exp(), "f32")
            put(d + "cpb.w1", get(s + "attn.cpb_mlp.0.weight"), "f32")
            put(d + "cpb.b1", get(s + "attn.cpb_mlp.0.bias"), "f32")
            put(d + "cpb.w2", get(s + "attn.cpb_mlp.2.weight"), "f32")
            put(d + "proj.w", lin(s + "attn.proj.weight"), "half")
            put(d + "proj.b", get(s + "attn.proj.bias"), "f32")
            put(d + "ln1.w", get(s + "norm1.weight"), "f32")
            put(d + "ln1.b", get(s + "norm1.bias"), "f32")
            put(d + "ln2.w", get(s + "norm2.weight"), "f32")
            put(d + "ln2.b", get(s + "norm2.bias"), "f32")
            put(d + "fc1.w", lin(s + "mlp.fc1.weight"), "half")
            put(d + "fc1.b", get(s + "mlp.fc1.bias"), "f32")
            put(d + "fc2.w", lin(s + "mlp.fc2.weight"), "half")
            put(d + "fc2.b", get(s + "mlp.fc2.bias"), "f32")
        if st < 3:
            s, d = f"pretrained.model.layers.{st}.downsample.", f"sw{st}.merge."
            put(d + "w", lin(s + "reduction.weight"), "half")
            put(d + "ln.w", get(s + "norm.weight"), "f32")
            put(d + "ln.b", get(s + "norm.bias"), "f32")
    for k in range(4):
        put(f"reasm{k}.fuse.w", conv(f"scratch.layer{k + 1}_rn.weight"), "half")
    for lvl in range(4):
        s, d = f"scratch.refinenet{lvl + 1}.", f"fus{lvl}."
        units = (("rcu1", "resConfUnit1"), ("rcu2", "resConfUnit2")) if lvl < 3 else (("rcu2", "resConfUnit2"),)
        for dn, sn in units:
            for cv in (1, 2):
                put(f"{d}{dn}.c{cv}.w", conv(f"{s}{sn}.conv{cv}.weight"), "half")
                put(f"{d}{dn}.c{cv}.b", get(f"{s}{sn}.conv{cv}.bias"), "f32")
        put(d + "out.w", lin(s + "out_conv.weight"), "half")
        put(d + "out.b", get(s + "out_conv.bias"), "f32")
    put("head.c1.w", conv("scratch.output_conv.0.weight"), "half")
    put("head.c1.b", get("scratch.output_conv.0.bias"), "f32")
    put("head.c2.w", conv("scratch.output_conv.2.weight"), "half")
    put("head.c2.b", get("scratch.output_conv.2.bias"), "f32")
    w = get("scratch.output_conv.4.weight")
    put("head.c3.w_host", w.reshape(-1) if w is not None else None, "host")
    put("head.c3.b_host", get("scratch.output_conv.4.bias"), "host")
    if missing:
        raise RuntimeError("Error(s) in loading state_dict: Missing key(s): " + ", ".join(missing[:8]) +
                           (" ..." if len(missing) > 8 else ""))
    if strict:
        _raise_unexpected(sd.unexpected(_SWIN_CONVERTED, _SWIN_DROPPED))
    return out
