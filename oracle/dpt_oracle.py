"""
ORACLE - TEST INFRASTRUCTURE ONLY. Nothing under muggled_dpt_b200/ may import this module; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs do, and only as the checker / the timed
CPU baseline - never as the product path.

A flat, functional, fp32, CPU restatement of the reference's single-image depth inference path
(muggled_dpt/dpt_model.py:61-83) for Depth-Anything V2 (DINOv2 ViT-S/B/L + DPT head), written directly against the
UPSTREAM checkpoint key names (the format make_dpt_from_state_dict() loads), so it shares no code and no key-renaming
logic with either the reference or the product. The arithmetic itself lives in PyTorch (third-party; the reference
pins torch>=2.1,<2.10, this image has 2.11): every op below is the same ATen call the reference module makes, cited
file:line. Parity pinning: the reference ships no tests / golden vectors, so this oracle is pinned
against the reference itself, imported in the build container from /root/reference by oracle/make_golden.py, on seeded
synthetic checkpoints; the resulting fixtures are committed under tests/golden/ and checked by
tests/test_oracle_golden.py (stage by stage).
"""

from __future__ import annotations

import math

import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------------------------------
# config inference (v2_depthanything/state_dict_conversion/config_from_original_state_dict.py:17-43)


def infer_config(sd: dict) -> dict:
    feats = int(sd["pretrained.patch_embed.proj.weight"].shape[0])
    patch = int(sd["pretrained.patch_embed.proj.weight"].shape[3])
    blocks = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("pretrained.blocks."))
    ntok = int(sd["pretrained.pos_embed"].shape[1]) - 1
    base = int(math.isqrt(ntok))
    reasm = [int(sd[f"depth_head.scratch.layer{i}_rn.weight"].shape[1]) for i in (1, 2, 3, 4)]
    return {
        "features_per_token": feats,
        "num_blocks": blocks,
        "num_heads": feats // 64,  # :78-90
        "reassembly_features_list": reasm,
        "fusion_channels": int(sd["depth_head.scratch.layer1_rn.weight"].shape[0]),
        "patch_size_px": patch,
        "base_patch_grid_hw": (base, base),
        "is_metric": "is_metric" in sd,
    }


# ---------------------------------------------------------------------------------------------------------------------
# stages


def patch_embed(sd: dict, img_bchw: torch.Tensor):
    """PatchEmbed.forward - v2_depthanything/patch_embed.py:77-99"""
    w, b = sd["pretrained.patch_embed.proj.weight"], sd["pretrained.patch_embed.proj.bias"]
    p = w.shape[-1]
    x = F.conv2d(img_bchw, w, b, stride=p)
    grid_hw = tuple(x.shape[2:])
    return x.flatten(2).transpose(1, 2), grid_hw


def position_table(sd: dict, grid_hw) -> tuple[torch.Tensor, torch.Tensor]:
    """PositionEncoder - components/position_encoder.py:55-76,108-143 (bicubic, align_corners=False, fp32)"""
    pos = sd["pretrained.pos_embed"].float()
    cls_pos, patch_pos = pos[:, :1], pos[:, 1:]
    n, c = patch_pos.shape[1:]
    base = int(math.isqrt(n))
    img = patch_pos.reshape(1, base, base, c).permute(0, 3, 1, 2)
    img = F.interpolate(img, size=tuple(grid_hw), mode="bicubic", antialias=False)
    return cls_pos, img.permute(0, 2, 3, 1).reshape(1, -1, c)


def layernorm(x, w, b, eps=1e-6):
    """LayerNormEPS6 - components/misc_helpers.py:190-210"""
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def attention(sd: dict, pre: str, x: torch.Tensor, heads: int, use_sdpa: bool = True):
    """OptimizedAttention.forward / Attention.forward - components/transformer_block.py:154-170 / 105-136"""
    B, N, C = x.shape
    d = C // heads
    qkv = F.linear(x, sd[pre + "attn.qkv.weight"], sd[pre + "attn.qkv.bias"])
    qkv = qkv.reshape(B, N, 3, heads, d).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.unbind(0)
    if use_sdpa:
        o = F.scaled_dot_product_attention(q, k, v)
    else:
        a = (q * d**-0.5) @ k.transpose(-2, -1)
        o = a.softmax(dim=-1) @ v
    o = o.transpose(1, 2).reshape(B, N, C)
    return F.linear(o, sd[pre + "attn.proj.weight"], sd[pre + "attn.proj.bias"])


def mlp(sd: dict, pre: str, x: torch.Tensor):
    """MLP2Layers.forward - components/misc_helpers.py:88-120 (exact-erf GELU); ViT-G checkpoints carry the SwiGLU FFN
    instead (SwiGLU.forward - components/misc_helpers.py:169-185: silu(first half) * second half of one doubled Linear)"""
    if pre + "mlp.w12.weight" in sd:
        inner = F.linear(x, sd[pre + "mlp.w12.weight"], sd[pre + "mlp.w12.bias"])
        gate, lin = inner.chunk(2, dim=-1)
        return F.linear(F.silu(gate) * lin, sd[pre + "mlp.w3.weight"], sd[pre + "mlp.w3.bias"])
    h = F.gelu(F.linear(x, sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"]))
    return F.linear(h, sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"])


def block(sd: dict, i: int, x: torch.Tensor, heads: int, use_sdpa: bool = True):
    """TransformerBlock.forward - components/transformer_block.py:53-65"""
    pre = f"pretrained.blocks.{i}."
    a = attention(sd, pre, layernorm(x, sd[pre + "norm1.weight"], sd[pre + "norm1.bias"]), heads, use_sdpa)
    x = x + sd[pre + "ls1.gamma"] * a
    m = mlp(sd, pre, layernorm(x, sd[pre + "norm2.weight"], sd[pre + "norm2.bias"]))
    return x + sd[pre + "ls2.gamma"] * m


def encoder(sd: dict, cfg: dict, tokens: torch.Tensor, grid_hw, use_sdpa: bool = True):
    """DinoV2Model4Stages.forward - v2_depthanything/image_encoder_model.py:80-94 (taps after uniform stages)"""
    cls_pos, patch_pos = position_table(sd, grid_hw)
    cls = (sd["pretrained.cls_token"] + cls_pos).to(tokens.dtype)
    x = torch.cat((cls.expand(tokens.shape[0], -1, -1), tokens + patch_pos.to(tokens.dtype)), dim=1)
    per_stage = int(round(cfg["num_blocks"] / 4))
    last4 = bool(cfg.get("taps_last4", False))  # Depth-Anything V1: v1_depthanything/image_encoder_model.py:92-103
    taps = []
    for i in range(cfg["num_blocks"]):
        x = block(sd, i, x, cfg["num_heads"], use_sdpa)
        if (i >= cfg["num_blocks"] - 4) if last4 else ((i + 1) % per_stage == 0):
            taps.append(x)
    nw, nb = sd["pretrained.norm.weight"], sd["pretrained.norm.bias"]
    return tuple(layernorm(t, nw, nb) for t in taps[:4])


def reassemble(sd: dict, taps, grid_hw):
    """ReassembleModel.forward - v2_depthanything/reassembly_model.py:61-94,139-149"""
    outs = []
    for k, t in enumerate(taps):
        x = t[:, 1:, :].transpose(1, 2).unflatten(2, tuple(grid_hw))  # :142, :208-211
        x = F.conv2d(x, sd[f"depth_head.projects.{k}.weight"], sd[f"depth_head.projects.{k}.bias"])
        if k in (0, 1):  # ConvTranspose2d k = s = 4 / 2 (:262-269)
            s = 4 if k == 0 else 2
            x = F.conv_transpose2d(
                x, sd[f"depth_head.resize_layers.{k}.weight"], sd[f"depth_head.resize_layers.{k}.bias"], stride=s
            )
        elif k == 3:  # Conv2d k=3 s=2 p=1 (:302-309)
            x = F.conv2d(
                x, sd["depth_head.resize_layers.3.weight"], sd["depth_head.resize_layers.3.bias"], stride=2, padding=1
            )
        x = F.conv2d(x, sd[f"depth_head.scratch.layer{k + 1}_rn.weight"], None, padding=1)  # fuse_proj, no bias
        outs.append(x)
    return tuple(outs)


def _rcu(sd: dict, pre: str, x: torch.Tensor):
    """ResidualConv2D.forward - v2_depthanything/fusion_model.py:187-220"""
    y = F.conv2d(F.relu(x), sd[pre + "conv1.weight"], sd[pre + "conv1.bias"], padding=1)
    y = F.conv2d(F.relu(y), sd[pre + "conv2.weight"], sd[pre + "conv2.bias"], padding=1)
    return y + x


def _upsample_project(sd: dict, rn: str, x: torch.Tensor):
    """UpsampleProjectionBlock - fusion_model.py:159-182"""
    x = _rcu(sd, rn + "resConfUnit2.", x)
    x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
    return F.conv2d(x, sd[rn + "out_conv.weight"], sd[rn + "out_conv.bias"])


def fusion(sd: dict, r1, r2, r3, r4):
    """FusionModel.forward - v2_depthanything/fusion_model.py:55-80 (refinenet4.resConfUnit1 unused)"""
    pre = "depth_head.scratch.refinenet"
    f = _upsample_project(sd, pre + "4.", r4)
    for idx, r in ((3, r3), (2, r2), (1, r1)):
        rn = f"{pre}{idx}."
        f = _upsample_project(sd, rn, _rcu(sd, rn + "resConfUnit1.", r) + f)
    return f


def head(sd: dict, cfg: dict, x: torch.Tensor):
    """MonocularDepthHead.forward - v2_depthanything/head_model.py:61-106"""
    p = "depth_head.scratch."
    x = F.conv2d(x, sd[p + "output_conv1.weight"], sd[p + "output_conv1.bias"], padding=1)
    x = F.interpolate(x, scale_factor=cfg["patch_size_px"] / 8, mode="bilinear", align_corners=True)
    x = F.relu(F.conv2d(x, sd[p + "output_conv2.0.weight"], sd[p + "output_conv2.0.bias"], padding=1))
    x = F.conv2d(x, sd[p + "output_conv2.2.weight"], sd[p + "output_conv2.2.bias"])
    x = torch.sigmoid(x) if cfg.get("is_metric", False) else F.relu(x)
    return x.squeeze(1)


def forward(sd: dict, img_bchw: torch.Tensor, cfg: dict | None = None, return_stages: bool = False, use_sdpa=True):
    """DPTModel.forward - muggled_dpt/dpt_model.py:61-83"""
    cfg = cfg or infer_config(sd)
    with torch.inference_mode():
        tokens, grid_hw = patch_embed(sd, img_bchw)
        taps = encoder(sd, cfg, tokens, grid_hw, use_sdpa)
        maps = reassemble(sd, taps, grid_hw)
        fused = fusion(sd, *maps)
        depth = head(sd, cfg, fused)
    if return_stages:
        return {"tokens": tokens, "taps": taps, "maps": maps, "fused": fused, "depth": depth, "grid_hw": grid_hw}
    return depth


# ---------------------------------------------------------------------------------------------------------------------
# pre / post-processing around the path

NORMALISATION = {  # v2_depthanything/patch_embed.py:38-39 ; v31_beit/patch_embed.py:38-39 ; v31_swinv2/patch_embed.py:39-40
    "depthanythingv2": ((0.485, 0.456, 0.406), (0.229, 0.224, 0.225)),
    "depthanythingv1": ((0.485, 0.456, 0.406), (0.229, 0.224, 0.225)),
    "beit": ((0.5, 0.5, 0.5), (0.5, 0.5, 0.5)),
    "swinv2": ((0.5, 0.5, 0.5), (0.5, 0.5, 0.5)),
}


def prepare_image(image_bgr, patch: int, base_grid: int, model_type: str = "depthanythingv2", max_side_length=None,
                  use_square_sizing: bool = True) -> torch.Tensor:
    """PatchEmbed.prepare_image - v2_depthanything/patch_embed.py:103-145, fp32 CPU (image_bgr: HxWx3 uint8 ndarray)"""
    import numpy as np

    tiling = round((8 if model_type == "swinv2" else 2) * patch)
    if max_side_length is None:
        max_side_length = base_grid * patch
    img_h, img_w = image_bgr.shape[0:2]
    largest = max(img_h, img_w)
    scale = max_side_length / largest
    targ_hw = (largest, largest) if use_square_sizing else (img_h, img_w)
    scaled_hw = [max(1, round(side * scale / tiling)) * tiling for side in targ_hw]
    rgb = np.ascontiguousarray(image_bgr[:, :, ::-1])
    chw = torch.tensor(np.transpose(rgb, (2, 0, 1)), dtype=torch.float32)
    bchw = F.interpolate(chw.unsqueeze(0), size=scaled_hw, align_corners=False, antialias=True, mode="bilinear")
    mean, std = NORMALISATION[model_type]
    mean = torch.tensor(mean).view(-1, 1, 1)
    inv_std = 1.0 / torch.tensor(std).view(-1, 1, 1)
    return ((bchw / 255.0) - mean) * inv_std


def postprocess_u8(prediction_bhw: torch.Tensor, target_wh) -> torch.Tensor:
    """convert_to_uint8(scale_prediction(prediction, target_wh)) - demo_helpers/postprocess.py:22-31,75-102"""
    target_hw = (int(target_wh[1]), int(target_wh[0]))
    scaled = F.interpolate(prediction_bhw.unsqueeze(1), size=target_hw, mode="bilinear").squeeze(1)
    lo, hi = scaled.min(), scaled.max()
    return (255.0 * ((scaled - lo) / (hi - lo))).byte()


# ---------------------------------------------------------------------------------------------------------------------
# synthetic upstream-format checkpoints (key schema + weight distributions that give O(1) stages)

STANDARD_CONFIGS = {
    # make_depthanythingv2_dpt.py:88-122
    "vits": dict(F=384, blocks=12, reasm=(48, 96, 192, 384), C=64),
    "vitb": dict(F=768, blocks=12, reasm=(96, 192, 384, 768), C=128),
    "vitl": dict(F=1024, blocks=24, reasm=(256, 512, 1024, 1024), C=256),
    # vit-giant dimensions (make_depthanythingv2_dpt.py:88-95); pass the result through giantify() for its SwiGLU FFN
    "vitg": dict(F=1536, blocks=40, reasm=(1536, 1536, 1536, 1536), C=384),
    # ViT-L widths with 4 blocks: the small-batch tile choices of the 1024-wide GEMMs at oracle-friendly cost
    "vitl_4blk": dict(F=1024, blocks=4, reasm=(256, 512, 1024, 1024), C=256),
    # not a real model: small enough to commit its weights-free fixtures and run anywhere in milliseconds
    "tiny": dict(F=128, blocks=4, reasm=(16, 32, 64, 128), C=32),
    # 8 blocks: the V1 tap rule (last four blocks) and the V2 rule (every second block) differ
    "tiny8": dict(F=128, blocks=8, reasm=(16, 32, 64, 128), C=32),
    # wide reassembly channels on a small encoder: the ViT-L-sized ConvTranspose path (256 channels) at test cost
    "tiny_r256": dict(F=128, blocks=4, reasm=(256, 256, 64, 128), C=32),
}


def make_synthetic_state_dict(name: str = "vits", seed: int = 0, base_grid: int = 37, patch: int = 14) -> dict:
    cfg = STANDARD_CONFIGS[name]
    Fdim, L, R, C = cfg["F"], cfg["blocks"], cfg["reasm"], cfg["C"]
    g = torch.Generator().manual_seed(seed)

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g) * std

    def fan(*shape):  # fan-in scaled normal
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        return rn(*shape, std=fan_in**-0.5)

    sd = {}
    sd["pretrained.cls_token"] = rn(1, 1, Fdim, std=0.5)
    sd["pretrained.pos_embed"] = rn(1, 1 + base_grid * base_grid, Fdim, std=0.5)
    sd["pretrained.mask_token"] = rn(1, Fdim)
    sd["pretrained.patch_embed.proj.weight"] = fan(Fdim, 3, patch, patch)
    sd["pretrained.patch_embed.proj.bias"] = rn(Fdim, std=0.1)
    for i in range(L):
        p = f"pretrained.blocks.{i}."
        for nrm in ("norm1", "norm2"):
            sd[p + nrm + ".weight"] = 1.0 + rn(Fdim, std=0.1)
            sd[p + nrm + ".bias"] = rn(Fdim, std=0.1)
        sd[p + "attn.qkv.weight"] = fan(3 * Fdim, Fdim) * 1.5
        sd[p + "attn.qkv.bias"] = rn(3 * Fdim, std=0.1)
        sd[p + "attn.proj.weight"] = fan(Fdim, Fdim)
        sd[p + "attn.proj.bias"] = rn(Fdim, std=0.1)
        sd[p + "ls1.gamma"] = 0.05 + 0.95 * torch.rand(Fdim, generator=g)
        sd[p + "ls2.gamma"] = 0.05 + 0.95 * torch.rand(Fdim, generator=g)
        sd[p + "mlp.fc1.weight"] = fan(4 * Fdim, Fdim)
        sd[p + "mlp.fc1.bias"] = rn(4 * Fdim, std=0.1)
        sd[p + "mlp.fc2.weight"] = fan(Fdim, 4 * Fdim)
        sd[p + "mlp.fc2.bias"] = rn(Fdim, std=0.1)
    sd["pretrained.norm.weight"] = 1.0 + rn(Fdim, std=0.1)
    sd["pretrained.norm.bias"] = rn(Fdim, std=0.1)
    for k in range(4):
        sd[f"depth_head.projects.{k}.weight"] = fan(R[k], Fdim, 1, 1)
        sd[f"depth_head.projects.{k}.bias"] = rn(R[k], std=0.1)
    sd["depth_head.resize_layers.0.weight"] = rn(R[0], R[0], 4, 4, std=R[0] ** -0.5)
    sd["depth_head.resize_layers.0.bias"] = rn(R[0], std=0.1)
    sd["depth_head.resize_layers.1.weight"] = rn(R[1], R[1], 2, 2, std=R[1] ** -0.5)
    sd["depth_head.resize_layers.1.bias"] = rn(R[1], std=0.1)
    sd["depth_head.resize_layers.3.weight"] = fan(R[3], R[3], 3, 3)
    sd["depth_head.resize_layers.3.bias"] = rn(R[3], std=0.1)
    for k in range(4):
        sd[f"depth_head.scratch.layer{k + 1}_rn.weight"] = fan(C, R[k], 3, 3)
    for i in (1, 2, 3, 4):
        for u in (1, 2):
            for cv in (1, 2):
                p = f"depth_head.scratch.refinenet{i}.resConfUnit{u}.conv{cv}."
                sd[p + "weight"] = fan(C, C, 3, 3) * (1.4 if cv == 1 else 0.7)
                sd[p + "bias"] = rn(C, std=0.1)
        sd[f"depth_head.scratch.refinenet{i}.out_conv.weight"] = fan(C, C, 1, 1)
        sd[f"depth_head.scratch.refinenet{i}.out_conv.bias"] = rn(C, std=0.1)
    sd["depth_head.scratch.output_conv1.weight"] = fan(C // 2, C, 3, 3)
    sd["depth_head.scratch.output_conv1.bias"] = rn(C // 2, std=0.1)
    sd["depth_head.scratch.output_conv2.0.weight"] = fan(32, C // 2, 3, 3) * 1.4
    sd["depth_head.scratch.output_conv2.0.bias"] = rn(32, std=0.1) + 0.2
    sd["depth_head.scratch.output_conv2.2.weight"] = fan(1, 32, 1, 1)
    sd["depth_head.scratch.output_conv2.2.bias"] = torch.full((1,), 2.0)
    return sd


def swiglu_hidden_features(features: int, ratio: float = 4) -> int:
    """SwiGLU.__init__ - components/misc_helpers.py:161-163"""
    return 8 * ((int(int(ratio * features) * 2 / 3) + 7) // 8)


def giantify(sd: dict, seed: int = 0) -> dict:
    """Turns a synthetic Depth-Anything checkpoint into the ViT-G schema: every block's fc1 / fc2 pair is replaced by
    the SwiGLU FFN tensors mlp.w12 [2h, F] / mlp.w3 [F, h] (config_from_original_state_dict.py:248-259 keys off
    pretrained.blocks.0.mlp.w12.weight). Separate RNG stream: the base checkpoint's tensors keep their values."""
    g = torch.Generator().manual_seed(1000 + seed)
    out = {k: v for k, v in sd.items() if ".mlp.fc" not in k}
    feats = int(sd["pretrained.patch_embed.proj.weight"].shape[0])
    h = swiglu_hidden_features(feats)
    blocks = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("pretrained.blocks."))
    for i in range(blocks):
        p = f"pretrained.blocks.{i}.mlp."
        out[p + "w12.weight"] = torch.randn(2 * h, feats, generator=g) * feats**-0.5 * 1.5
        out[p + "w12.bias"] = torch.randn(2 * h, generator=g) * 0.1
        out[p + "w3.weight"] = torch.randn(feats, h, generator=g) * h**-0.5
        out[p + "w3.bias"] = torch.randn(feats, generator=g) * 0.1
    return out


def make_input(B: int, H: int, W: int, seed: int = 0) -> torch.Tensor:
    g = torch.Generator().manual_seed(1000 + seed)
    return torch.randn(B, 3, H, W, generator=g)


# =====================================================================================================================
# MiDaS v3.1 BEiT: upstream keys `pretrained.model.*`, `pretrained.act_postprocess*`,
# `scratch.*`
# =====================================================================================================================


def infer_config_beit(sd: dict) -> dict:
    """v31_beit/state_dict_conversion/config_from_midas_state_dict.py (heads from the bias-table width :67-80, base grid
    from its length :205-246)"""
    feats = int(sd["pretrained.model.patch_embed.proj.weight"].shape[0])
    patch = int(sd["pretrained.model.patch_embed.proj.weight"].shape[3])
    blocks = 1 + max(int(k.split(".")[3]) for k in sd if k.startswith("pretrained.model.blocks."))
    table = sd["pretrained.model.blocks.0.attn.relative_position_bias_table"]
    heads = int(table.shape[1])
    side = int(math.isqrt(int(table.shape[0]) - 3))  # (2g-1)
    base = (side + 1) // 2
    return {
        "features_per_token": feats,
        "num_blocks": blocks,
        "num_heads": heads,
        "reassembly_features_list": [int(sd[f"scratch.layer{i}_rn.weight"].shape[1]) for i in (1, 2, 3, 4)],
        "fusion_channels": int(sd["scratch.layer1_rn.weight"].shape[0]),
        "patch_size_px": patch,
        "base_patch_grid_hw": (base, base),
    }


def beit_relative_position_index(grid_hw) -> torch.Tensor:
    """RelativePositionEncoding._generate_relative_position_index - v31_beit/components/relative_positional_encoder.py:117-238"""
    gh, gw = grid_hw
    n = gh * gw + 1
    ys, xs = torch.meshgrid(torch.arange(gh), torch.arange(gw), indexing="ij")
    coords = torch.stack((ys.flatten(), xs.flatten()))  # [2, g]
    rel = coords[:, :, None] - coords[:, None, :]
    idx_tok = (rel[0] + gh - 1) * (2 * gw - 1) + (rel[1] + gw - 1)
    max_idx = (2 * gh - 1) * (2 * gw - 1) - 1
    idx = torch.zeros((n, n), dtype=torch.long)
    idx[1:, 1:] = idx_tok
    idx[0, :] = max_idx + 1  # cls -> token
    idx[:, 0] = max_idx + 2  # token -> cls
    idx[0, 0] = max_idx + 3  # cls -> cls
    return idx


def beit_position_bias(table: torch.Tensor, base_hw, grid_hw) -> torch.Tensor:
    """_generate_position_bias_lut - relative_positional_encoder.py:242-309 (bilinear, align_corners=False) -> [H,N,N]"""
    heads = table.shape[1]
    rh, rw = 2 * base_hw[0] - 1, 2 * base_hw[1] - 1
    nh, nw = 2 * grid_hw[0] - 1, 2 * grid_hw[1] - 1
    tok, cls = table[: rh * rw], table[rh * rw:]
    t2d = tok.reshape(1, rh, rw, heads).permute(0, 3, 1, 2)
    t2d = F.interpolate(t2d, size=(nh, nw), mode="bilinear")
    lut = torch.cat([t2d.permute(0, 2, 3, 1).reshape(nh * nw, heads), cls])
    n = grid_hw[0] * grid_hw[1] + 1
    return lut[beit_relative_position_index(grid_hw).reshape(-1)].reshape(n, n, heads).permute(2, 0, 1).contiguous()


def beit_block(sd: dict, i: int, x: torch.Tensor, heads: int, base_hw, grid_hw):
    """TransformerBlock / SelfAttentionRelPos - v31_beit/image_encoder_model.py:233-251,311-356"""
    pre = f"pretrained.model.blocks.{i}."
    B, N, C = x.shape
    d = C // heads
    t = F.layer_norm(x, (C,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], 1e-6)
    qkv = F.linear(t, sd[pre + "attn.qkv.weight"]).reshape(B, N, 3, heads, d).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.unbind(0)
    q = (q + sd[pre + "attn.q_bias"].reshape(1, heads, 1, d)) * d**-0.5
    v = v + sd[pre + "attn.v_bias"].reshape(1, heads, 1, d)
    a = q @ k.transpose(-2, -1) + beit_position_bias(sd[pre + "attn.relative_position_bias_table"], base_hw, grid_hw)
    o = (a.softmax(dim=-1) @ v).transpose(1, 2).reshape(B, N, C)
    o = F.linear(o, sd[pre + "attn.proj.weight"], sd[pre + "attn.proj.bias"])
    x = x + sd[pre + "gamma_1"] * o
    t = F.layer_norm(x, (C,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], 1e-6)
    m = F.linear(F.gelu(F.linear(t, sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"])),
                 sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"])
    return x + sd[pre + "gamma_2"] * m


def beit_encoder(sd: dict, cfg: dict, tokens: torch.Tensor, grid_hw):
    """BEiTModel4Stage.forward - v31_beit/image_encoder_model.py:68-91 (no position embedding, no output norm)"""
    x = torch.cat((sd["pretrained.model.cls_token"].expand(tokens.shape[0], -1, -1), tokens), dim=1)
    per_stage = int(round(cfg["num_blocks"] / 4))
    taps = []
    for i in range(cfg["num_blocks"]):
        x = beit_block(sd, i, x, cfg["num_heads"], cfg["base_patch_grid_hw"], grid_hw)
        if (i + 1) % per_stage == 0:
            taps.append(x)
    return tuple(taps[:4])


def beit_reassemble(sd: dict, taps, grid_hw):
    """ReassembleBlock.forward - v31_beit/reassembly_model.py:118-128 + ReadoutProjectLayer (readout_projection.py)"""
    outs = []
    for k, t in enumerate(taps):
        p = f"pretrained.act_postprocess{k + 1}."
        cat = torch.cat((t[:, 1:], t[:, :1].expand(-1, t.shape[1] - 1, -1)), dim=-1)  # [patch, cls] (:74-79)
        x = F.gelu(F.linear(cat, sd[p + "0.project.0.weight"], sd[p + "0.project.0.bias"]))
        x = x.transpose(1, 2).unflatten(2, tuple(grid_hw))
        x = F.conv2d(x, sd[p + "3.weight"], sd[p + "3.bias"])
        if k in (0, 1):
            x = F.conv_transpose2d(x, sd[p + "4.weight"], sd[p + "4.bias"], stride=4 if k == 0 else 2)
        elif k == 3:
            x = F.conv2d(x, sd[p + "4.weight"], sd[p + "4.bias"], stride=2, padding=1)
        outs.append(F.conv2d(x, sd[f"scratch.layer{k + 1}_rn.weight"], None, padding=1))
    return tuple(outs)


def _midas_fusion(sd: dict, r1, r2, r3, r4):
    """FusionModel.forward - v31_beit/fusion_model.py (same structure as the Depth-Anything one, `scratch.` prefix)"""
    def rcu(pre, x):
        y = F.conv2d(F.relu(x), sd[pre + "conv1.weight"], sd[pre + "conv1.bias"], padding=1)
        y = F.conv2d(F.relu(y), sd[pre + "conv2.weight"], sd[pre + "conv2.bias"], padding=1)
        return y + x

    def up(rn, x):
        x = rcu(rn + "resConfUnit2.", x)
        x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
        return F.conv2d(x, sd[rn + "out_conv.weight"], sd[rn + "out_conv.bias"])

    f = up("scratch.refinenet4.", r4)
    for idx, r in ((3, r3), (2, r2), (1, r1)):
        rn = f"scratch.refinenet{idx}."
        f = up(rn, rcu(rn + "resConfUnit1.", r) + f)
    return f


def _midas_head(sd: dict, x: torch.Tensor):
    """MonocularDepthHead.forward - v31_beit/head_model.py:37-74 (x2 upsample)"""
    x = F.conv2d(x, sd["scratch.output_conv.0.weight"], sd["scratch.output_conv.0.bias"], padding=1)
    x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
    x = F.relu(F.conv2d(x, sd["scratch.output_conv.2.weight"], sd["scratch.output_conv.2.bias"], padding=1))
    x = F.relu(F.conv2d(x, sd["scratch.output_conv.4.weight"], sd["scratch.output_conv.4.bias"]))
    return x.squeeze(1)


def forward_beit(sd: dict, img_bchw: torch.Tensor, cfg: dict | None = None, return_stages: bool = False):
    cfg = cfg or infer_config_beit(sd)
    with torch.inference_mode():
        w, b = sd["pretrained.model.patch_embed.proj.weight"], sd["pretrained.model.patch_embed.proj.bias"]
        x = F.conv2d(img_bchw, w, b, stride=w.shape[-1])  # v31_beit/patch_embed.py:74-91
        grid_hw = tuple(x.shape[2:])
        tokens = x.flatten(2).transpose(1, 2)
        taps = beit_encoder(sd, cfg, tokens, grid_hw)
        maps = beit_reassemble(sd, taps, grid_hw)
        fused = _midas_fusion(sd, *maps)
        depth = _midas_head(sd, fused)
    if return_stages:
        return {"tokens": tokens, "taps": taps, "maps": maps, "fused": fused, "depth": depth, "grid_hw": grid_hw}
    return depth


BEIT_CONFIGS = {
    # make_beit_dpt.py docstring
    "beit_large_384": dict(F=1024, heads=16, blocks=24, reasm=(256, 512, 1024, 1024), C=256, base=24),
    "beit_base_384": dict(F=768, heads=12, blocks=12, reasm=(96, 192, 384, 768), C=256, base=24),
    "beit_tiny": dict(F=128, heads=2, blocks=4, reasm=(16, 32, 64, 128), C=32, base=6),
}


def make_synthetic_state_dict_beit(name: str = "beit_tiny", seed: int = 0) -> dict:
    """MiDaS v3.1 BEiT key schema, fan-in scaled weights"""
    cfg = BEIT_CONFIGS[name]
    Fd, H, L, R, C, g0 = cfg["F"], cfg["heads"], cfg["blocks"], cfg["reasm"], cfg["C"], cfg["base"]
    g = torch.Generator().manual_seed(seed)

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g) * std

    def fan(*shape):
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        return rn(*shape, std=fan_in**-0.5)

    sd = {}
    sd["pretrained.model.cls_token"] = rn(1, 1, Fd, std=0.5)
    sd["pretrained.model.patch_embed.proj.weight"] = fan(Fd, 3, 16, 16)
    sd["pretrained.model.patch_embed.proj.bias"] = rn(Fd, std=0.1)
    for i in range(L):
        p = f"pretrained.model.blocks.{i}."
        sd[p + "gamma_1"] = 0.05 + 0.95 * torch.rand(Fd, generator=g)
        sd[p + "gamma_2"] = 0.05 + 0.95 * torch.rand(Fd, generator=g)
        for nrm in ("norm1", "norm2"):
            sd[p + nrm + ".weight"] = 1.0 + rn(Fd, std=0.1)
            sd[p + nrm + ".bias"] = rn(Fd, std=0.1)
        sd[p + "attn.q_bias"] = rn(Fd, std=0.2)
        sd[p + "attn.v_bias"] = rn(Fd, std=0.2)
        sd[p + "attn.qkv.weight"] = fan(3 * Fd, Fd) * 1.5
        sd[p + "attn.relative_position_bias_table"] = rn((2 * g0 - 1) ** 2 + 3, H, std=1.0)
        sd[p + "attn.relative_position_index"] = torch.zeros(g0 * g0 + 1, g0 * g0 + 1, dtype=torch.long)  # dropped
        sd[p + "attn.proj.weight"] = fan(Fd, Fd)
        sd[p + "attn.proj.bias"] = rn(Fd, std=0.1)
        sd[p + "mlp.fc1.weight"] = fan(4 * Fd, Fd)
        sd[p + "mlp.fc1.bias"] = rn(4 * Fd, std=0.1)
        sd[p + "mlp.fc2.weight"] = fan(Fd, 4 * Fd)
        sd[p + "mlp.fc2.bias"] = rn(Fd, std=0.1)
    for k in range(4):
        p = f"pretrained.act_postprocess{k + 1}."
        sd[p + "0.project.0.weight"] = fan(Fd, 2 * Fd) * 1.4
        sd[p + "0.project.0.bias"] = rn(Fd, std=0.1)
        sd[p + "3.weight"] = fan(R[k], Fd, 1, 1) * 1.5
        sd[p + "3.bias"] = rn(R[k], std=0.1)
    sd["pretrained.act_postprocess1.4.weight"] = rn(R[0], R[0], 4, 4, std=R[0] ** -0.5)
    sd["pretrained.act_postprocess1.4.bias"] = rn(R[0], std=0.1)
    sd["pretrained.act_postprocess2.4.weight"] = rn(R[1], R[1], 2, 2, std=R[1] ** -0.5)
    sd["pretrained.act_postprocess2.4.bias"] = rn(R[1], std=0.1)
    sd["pretrained.act_postprocess4.4.weight"] = fan(R[3], R[3], 3, 3)
    sd["pretrained.act_postprocess4.4.bias"] = rn(R[3], std=0.1)
    for k in range(4):
        sd[f"scratch.layer{k + 1}_rn.weight"] = fan(C, R[k], 3, 3)
    for i in (1, 2, 3, 4):
        for u in (1, 2):
            for cv in (1, 2):
                p = f"scratch.refinenet{i}.resConfUnit{u}.conv{cv}."
                sd[p + "weight"] = fan(C, C, 3, 3) * (1.4 if cv == 1 else 0.7)
                sd[p + "bias"] = rn(C, std=0.1)
        sd[f"scratch.refinenet{i}.out_conv.weight"] = fan(C, C, 1, 1)
        sd[f"scratch.refinenet{i}.out_conv.bias"] = rn(C, std=0.1)
    sd["scratch.output_conv.0.weight"] = fan(C // 2, C, 3, 3)
    sd["scratch.output_conv.0.bias"] = rn(C // 2, std=0.1)
    sd["scratch.output_conv.2.weight"] = fan(32, C // 2, 3, 3) * 1.4
    sd["scratch.output_conv.2.bias"] = rn(32, std=0.1) + 0.2
    sd["scratch.output_conv.4.weight"] = fan(1, 32, 1, 1)
    sd["scratch.output_conv.4.bias"] = torch.full((1,), 2.0)
    return sd


# =====================================================================================================================
# MiDaS v3.1 SwinV2: hierarchical windowed cosine attention, post-norm blocks
# =====================================================================================================================


def infer_config_swinv2(sd: dict) -> dict:
    """v31_swinv2/state_dict_conversion/config_from_midas_state_dict.py:17-214"""
    f0 = int(sd["pretrained.model.patch_embed.proj.weight"].shape[0])
    patch = int(sd["pretrained.model.patch_embed.proj.weight"].shape[3])
    heads, layers = {}, {}
    for k in sd:
        if k.startswith("pretrained.model.layers.") and ".blocks." in k:
            s, b = int(k.split(".")[3]), int(k.split(".")[5])
            layers[s] = max(layers.get(s, 0), b + 1)
            if k.endswith("logit_scale"):
                heads[s] = int(sd[k].shape[0])
    mask_key = sorted(k for k in sd if k.endswith("attn_mask"))[0]
    nw, area = sd[mask_key].shape[:2]
    win = int(math.isqrt(int(area)))
    base = int(math.isqrt(int(nw) * int(area)))
    pre_lut = {16: [16, 16, 16, 8], 24: [12, 12, 12, 6]}
    return {
        "features_per_stage": [f0 * 2**i for i in range(4)],
        "heads_per_stage": [heads[i] for i in range(4)],
        "layers_per_stage": [layers[i] for i in range(4)],
        "base_patch_grid_hw": (base, base),
        "window_size_hw": (win, win),
        "pretrained_window_sizes_per_stage": pre_lut.get(win, [None] * 4),
        "fusion_channels": int(sd["scratch.layer1_rn.weight"].shape[0]),
        "patch_size_px": patch,
    }


def swin_window_and_shift(grid_hw, target_hw):
    """adjust_window_and_shift_sizes - v31_swinv2/components/windowed_attention.py:345-388"""
    out_win, out_shift = [], []
    for patch, targ in zip(grid_hw, target_hw):
        win = min(targ, patch)
        if patch % win != 0:
            divs = [d for d in range(win // 2, 2 * win) if patch % d == 0]
            win = min(divs, key=lambda d: abs(patch - d))
        out_win.append(win)
        out_shift.append(0 if patch <= win else win // 2)
    return tuple(out_win), tuple(out_shift)


def swin_shift_mask(grid_hw, win_hw, shift_hw):
    """make_shift_mask - windowed_attention.py:394-439 -> [nW, A, A] of 0 / -100 (None when no shift)"""
    (gh, gw), (wh, ww), (sh, sw) = grid_hw, win_hw, shift_hw
    if sh == 0 and sw == 0:
        return None
    img = torch.zeros((gh, gw))
    cnt = 0
    for hs in (slice(0, -wh), slice(-wh, -sh), slice(-sh, None)):
        for ws in (slice(0, -ww), slice(-ww, -sw), slice(-sw, None)):
            img[hs, ws] = cnt
            cnt += 1
    win = img.reshape(gh // wh, wh, gw // ww, ww).permute(0, 2, 1, 3).reshape(-1, wh * ww)
    diff = win.unsqueeze(1) - win.unsqueeze(2)
    return torch.where(diff != 0, torch.tensor(-100.0), torch.tensor(0.0))


def swin_cpb_bias(sd: dict, pre: str, win_hw, pretrained_window):
    """RelativePositionEncoding._get_position_bias - v31_swinv2/components/relative_positional_encoder.py:60-93,121-283"""
    wh, ww = win_hw
    ys = torch.arange(-(wh - 1), wh, dtype=torch.float32)
    xs = torch.arange(-(ww - 1), ww, dtype=torch.float32)
    table = torch.stack(torch.meshgrid([ys, xs], indexing="ij")).permute(1, 2, 0).contiguous()
    table[:, :, 0] /= max((wh if pretrained_window is None else pretrained_window) - 1, 1)
    table[:, :, 1] /= max((ww if pretrained_window is None else pretrained_window) - 1, 1)
    table = torch.sign(table) * torch.log2(torch.abs(table * 8) + 1.0) / math.log2(8)
    h = F.relu(F.linear(table.reshape(-1, 2), sd[pre + "cpb_mlp.0.weight"], sd[pre + "cpb_mlp.0.bias"]))
    bias_table = F.linear(h, sd[pre + "cpb_mlp.2.weight"])  # [(2wh-1)(2ww-1), H]
    cy, cx = torch.meshgrid(torch.arange(wh), torch.arange(ww), indexing="ij")
    coords = torch.stack((cy.flatten(), cx.flatten()))
    rel = coords[:, :, None] - coords[:, None, :]
    idx = (rel[0] + wh - 1) * (2 * ww - 1) + (rel[1] + ww - 1)
    area = wh * ww
    bias = 16 * torch.sigmoid(bias_table[idx.reshape(-1)])
    return bias.reshape(area, area, -1).permute(2, 0, 1).contiguous()  # [H, A, A]


def swin_block(sd: dict, pre: str, x: torch.Tensor, grid_hw, heads: int, target_win, pretrained_window, is_shift_block):
    """SwinTransformerBlock.forward (post-norm) - v31_swinv2/image_encoder_model.py:213-225 and
    WindowAttentionWithRelPos - components/windowed_attention.py:65-123"""
    B, N, C = x.shape
    gh, gw = grid_hw
    (wh, ww), (sh, sw) = swin_window_and_shift(grid_hw, target_win)
    need_shift = is_shift_block and (sh > 0 or sw > 0)
    img = x.reshape(B, gh, gw, C)
    if need_shift:
        img = torch.roll(img, shifts=(-sh, -sw), dims=(1, 2))
    nwy, nwx = gh // wh, gw // ww
    win = img.reshape(B, nwy, wh, nwx, ww, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, wh * ww, C)
    P, A, _ = win.shape
    d = C // heads
    a = pre + "attn."
    qkv = F.linear(win, sd[a + "qkv.weight"]).reshape(P, A, 3, heads, d).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.unbind(0)
    q = q + sd[a + "q_bias"].reshape(1, heads, 1, d)
    v = v + sd[a + "v_bias"].reshape(1, heads, 1, d)
    att = F.normalize(q, dim=-1) @ F.normalize(k, dim=-1).transpose(-2, -1)
    # logit_scale is clamped (<= ln 100) and exponentiated at load time (convert_midas_state_dict_keys.py:115-131)
    att = att * torch.clamp(sd[a + "logit_scale"], max=math.log(100.0)).exp()
    att = att + swin_cpb_bias(sd, a, (wh, ww), pretrained_window).unsqueeze(0)
    if need_shift:
        mask = swin_shift_mask(grid_hw, (wh, ww), (sh, sw))
        att = att + mask.unsqueeze(1).repeat(B, 1, 1, 1)
    o = (att.softmax(dim=-1) @ v).transpose(1, 2).reshape(P, A, C)
    o = F.linear(o, sd[a + "proj.weight"], sd[a + "proj.bias"])
    img = o.reshape(B, nwy, nwx, wh, ww, C).permute(0, 1, 3, 2, 4, 5).reshape(B, gh, gw, C)
    if need_shift:
        img = torch.roll(img, shifts=(sh, sw), dims=(1, 2))
    t = img.reshape(B, N, C)
    x = x + F.layer_norm(t, (C,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], 1e-5)
    m = F.linear(F.gelu(F.linear(x, sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"])),
                 sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"])
    return x + F.layer_norm(m, (C,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], 1e-5)


def swin_patch_merge(sd: dict, s: int, x: torch.Tensor, grid_hw):
    """PatchMerge.forward - v31_swinv2/components/patch_merge.py:49-103 (concat order TL, BL, TR, BR)"""
    B, N, C = x.shape
    gh, gw = grid_hw
    img = x.reshape(B, gh, gw, C)
    cat = torch.cat([img[:, 0::2, 0::2], img[:, 1::2, 0::2], img[:, 0::2, 1::2], img[:, 1::2, 1::2]], dim=-1)
    t = F.linear(cat.reshape(B, N // 4, 4 * C), sd[f"pretrained.model.layers.{s}.downsample.reduction.weight"])
    p = f"pretrained.model.layers.{s}.downsample.norm."
    return F.layer_norm(t, (t.shape[-1],), sd[p + "weight"], sd[p + "bias"], 1e-5), (gh // 2, gw // 2)


def forward_swinv2(sd: dict, img_bchw: torch.Tensor, cfg: dict | None = None, return_stages: bool = False):
    """DPTModel.forward with the SwinV2 sub-models (make_swinv2_dpt.py:61-139)"""
    cfg = cfg or infer_config_swinv2(sd)
    with torch.inference_mode():
        w, b = sd["pretrained.model.patch_embed.proj.weight"], sd["pretrained.model.patch_embed.proj.bias"]
        x = F.conv2d(img_bchw, w, b, stride=w.shape[-1])  # v31_swinv2/patch_embed.py:76-94 (conv + LayerNorm)
        grid_hw = tuple(x.shape[2:])
        x = x.flatten(2).transpose(1, 2)
        tokens = F.layer_norm(x, (x.shape[-1],), sd["pretrained.model.patch_embed.norm.weight"],
                              sd["pretrained.model.patch_embed.norm.bias"], 1e-5)
        taps, g, x = [], grid_hw, tokens
        for s in range(4):
            if s > 0:
                x, g = swin_patch_merge(sd, s - 1, x, g)
            for bi in range(cfg["layers_per_stage"][s]):
                x = swin_block(sd, f"pretrained.model.layers.{s}.blocks.{bi}.", x, g, cfg["heads_per_stage"][s],
                               cfg["window_size_hw"], cfg["pretrained_window_sizes_per_stage"][s], bi % 2 == 1)
            taps.append(x)
        maps = []
        for k, t in enumerate(taps):  # ReassembleBlock - v31_swinv2/reassembly_model.py:113-122
            gk = (grid_hw[0] // 2**k, grid_hw[1] // 2**k)
            maps.append(F.conv2d(t.transpose(1, 2).unflatten(2, gk), sd[f"scratch.layer{k + 1}_rn.weight"], None, padding=1))
        fused = _midas_fusion(sd, *maps)
        depth = _midas_head(sd, fused)
    if return_stages:
        return {"tokens": tokens, "taps": tuple(taps), "maps": tuple(maps), "fused": fused, "depth": depth, "grid_hw": grid_hw}
    return depth


SWINV2_CONFIGS = {
    # make_swinv2_dpt.py docstring
    "swinv2_large_384": dict(F0=192, heads=(6, 12, 24, 48), layers=(2, 2, 18, 2), base=96, win=24, C=256),
    "swinv2_base_384": dict(F0=128, heads=(4, 8, 16, 32), layers=(2, 2, 18, 2), base=96, win=24, C=256),
    "swinv2_tiny_256": dict(F0=96, heads=(3, 6, 12, 24), layers=(2, 2, 6, 2), base=64, win=16, C=256),
    "swinv2_micro": dict(F0=32, heads=(1, 2, 4, 8), layers=(2, 2, 2, 2), base=32, win=8, C=32),
}


def make_synthetic_state_dict_swinv2(name: str = "swinv2_micro", seed: int = 0, logit_std: float = 1.5) -> dict:
    """MiDaS v3.1 SwinV2 key schema. logit_std spreads the per-head logit scales around ln(10):
    the default reaches the ln(100) clamp (logits up to +-100, a very peaked softmax that amplifies 16-bit rounding of
    the normalised q/k - the reference's own bf16 forward is off by 2e-2..1e-1 on these weights); 0.3 keeps them mild."""
    cfg = SWINV2_CONFIGS[name]
    F0, Hs, Ls, base, win, C = cfg["F0"], cfg["heads"], cfg["layers"], cfg["base"], cfg["win"], cfg["C"]
    g = torch.Generator().manual_seed(seed)

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g) * std

    def fan(*shape):
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        return rn(*shape, std=fan_in**-0.5)

    sd = {}
    sd["pretrained.model.patch_embed.proj.weight"] = fan(F0, 3, 4, 4)
    sd["pretrained.model.patch_embed.proj.bias"] = rn(F0, std=0.1)
    sd["pretrained.model.patch_embed.norm.weight"] = 1.0 + rn(F0, std=0.1)
    sd["pretrained.model.patch_embed.norm.bias"] = rn(F0, std=0.1)
    for s in range(4):
        Fs, H = F0 * 2**s, Hs[s]
        grid = base // 2**s
        for bi in range(Ls[s]):
            p = f"pretrained.model.layers.{s}.blocks.{bi}."
            sd[p + "attn.logit_scale"] = math.log(10.0) + rn(H, 1, 1, std=logit_std)  # default: some above ln 100 (clamp)
            sd[p + "attn.q_bias"] = rn(Fs, std=0.3)
            sd[p + "attn.v_bias"] = rn(Fs, std=0.3)
            sd[p + "attn.qkv.weight"] = fan(3 * Fs, Fs) * 1.5
            sd[p + "attn.cpb_mlp.0.weight"] = rn(512, 2, std=1.0)
            sd[p + "attn.cpb_mlp.0.bias"] = rn(512, std=0.5)
            sd[p + "attn.cpb_mlp.2.weight"] = rn(H, 512, std=0.08)
            sd[p + "attn.proj.weight"] = fan(Fs, Fs)
            sd[p + "attn.proj.bias"] = rn(Fs, std=0.1)
            for nrm in ("norm1", "norm2"):
                sd[p + nrm + ".weight"] = 0.5 + rn(Fs, std=0.1)
                sd[p + nrm + ".bias"] = rn(Fs, std=0.1)
            sd[p + "mlp.fc1.weight"] = fan(4 * Fs, Fs)
            sd[p + "mlp.fc1.bias"] = rn(4 * Fs, std=0.1)
            sd[p + "mlp.fc2.weight"] = fan(Fs, 4 * Fs)
            sd[p + "mlp.fc2.bias"] = rn(Fs, std=0.1)
            w = min(win, grid)
            if bi % 2 == 1 and grid > w:  # the stored masks are dropped by the loader but drive config inference
                sd[p + "attn_mask"] = torch.zeros((grid // w) ** 2, w * w, w * w)
        if s < 3:
            p = f"pretrained.model.layers.{s}.downsample."
            sd[p + "reduction.weight"] = fan(2 * Fs, 4 * Fs)
            sd[p + "norm.weight"] = 1.0 + rn(2 * Fs, std=0.1)
            sd[p + "norm.bias"] = rn(2 * Fs, std=0.1)
    for k in range(4):
        sd[f"scratch.layer{k + 1}_rn.weight"] = fan(C, F0 * 2**k, 3, 3)
    for i in (1, 2, 3, 4):
        for u in (1, 2):
            for cv in (1, 2):
                p = f"scratch.refinenet{i}.resConfUnit{u}.conv{cv}."
                sd[p + "weight"] = fan(C, C, 3, 3) * (1.4 if cv == 1 else 0.7)
                sd[p + "bias"] = rn(C, std=0.1)
        sd[f"scratch.refinenet{i}.out_conv.weight"] = fan(C, C, 1, 1)
        sd[f"scratch.refinenet{i}.out_conv.bias"] = rn(C, std=0.1)
    sd["scratch.output_conv.0.weight"] = fan(C // 2, C, 3, 3)
    sd["scratch.output_conv.0.bias"] = rn(C // 2, std=0.1)
    sd["scratch.output_conv.2.weight"] = fan(32, C // 2, 3, 3) * 1.4
    sd["scratch.output_conv.2.bias"] = rn(32, std=0.1) + 0.2
    sd["scratch.output_conv.4.weight"] = fan(1, 32, 1, 1)
    sd["scratch.output_conv.4.bias"] = torch.full((1,), 2.0)
    return sd
