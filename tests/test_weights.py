"""CPU: weight loader / packer. Every packed layout is checked by evaluating the GEMM formulation the kernels use
(gemm_tc.cuh) in torch and comparing with the reference op (F.conv2d / F.conv_transpose2d / F.linear)."""
import os
import tempfile

import pytest
import torch
import torch.nn.functional as F

from muggled_dpt_b200 import weights as Wt
from oracle import dpt_oracle as O


def _im2col3x3(x_nhwc, kpad):
    B, H, W, C = x_nhwc.shape
    xp = F.pad(x_nhwc, (0, kpad - C, 1, 1, 1, 1))
    cols = [xp[:, ky:ky + H, kx:kx + W, :] for ky in range(3) for kx in range(3)]
    return torch.cat(cols, dim=-1)  # [B,H,W,9*kpad], column = tap*kpad + c


def test_pack_conv_matches_conv2d():
    torch.manual_seed(0)
    x = torch.randn(2, 7, 9, 24)
    w = torch.randn(40, 24, 3, 3)
    packed = Wt.pack_conv(w)
    assert packed.shape == (40, 9 * 64)
    got = _im2col3x3(x, 64) @ packed.t()
    ref = F.conv2d(x.permute(0, 3, 1, 2), w, padding=1).permute(0, 2, 3, 1)
    torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("s", [2, 4])
def test_pack_conv_transpose_matches_conv_transpose2d(s):
    torch.manual_seed(1)
    ci = co = 24
    x = torch.randn(2, 5, 6, ci)
    w = torch.randn(ci, co, s, s)
    b = torch.randn(co)
    packed = Wt.pack_conv_transpose(w)
    cop = 32  # output channels padded to whole 32-column tiles (zero rows), the bias with them
    assert packed.shape == (s * s * cop, 64)
    bp = Wt.pad_conv_transpose_bias(b)
    assert bp.shape == (cop,) and torch.count_nonzero(bp[co:]) == 0
    xp = F.pad(x, (0, 64 - ci))
    out = torch.zeros(2, 5 * s, 6 * s, co)
    for sub in range(s * s):
        ky, kx = sub // s, sub % s
        full = xp @ packed[sub * cop:(sub + 1) * cop].t() + bp
        assert torch.count_nonzero(full[..., co:]) == 0
        out[:, ky::s, kx::s, :] = full[..., :co]
    ref = F.conv_transpose2d(x.permute(0, 3, 1, 2), w, b, stride=s).permute(0, 2, 3, 1)
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-4)


def test_pack_patch_embed_matches_strided_conv():
    torch.manual_seed(2)
    P, Fd = 14, 32
    img = torch.randn(2, 3, 28, 42)
    w = torch.randn(Fd, 3, P, P)
    packed = Wt.pack_patch_embed(w)
    assert packed.shape == (Fd, 640)
    gh, gw = 2, 3
    cols = img.reshape(2, 3, gh, P, gw, P).permute(0, 2, 4, 1, 3, 5).reshape(2, gh * gw, 3 * P * P)
    got = F.pad(cols, (0, 640 - 588)) @ packed.t()
    ref = F.conv2d(img, w, stride=P).flatten(2).transpose(1, 2)
    torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-4)


def test_layerscale_folding():
    torch.manual_seed(3)
    sd = O.make_synthetic_state_dict("tiny", seed=3, base_grid=5)
    cfg = Wt.get_model_config_from_state_dict(sd, False, True)
    packed = Wt.pack_depthanything_v2(sd, cfg)
    x = torch.randn(5, 128)
    ref = sd["pretrained.blocks.1.ls1.gamma"] * F.linear(x, sd["pretrained.blocks.1.attn.proj.weight"],
                                                         sd["pretrained.blocks.1.attn.proj.bias"])
    got = x @ packed["blk1.proj.w"][0].t() + packed["blk1.proj.b"][0]
    torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-5)


def test_config_matches_reference_keys_and_values():
    sd = O.make_synthetic_state_dict("vits", seed=11)
    cfg = Wt.get_model_config_from_state_dict(sd, enable_cache=False, enable_optimizations=True)
    # key set and order of the reference's config dict (config_from_original_state_dict.py:29-41)
    assert list(cfg.keys()) == ["features_per_token", "num_blocks", "num_heads", "reassembly_features_list",
                                "fusion_channels", "patch_size_px", "base_patch_grid_hw", "is_giant", "is_metric",
                                "enable_cache", "enable_optimizations"]
    ocfg = O.infer_config(sd)
    for k, v in ocfg.items():
        assert cfg[k] == v, k


def test_dropped_and_missing_keys():
    sd = O.make_synthetic_state_dict("tiny", seed=3, base_grid=5)
    cfg = Wt.get_model_config_from_state_dict(sd, False, True)
    packed = Wt.pack_depthanything_v2(sd, cfg)
    assert not any(k.startswith("fus3.rcu1") for k in packed)  # refinenet4.resConfUnit1 is dropped
    sd2 = dict(sd)
    del sd2["pretrained.blocks.2.mlp.fc2.bias"]
    with pytest.raises(RuntimeError):
        Wt.pack_depthanything_v2(sd2, cfg, strict=True)
    Wt.pack_depthanything_v2(sd2, cfg, strict=False)


def test_unexpected_keys_follow_the_reference_converter():
    """strict loading: a stray key under a prefix the reference's converter renames is an "Unexpected key(s)" error (it
    would reach load_state_dict(strict=True)); a key the converter ignores, or one on its drop list, is not."""
    sd = O.make_synthetic_state_dict("tiny", seed=3, base_grid=5)
    cfg = Wt.get_model_config_from_state_dict(sd, False, True)
    ignored = dict(sd, **{"something.else": torch.zeros(1), "pretrained.mask_token": torch.zeros(1, 1, 4),
                          "depth_head.scratch.refinenet4.resConfUnit1.conv1.bias": torch.zeros(4)})
    Wt.pack_depthanything_v2(ignored, cfg, strict=True)
    stray = dict(sd, **{"pretrained.blocks.1.attn.extra.weight": torch.zeros(2)})
    with pytest.raises(RuntimeError, match="Unexpected key"):
        Wt.pack_depthanything_v2(stray, cfg, strict=True)
    Wt.pack_depthanything_v2(stray, cfg, strict=False)
    ref_root = "/root/reference"
    if os.path.isdir(ref_root):  # the live reference agrees (build container only)
        import sys

        sys.path.insert(0, ref_root)
        try:
            from muggled_dpt.make_depthanythingv2_dpt import make_depthanythingv2_dpt_from_original_state_dict as mk
        finally:
            sys.path.remove(ref_root)
        mk(dict(ignored), False, True, True)
        with pytest.raises(RuntimeError, match="Unexpected key"):
            mk(dict(stray), False, True, True)
    bsd = O.make_synthetic_state_dict_beit("beit_tiny", seed=1)
    bcfg = Wt.get_model_config_from_midas_beit_state_dict(bsd, False, True)
    Wt.pack_beit(dict(bsd, **{"pretrained.model.blocks.0.attn.relative_position_index": torch.zeros(3)}), bcfg)
    with pytest.raises(RuntimeError, match="Unexpected key"):
        Wt.pack_beit(dict(bsd, **{"scratch.output_conv.9.weight": torch.zeros(3)}), bcfg)
    ssd = O.make_synthetic_state_dict_swinv2("swinv2_micro", seed=1)
    scfg = Wt.get_model_config_from_midas_swinv2_state_dict(ssd, False, True)
    Wt.pack_swinv2(ssd, scfg)  # carries attn_mask keys (dropped)
    with pytest.raises(RuntimeError, match="Unexpected key"):
        Wt.pack_swinv2(dict(ssd, **{"pretrained.model.layers.0.blocks.0.attn.stray": torch.zeros(3)}), scfg)


def test_model_type_sniffing():
    f = Wt.determine_model_type_from_state_dict
    assert f("x.pth", {"pretrained.model.layers.0.blocks.0.attn.logit_scale": 0}) == "swinv2"
    assert f("x.pth", {"pretrained.model.blocks.0.attn.relative_position_bias_table": 0}) == "beit"
    assert f("depth_anything_v2_vits.pth", {"pretrained.blocks.0.ls1.gamma": 0}) == "depthanythingv2"
    assert f("depth_anything_vitl14.pth", {"pretrained.blocks.0.ls1.gamma": 0}) == "depthanythingv1"
    assert f("x.pth", {}) == "unknown"


def test_factory_surface_on_cpu():
    from muggled_dpt_b200 import make_dpt_from_state_dict

    sd = O.make_synthetic_state_dict("tiny", seed=3, base_grid=5)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "depth_anything_v2_tiny.pth")
        torch.save(sd, path)
        cfg, model = make_dpt_from_state_dict(path)
        with pytest.raises(NotImplementedError):
            make_dpt_from_state_dict(path, model_type="nonsense")
    assert cfg["num_blocks"] == 4
    for attr in ("patch_embed", "imgencoder", "reassemble", "fusion", "head", "inference", "prepare_image_bgr",
                 "verify_input", "to"):
        assert hasattr(model, attr)
    with pytest.raises(RuntimeError):  # no CPU fallback
        model(torch.zeros(1, 3, 56, 56))
    with pytest.raises(RuntimeError):
        model.to("cpu")
    with pytest.raises(RuntimeError):
        model.to(dtype=torch.float32)


def test_beit_config_and_packing():
    sd = O.make_synthetic_state_dict_beit("beit_tiny", seed=5)
    cfg = Wt.get_model_config_from_midas_beit_state_dict(sd, False, True)
    assert list(cfg.keys()) == ["features_per_token", "num_blocks", "num_heads", "reassembly_features_list",
                                "fusion_channels", "patch_size_px", "base_patch_grid_hw", "enable_cache",
                                "enable_optimizations"]
    ocfg = O.infer_config_beit(sd)
    for k, v in ocfg.items():
        assert cfg[k] == v, k
    packed = Wt.pack_beit(sd, cfg)
    # q/v bias -> fused QKV bias with a zero K part (v31_beit/image_encoder_model.py:341-342)
    # ... and LayerNorm 1 folded in: W' = W * ln_w, b' = b + W @ ln_b (weights.fold_layernorm)
    qkv_b = packed["blk2.qkv.b"][0]
    Fd = cfg["features_per_token"]
    Wq = sd["pretrained.model.blocks.2.attn.qkv.weight"].float()
    ln_w, ln_b = sd["pretrained.model.blocks.2.norm1.weight"].float(), sd["pretrained.model.blocks.2.norm1.bias"].float()
    shift = Wq @ ln_b
    assert torch.allclose(qkv_b[:Fd], sd["pretrained.model.blocks.2.attn.q_bias"] + shift[:Fd], atol=1e-6)
    assert torch.allclose(qkv_b[Fd:2 * Fd], shift[Fd:2 * Fd], atol=1e-6)
    assert torch.allclose(qkv_b[2 * Fd:], sd["pretrained.model.blocks.2.attn.v_bias"] + shift[2 * Fd:], atol=1e-6)
    assert packed["blk2.qkv.w"][1] == "half_colsum"
    assert torch.allclose(packed["blk2.qkv.w"][0][:, :Fd], Wq * ln_w[None, :])
    # the fold is exact in fp32: Linear(LN(x)) == rstd * (x W'^T) - rstd * mean * colsum(W') + b'
    torch.manual_seed(1)
    xr = torch.randn(7, Fd) * 2 + 0.3
    ref = torch.nn.functional.layer_norm(xr, (Fd,), ln_w, ln_b, 1e-6) @ Wq.T + torch.cat(
        [sd["pretrained.model.blocks.2.attn.q_bias"], torch.zeros(Fd), sd["pretrained.model.blocks.2.attn.v_bias"]])
    Wp = packed["blk2.qkv.w"][0][:, :Fd]
    mean, var = xr.mean(1, keepdim=True), xr.var(1, unbiased=False, keepdim=True)
    rstd = (var + 1e-6).rsqrt()
    got = rstd * (xr @ Wp.T) - rstd * mean * Wp.sum(1)[None, :] + qkv_b[None, :]
    assert torch.allclose(got, ref, atol=2e-4, rtol=1e-4)
    # readout split: W1 patch + W2 cls + b == Linear(2F,F)(cat(patch, cls))
    torch.manual_seed(0)
    patch, cls = torch.randn(5, Fd), torch.randn(1, Fd)
    W = sd["pretrained.act_postprocess2.0.project.0.weight"]
    b = sd["pretrained.act_postprocess2.0.project.0.bias"]
    ref = F.linear(torch.cat([patch, cls.expand(5, -1)], dim=-1), W, b)
    got = patch @ packed["reasm1.readout.w1"][0][:, :Fd].t() + cls @ packed["reasm1.readout.w2"][0].t() + packed["reasm1.readout.b"][0]
    torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-5)
    assert not any("relative_position_index" in k for k in packed)
    assert not any(k.startswith("fus3.rcu1") for k in packed)


def test_beit_factory_surface_on_cpu():
    from muggled_dpt_b200 import make_dpt_from_state_dict

    sd = O.make_synthetic_state_dict_beit("beit_tiny", seed=5)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "dpt_beit_tiny.pt")
        torch.save(sd, path)
        cfg, model = make_dpt_from_state_dict(path)
    assert model.model_type == "beit" and cfg["num_heads"] == 2 and cfg["base_patch_grid_hw"] == (6, 6)
    assert model.patch_embed.rgb_offset == (0.5, 0.5, 0.5)


def test_swinv2_config_and_packing():
    import math

    sd = O.make_synthetic_state_dict_swinv2("swinv2_micro", seed=4)
    cfg = Wt.get_model_config_from_midas_swinv2_state_dict(sd, False, True)
    assert list(cfg.keys()) == ["features_per_stage", "heads_per_stage", "layers_per_stage", "base_patch_grid_hw",
                                "window_size_hw", "pretrained_window_sizes_per_stage", "fusion_channels",
                                "patch_size_px", "enable_cache", "enable_optimizations"]
    ocfg = O.infer_config_swinv2(sd)
    for k, v in ocfg.items():
        assert cfg[k] == v, k
    packed = Wt.pack_swinv2(sd, cfg)
    assert not any("attn_mask" in k for k in packed)
    # logit_scale: clamp at ln(100) then exp, once, at load (convert_midas_state_dict_keys.py:115-131)
    raw = sd["pretrained.model.layers.3.blocks.1.attn.logit_scale"].reshape(-1)
    got = packed["sw3.1.logit"][0]
    torch.testing.assert_close(got, torch.clamp(raw, max=math.log(100.0)).exp())
    assert got.max() <= 100.0 + 1e-4
    assert packed["sw0.merge.w"][0].shape == (64, 128)  # Linear(4C, 2C), K padded to 64-multiples
    # known pretrained-window LUT for the shipped checkpoints
    big = {"pretrained.model.patch_embed.proj.weight": torch.zeros(192, 3, 4, 4), "scratch.layer1_rn.weight": torch.zeros(256, 192, 3, 3)}
    for st, (h, n) in enumerate(zip((6, 12, 24, 48), (2, 2, 18, 2))):
        for bi in range(n):
            big[f"pretrained.model.layers.{st}.blocks.{bi}.attn.logit_scale"] = torch.zeros(h, 1, 1)
    big["pretrained.model.layers.0.blocks.1.attn_mask"] = torch.zeros(16, 576, 1)
    c2 = Wt.get_model_config_from_midas_swinv2_state_dict(big, False, True)
    assert c2["window_size_hw"] == (24, 24) and c2["base_patch_grid_hw"] == (96, 96)
    assert c2["pretrained_window_sizes_per_stage"] == [12, 12, 12, 6] and c2["layers_per_stage"] == [2, 2, 18, 2]
