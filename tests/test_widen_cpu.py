"""CPU: the rows added after the core path - the oracle restatements of the Depth-Anything V1 tap
rule, the metric head, PatchEmbed.prepare_image and the demo post-processing chain against golden vectors produced by
the real reference (oracle/make_golden_widen.py), plus the host logic (type sniffing, V1 config keys)."""
import os
import tempfile

import pytest
import torch

from oracle import dpt_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_oracle_v1_taps_match_reference_every_stage():
    fix = torch.load(os.path.join(GOLDEN, "da_v1_tiny8.pt"))
    sd = O.make_synthetic_state_dict(fix["sd_name"], fix["sd_seed"], fix["sd_base_grid"])
    cfg = O.infer_config(sd)
    cfg["taps_last4"] = True
    st = O.forward(sd, fix["img"], cfg=cfg, return_stages=True)
    for a, b in zip(st["taps"], fix["taps"]):
        torch.testing.assert_close(a, b, rtol=0, atol=2e-5)
    torch.testing.assert_close(st["depth"], fix["depth"], rtol=0, atol=1e-4)
    # the V2 rule on the same weights gives different taps: the fixture discriminates the two rules
    st2 = O.forward(sd, fix["img"], return_stages=True)
    assert (st2["taps"][0] - fix["taps"][0]).abs().max() > 1e-2


def test_oracle_metric_head_matches_reference():
    fix = torch.load(os.path.join(GOLDEN, "da_v2_metric.pt"))
    sd = O.make_synthetic_state_dict(fix["sd_name"], fix["sd_seed"], fix["sd_base_grid"])
    sd["is_metric"] = torch.tensor(1.0)
    depth = O.forward(sd, fix["img"])
    torch.testing.assert_close(depth, fix["depth"], rtol=0, atol=1e-5)
    assert 0.0 < depth.min() and depth.max() < 1.0  # Sigmoid range


def test_oracle_prepare_image_matches_reference():
    for case in torch.load(os.path.join(GOLDEN, "prepare_image.pt")):
        out = O.prepare_image(case["bgr"].numpy(), case["patch"], case["base_grid"], case["model_type"], **case["kwargs"])
        assert out.shape == case["out"].shape, case["name"]
        torch.testing.assert_close(out, case["out"], rtol=0, atol=1e-6)


def test_oracle_postprocess_matches_reference():
    for case in torch.load(os.path.join(GOLDEN, "postprocess.pt")):
        out = O.postprocess_u8(case["pred"].float(), case["target_wh"])
        assert out.dtype == torch.uint8 and torch.equal(out, case["out"])


def test_v1_factory_type_sniffing_and_config_keys():
    from muggled_dpt_b200 import make_dpt_from_state_dict
    from muggled_dpt_b200.weights import determine_model_type_from_state_dict

    fix = torch.load(os.path.join(GOLDEN, "da_v1_tiny8.pt"))
    sd = O.make_synthetic_state_dict(fix["sd_name"], fix["sd_seed"], fix["sd_base_grid"])
    # make_dpt.py:98-104: the version comes from the FILE NAME
    assert determine_model_type_from_state_dict("/x/depth_anything_vitl14.pth", sd) == "depthanythingv1"
    assert determine_model_type_from_state_dict("/x/depth_anything_v1_vits.pth", sd) == "depthanythingv1"
    assert determine_model_type_from_state_dict("/x/depth_anything_v2_vits.pth", sd) == "depthanythingv2"
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "depth_anything_v1_synthetic.pth")
        torch.save(sd, path)
        cfg, model = make_dpt_from_state_dict(path)
    assert model.model_type == "depthanythingv1"
    assert list(cfg.keys()) == list(fix["config"].keys())  # no is_giant / is_metric keys in the V1 config
    for k, v in fix["config"].items():
        assert (list(cfg[k]) if isinstance(cfg[k], (list, tuple)) else cfg[k]) == v, k
    with pytest.raises(RuntimeError):
        model.to("cpu")  # CUDA-only, like every other variant


def test_oracle_vit_giant_swiglu_matches_reference():
    fix = torch.load(os.path.join(GOLDEN, "da_v2_giant_tiny.pt"))
    sd = O.giantify(O.make_synthetic_state_dict(fix["sd_name"], fix["sd_seed"], fix["sd_base_grid"]), seed=fix["sd_seed"])
    from oracle.make_golden import state_dict_checksum

    assert state_dict_checksum(sd) == fix["sd_checksum"]
    st = O.forward(sd, fix["img"], return_stages=True)
    for a, b in zip(st["taps"], fix["taps"]):
        torch.testing.assert_close(a, b, rtol=0, atol=2e-5)
    torch.testing.assert_close(st["depth"], fix["depth"], rtol=0, atol=1e-4)


def test_vit_giant_config_and_packing():
    from muggled_dpt_b200 import weights as Wt

    fix = torch.load(os.path.join(GOLDEN, "da_v2_giant_tiny.pt"))
    sd = O.giantify(O.make_synthetic_state_dict(fix["sd_name"], fix["sd_seed"], fix["sd_base_grid"]), seed=fix["sd_seed"])
    cfg = Wt.get_model_config_from_state_dict(sd, False, True)
    assert cfg["is_giant"] is True and list(cfg.keys()) == list(fix["config"].keys())
    packed = Wt.pack_depthanything_v2(sd, cfg)
    Fd, h = cfg["features_per_token"], O.swiglu_hidden_features(cfg["features_per_token"])
    hp = (h + 63) // 64 * 64
    w1, b1 = packed["blk1.fc1.w"][0], packed["blk1.fc1.b"][0]
    assert w1.shape == (2 * hp, Fd) and packed["blk1.fc1.w"][1] == "half_colsum" and b1.shape == (2 * hp,)
    # rows interleaved for the fused gate epilogue: blocks of 32 gate rows + the 32 linear rows of the same features;
    # undoing the interleave gives back LN-folded w12 (gate half first), the padding rows are zero
    ln_w, ln_b = sd["pretrained.blocks.1.norm2.weight"], sd["pretrained.blocks.1.norm2.bias"]
    w12, b12 = sd["pretrained.blocks.1.mlp.w12.weight"], sd["pretrained.blocks.1.mlp.w12.bias"]
    blocks = w1.reshape(hp // 32, 2, 32, Fd)
    torch.testing.assert_close(blocks[:, 0].reshape(hp, Fd)[:h], w12[:h] * ln_w[None, :])
    torch.testing.assert_close(blocks[:, 1].reshape(hp, Fd)[:h], w12[h:] * ln_w[None, :])
    assert torch.count_nonzero(blocks[:, 0].reshape(hp, Fd)[h:]) == 0 and torch.count_nonzero(blocks[:, 1].reshape(hp, Fd)[h:]) == 0
    bb = b1.reshape(hp // 32, 2, 32)
    torch.testing.assert_close(bb[:, 0].reshape(hp)[:h], b12[:h] + w12[:h] @ ln_b)
    torch.testing.assert_close(bb[:, 1].reshape(hp)[:h], b12[h:] + w12[h:] @ ln_b)
    assert packed["blk1.fc2.w"][0].shape == (Fd, (h + 63) // 64 * 64)  # K padded to the GEMM chunk with zeros
    assert torch.count_nonzero(packed["blk1.fc2.w"][0][:, h:]) == 0
