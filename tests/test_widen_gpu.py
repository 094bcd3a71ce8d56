"""-m gpu: the CUDA implementations of the rows added after the core path against the golden
vectors of the real reference and the fp32 oracle: Depth-Anything V1 taps, the metric head, the pre-processing kernel
(dpt_prepare_image) and the post-processing kernels (dpt_postprocess_u8), all through the C ABI."""
import os
import tempfile

import numpy as np
import pytest
import torch

from gpu_util import gate

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REL_L2 = {torch.bfloat16: 1.5e-2, torch.float16: 3e-3}  # unpinned defaults; real limits: tests/golden/parity_gates.json


def _dt(dtype):
    return "bf16" if dtype == torch.bfloat16 else "fp16"


def _load(sd, dtype, name):
    from muggled_dpt_b200 import make_dpt_from_state_dict

    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, name)
        torch.save(sd, path)
        cfg, model = make_dpt_from_state_dict(path)
    model.to(device="cuda", dtype=dtype)
    return cfg, model


def _rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_depth_anything_v1_taps_against_reference_golden(dtype):
    from oracle import dpt_oracle as O

    fix = torch.load(os.path.join(GOLDEN, "da_v1_tiny8.pt"))
    sd = O.make_synthetic_state_dict(fix["sd_name"], fix["sd_seed"], fix["sd_base_grid"])
    cfg, model = _load(sd, dtype, "depth_anything_v1_synthetic.pth")
    assert model.model_type == "depthanythingv1"
    img = fix["img"].to("cuda", dtype)
    with torch.inference_mode():
        tokens, grid_hw = model.patch_embed(img)
        taps = model.imgencoder(tokens, grid_hw)
        depth = model(img)
    for k, (a, b) in enumerate(zip(taps, fix["taps"])):
        e = _rel_l2(a, b)
        gate(f"dav1.tiny8.{_dt(dtype)}.tap{k}.rel_l2", e, REL_L2[dtype])
    e = _rel_l2(depth, fix["depth"])
    gate(f"dav1.tiny8.{_dt(dtype)}.depth.rel_l2", e, 2 * REL_L2[dtype])


def test_metric_head_against_reference_golden():
    from oracle import dpt_oracle as O

    fix = torch.load(os.path.join(GOLDEN, "da_v2_metric.pt"))
    sd = O.make_synthetic_state_dict(fix["sd_name"], fix["sd_seed"], fix["sd_base_grid"])
    cfg, model = _load(sd, torch.bfloat16, "depth_anything_v2_metric_synthetic.pth")
    assert cfg["is_metric"] is True
    with torch.inference_mode():
        depth = model(fix["img"].to("cuda", torch.bfloat16))
    e = _rel_l2(depth, fix["depth"])
    gate("dav2.metric.bf16.depth.rel_l2", e, 1.5e-2)
    assert depth.float().min() > 0 and depth.float().max() < 1


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_prepare_image_kernel_against_reference_golden(dtype):
    from oracle import dpt_oracle as O

    sds = {"depthanythingv2": (O.make_synthetic_state_dict("tiny", seed=3, base_grid=5), "depth_anything_v2_t.pth"),
           "beit": (O.make_synthetic_state_dict_beit("beit_tiny", seed=5), "dpt_beit_t.pt"),
           "swinv2": (O.make_synthetic_state_dict_swinv2("swinv2_micro", seed=7), "dpt_swinv2_t.pt")}
    models = {}
    for case in torch.load(os.path.join(GOLDEN, "prepare_image.pt")):
        mt = case["model_type"]
        if mt not in models:
            models[mt] = _load(sds[mt][0], dtype, sds[mt][1])[1]
        model = models[mt]
        assert model.config["patch_size_px"] == case["patch"]  # (every case passes max_side_length: base grid unused)
        out = model.prepare_image_bgr(case["bgr"].numpy(), **case["kwargs"])
        ref = case["out"]
        assert tuple(out.shape) == tuple(ref.shape) and out.dtype == dtype, case["name"]
        err = (out.float().cpu() - ref).abs().max().item()
        # fp32 arithmetic, one rounding to the 16-bit output: values are within +-2.7
        tol = 2.7 * (2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11) + 2e-4
        print(f"prepare_image {case['name']} {dtype}: max abs err {err:.2e} (tol {tol:.2e})")
        assert err <= tol, (case["name"], err)


def test_prepare_image_rejects_bad_input():
    from oracle import dpt_oracle as O

    _, model = _load(O.make_synthetic_state_dict("tiny", seed=3, base_grid=5), torch.bfloat16, "depth_anything_v2_t.pth")
    with pytest.raises(ValueError):
        model.prepare_image_bgr(np.zeros((20, 20, 3), dtype=np.float32))
    with pytest.raises(ValueError):
        model.prepare_image_bgr(np.zeros((20, 20), dtype=np.uint8))


def test_postprocess_kernels_against_reference_golden():
    from muggled_dpt_b200.postprocess import scale_normalize_to_uint8

    for case in torch.load(os.path.join(GOLDEN, "postprocess.pt")):
        out = scale_normalize_to_uint8(case["pred"].cuda(), case["target_wh"])
        ref = case["out"]
        assert out.dtype == torch.uint8 and tuple(out.shape) == tuple(ref.shape)
        d = (out.cpu().int() - ref.int()).abs()
        # truncation to uint8: fp32 rounding order can flip a value sitting on an integer boundary by one code
        print(f"postprocess {tuple(ref.shape)}: max code diff {d.max().item()}, differing {(d > 0).float().mean().item():.2e}")
        assert d.max().item() <= 1 and (d > 0).float().mean().item() < 2e-3
        assert out.min().item() == 0 and out.max().item() == 255
    with pytest.raises(RuntimeError):
        scale_normalize_to_uint8(torch.zeros(1, 4, 4, dtype=torch.bfloat16), (8, 8))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_vit_giant_swiglu_against_reference_golden(dtype):
    from oracle import dpt_oracle as O

    fix = torch.load(os.path.join(GOLDEN, "da_v2_giant_tiny.pt"))
    sd = O.giantify(O.make_synthetic_state_dict(fix["sd_name"], fix["sd_seed"], fix["sd_base_grid"]), seed=fix["sd_seed"])
    cfg, model = _load(sd, dtype, "depth_anything_v2_vitg_synthetic.pth")
    assert cfg["is_giant"] is True
    img = fix["img"].to("cuda", dtype)
    with torch.inference_mode():
        taps = model.imgencoder(fix["tokens"].to("cuda", dtype), tuple(fix["grid_hw"]))
        depth = model(img)
    for k, (a, b) in enumerate(zip(taps, fix["taps"])):
        e = _rel_l2(a, b)
        gate(f"dav2.giant_tiny.{_dt(dtype)}.tap{k}.rel_l2", e, REL_L2[dtype])
    e = _rel_l2(depth, fix["depth"])
    gate(f"dav2.giant_tiny.{_dt(dtype)}.depth.rel_l2", e, 2 * REL_L2[dtype])
