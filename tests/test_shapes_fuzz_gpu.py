"""-m gpu: many input shapes (batch, height, width; square and not; token counts on and off the 64 / 128 tile
boundaries; SwinV2 grids that force the reference's window-size adjustment) for a small model of each encoder family,
depth map against the fp32 CPU oracle. The shapes are fixed (seeded), the point is coverage of tails and index maps."""
import os
import tempfile

import pytest
import torch

from gpu_util import gate

pytestmark = pytest.mark.gpu


def _load(sd, fname, dtype):
    from muggled_dpt_b200 import make_dpt_from_state_dict

    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, fname)
        torch.save(sd, path)
        _, model = make_dpt_from_state_dict(path)
    model.to(device="cuda", dtype=dtype)
    return model


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


# (B, H, W): patch 14, even grids; 2x2 ... 18x10 tokens (+cls): N = 5, 17, 25, 37, 49, 65, 101, 129, 181
DA_SHAPES = [(1, 28, 28), (3, 56, 56), (2, 56, 84), (1, 84, 84), (1, 56, 168), (2, 112, 112), (1, 140, 140), (1, 224, 112), (1, 252, 140)]
# patch 16, even grids: N = 9 ... 193 (+cls)
BEIT_SHAPES = [(2, 32, 32), (1, 64, 32), (2, 64, 96), (1, 128, 128), (1, 96, 160), (1, 192, 256)]
# patch 4, grid % 8 == 0; base window 8 (swinv2_micro): 32 -> one window, 160 / 96 -> several, shifted
SWIN_SHAPES = [(2, 32, 32), (1, 64, 64), (1, 128, 160), (2, 96, 64), (1, 160, 96), (1, 224, 32)]


@pytest.mark.parametrize("shape", DA_SHAPES)
def test_depth_anything_tiny_shapes(shape):
    from oracle import dpt_oracle as O

    sd = O.make_synthetic_state_dict("tiny", seed=3, base_grid=5)
    img = O.make_input(*shape, seed=sum(shape))
    ref = O.forward(sd, img)
    model = _load(sd, "depth_anything_v2_tiny.pth", torch.float16)
    with torch.inference_mode():
        out = model(img.to("cuda", torch.float16))
    assert tuple(out.shape) == tuple(shape)
    gate("fuzz.dav2_tiny.fp16.%dx%dx%d.rel_l2" % shape, _rel(out, ref), 3e-3)


@pytest.mark.parametrize("shape", BEIT_SHAPES)
def test_beit_tiny_shapes(shape):
    from oracle import dpt_oracle as O

    sd = O.make_synthetic_state_dict_beit("beit_tiny", seed=5)
    img = O.make_input(*shape, seed=sum(shape))
    ref = O.forward_beit(sd, img)
    model = _load(sd, "dpt_beit_tiny.pt", torch.float16)
    with torch.inference_mode():
        out = model(img.to("cuda", torch.float16))
    assert tuple(out.shape) == tuple(shape)
    gate("fuzz.beit_tiny.fp16.%dx%dx%d.rel_l2" % shape, _rel(out, ref), 4e-3)


@pytest.mark.parametrize("shape", SWIN_SHAPES)
def test_swinv2_micro_shapes(shape):
    from oracle import dpt_oracle as O

    sd = O.make_synthetic_state_dict_swinv2("swinv2_micro", seed=21, logit_std=0.3)
    img = O.make_input(*shape, seed=sum(shape))
    ref = O.forward_swinv2(sd, img)
    model = _load(sd, "dpt_swin2_micro.pt", torch.float16)
    with torch.inference_mode():
        out = model(img.to("cuda", torch.float16))
    assert tuple(out.shape) == tuple(shape)
    gate("fuzz.swinv2_micro.fp16.%dx%dx%d.rel_l2" % shape, _rel(out, ref), 6e-3)


def test_large_images_many_kv_steps():
    """far more tokens than the benchmark sizes: ViT-S at 1008 x 756 (3889 tokens, 61 kv steps per q tile, 31 q tiles),
    a BEiT-shaped model at 512 x 512 (1025 tokens with bias tables resized 4x) and a SwinV2-shaped one at 512 x 384"""
    from oracle import dpt_oracle as O

    sd = O.make_synthetic_state_dict("vits", seed=11)
    img = O.make_input(1, 1008, 756, seed=12)
    ref = O.forward(sd, img)
    model = _load(sd, "depth_anything_v2_vits.pth", torch.float16)
    with torch.inference_mode():
        out = model(img.to("cuda", torch.float16))
    gate("fuzz.vits.fp16.1x1008x756.rel_l2", _rel(out, ref), 3e-3)
    del model

    sd = O.make_synthetic_state_dict_beit("beit_tiny", seed=5)
    img = O.make_input(1, 512, 512, seed=13)
    ref = O.forward_beit(sd, img)
    model = _load(sd, "dpt_beit_tiny.pt", torch.float16)
    with torch.inference_mode():
        out = model(img.to("cuda", torch.float16))
    gate("fuzz.beit_tiny.fp16.1x512x512.rel_l2", _rel(out, ref), 4e-3)
    del model

    sd = O.make_synthetic_state_dict_swinv2("swinv2_micro", seed=21, logit_std=0.3)
    img = O.make_input(1, 512, 384, seed=14)
    ref = O.forward_swinv2(sd, img)
    model = _load(sd, "dpt_swin2_micro.pt", torch.float16)
    with torch.inference_mode():
        out = model(img.to("cuda", torch.float16))
    gate("fuzz.swinv2_micro.fp16.1x512x384.rel_l2", _rel(out, ref), 6e-3)
