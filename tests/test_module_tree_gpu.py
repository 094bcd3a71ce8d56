"""-m gpu: SURVEY.md section 8f-3 - the module tree the reference's experiments/ reach into. The product runs next to
the UNMODIFIED reference (oracle/_ref, fp32 CPU, enable_optimizations=False so its manual attention path with the
nn.Softmax module is used) on the same synthetic checkpoint and input; both are instrumented with the reference's own
ModelOutputCapture (demo_helpers/model_capture.py) exactly as experiments/attention_visualization.py:324-332 and
experiments/block_norm_visualization.py:265-300 do, and experiments/fusion_scaling.py:330-333's per-block fusion calls
are replayed on both."""
import os
import sys
import tempfile

import pytest
import torch

from gpu_util import gate

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference():
    from oracle.build_ref import import_reference, ref_available

    if not ref_available():
        pytest.skip("oracle/_ref is absent (run oracle/build_ref.py where /root/reference exists)")
    return import_reference()


def _both_models(sd, fname, dtype):
    from muggled_dpt_b200 import make_dpt_from_state_dict

    ref_make = _reference()
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, fname)
        torch.save(sd, path)
        _, ref_model = ref_make.make_dpt_from_state_dict(path, enable_optimizations=False)
        _, model = make_dpt_from_state_dict(path, enable_optimizations=False)
    model.to(device="cuda", dtype=dtype)
    return ref_model, model


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


CASES = {
    "dav2": ("tiny", "depth_anything_v2_tiny.pth", (2, 56, 84)),
    "beit": ("beit_tiny", "dpt_beit_tiny.pt", (2, 64, 96)),
    "swin": ("swinv2_micro", "dpt_swin2_micro.pt", (1, 128, 160)),
}


def _case(fam):
    from oracle import dpt_oracle as O

    name, fname, shape = CASES[fam]
    if fam == "dav2":
        sd = O.make_synthetic_state_dict(name, seed=3, base_grid=5)
    elif fam == "beit":
        sd = O.make_synthetic_state_dict_beit(name, seed=5)
    else:
        sd = O.make_synthetic_state_dict_swinv2(name, seed=21, logit_std=0.3)
    return sd, fname, O.make_input(*shape, seed=4)


@pytest.mark.parametrize("fam", ["dav2", "beit", "swin"])
def test_softmax_hooks_receive_the_attention_probabilities(fam):
    _reference()
    from muggled_dpt.demo_helpers.model_capture import ModelOutputCapture  # the reference's helper, from oracle/_ref

    sd, fname, img = _case(fam)
    dtype = torch.float16
    ref_model, model = _both_models(sd, fname, dtype)
    ref_cap = ModelOutputCapture(ref_model, torch.nn.Softmax)
    cap = ModelOutputCapture(model, torch.nn.Softmax)
    with torch.inference_mode():
        rt, rgrid = ref_model.patch_embed(img)
        ref_taps = ref_model.imgencoder(rt, rgrid)
        tokens, grid = model.patch_embed(img.to("cuda", dtype))
        taps = model.imgencoder(rt.to("cuda", dtype), grid)
    assert len(cap) == len(ref_cap) > 0
    worst = 0.0
    for i, (a, b) in enumerate(zip(cap, ref_cap)):
        assert tuple(a.shape) == tuple(b.shape), (i, a.shape, b.shape)
        row_sums = a.float().sum(-1)
        assert torch.allclose(row_sums, torch.ones_like(row_sums), atol=4e-3)
        worst = max(worst, _rel(a, b))
    gate(f"hooks.softmax.{fam}.fp16.rel_l2", worst, 2e-2)
    for i in range(4):  # the captured run returns the same taps as the plain one
        gate(f"hooks.softmax.{fam}.fp16.tap{i}.rel_l2", _rel(taps[i], ref_taps[i]), 2e-2)
    plain = type(model.imgencoder).forward  # no hooks on a fresh model -> single C call, identical taps
    _, model2 = _both_models(sd, fname, dtype)
    with torch.inference_mode():
        taps2 = model2.imgencoder(rt.to("cuda", dtype), grid)
    for a, b in zip(taps, taps2):
        assert torch.equal(a, b)
    assert plain is not None


@pytest.mark.parametrize("fam", ["dav2", "beit", "swin"])
def test_block_hooks_receive_each_block_output(fam):
    _reference()
    from muggled_dpt.demo_helpers.model_capture import ModelOutputCapture
    from muggled_dpt_b200.module_tree import TransformerBlock

    if fam == "dav2":
        from muggled_dpt.v2_depthanything.image_encoder_model import TransformerBlock as RefBlock
    elif fam == "beit":
        from muggled_dpt.v31_beit.image_encoder_model import TransformerBlock as RefBlock
    else:
        from muggled_dpt.v31_swinv2.image_encoder_model import SwinTransformerBlock as RefBlock
    sd, fname, img = _case(fam)
    dtype = torch.float16
    ref_model, model = _both_models(sd, fname, dtype)
    ref_cap = ModelOutputCapture(ref_model, RefBlock)
    cap = ModelOutputCapture(model, TransformerBlock)
    with torch.inference_mode():
        rt, rgrid = ref_model.patch_embed(img)
        ref_model.imgencoder(rt, rgrid)
        _, grid = model.patch_embed(img.to("cuda", dtype))
        model.imgencoder(rt.to("cuda", dtype), grid)
    assert len(cap) == len(ref_cap) > 0
    worst = 0.0
    for i, (a, b) in enumerate(zip(cap, ref_cap)):
        assert tuple(a.shape) == tuple(b.shape), (i, a.shape, b.shape)
        worst = max(worst, _rel(a, b))
    gate(f"hooks.block.{fam}.fp16.rel_l2", worst, 2e-2)


@pytest.mark.parametrize("fam", ["dav2", "beit", "swin"])
def test_fusion_blocks_are_callable_one_by_one(fam):
    """experiments/fusion_scaling.py:330-334 with scale factors != 1"""
    sd, fname, img = _case(fam)
    dtype = torch.float16
    ref_model, model = _both_models(sd, fname, dtype)
    scales = (0.5, 1.5, 0.75, 1.25)
    with torch.inference_mode():
        rt, rgrid = ref_model.patch_embed(img)
        r = ref_model.reassemble(*ref_model.imgencoder(rt, rgrid), rgrid)
        f3 = ref_model.fusion.blocks[3](r[3] * scales[3])
        f2 = ref_model.fusion.blocks[2](r[2], f3 * scales[2])
        f1 = ref_model.fusion.blocks[1](r[1], f2 * scales[1])
        f0 = ref_model.fusion.blocks[0](r[0], f1 * scales[0])
        ref_pred = ref_model.head(f0)
        m = [t.to("cuda", dtype) for t in r]
        g3 = model.fusion.blocks[3](m[3] * scales[3])
        g2 = model.fusion.blocks[2](m[2], g3 * scales[2])
        g1 = model.fusion.blocks[1](m[1], g2 * scales[1])
        g0 = model.fusion.blocks[0](m[0], g1 * scales[0])
        pred = model.head(g0)
        # unit scales: the four calls reproduce model.fusion(...) bit for bit
        h3 = model.fusion.blocks[3](m[3])
        h2 = model.fusion.blocks[2](m[2], h3)
        h1 = model.fusion.blocks[1](m[1], h2)
        h0 = model.fusion.blocks[0](m[0], h1)
        whole = model.fusion(*m)
    assert torch.equal(h0, whole)
    for name, a, b in (("f3", g3, f3), ("f2", g2, f2), ("f1", g1, f1), ("f0", g0, f0), ("depth", pred, ref_pred)):
        assert tuple(a.shape) == tuple(b.shape), (name, a.shape, b.shape)
        gate(f"fusion_blocks.{fam}.fp16.{name}.rel_l2", _rel(a, b), 6e-3)
    with pytest.raises(TypeError):
        model.fusion.blocks[1](m[1])
