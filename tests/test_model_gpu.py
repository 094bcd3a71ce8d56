"""Model-level parity (-m gpu): the CUDA path, through the reference-shaped Python surface (make_dpt_from_state_dict /
DPTModel stage calls / forward) and the C ABI underneath, against (a) the committed golden fixtures produced by the
real reference (oracle/make_golden.py) and (b) the fp32 CPU oracle on the same seeded inputs.

Tolerances (stated, per the north star's "fp16/bf16 tolerance"): the reference's own bf16-vs-fp32 CPU gap is
3.8e-3 max-rel on the depth map. We gate on relative L2 error per stage and max-rel on depth.
"""
import os
import tempfile

import pytest
import torch

from gpu_util import gate

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# loose defaults for a gate that has not been pinned yet; the real limits (<= 1.5 x the value measured on B200) live in
# tests/golden/parity_gates.json, see gpu_util.gate
REL_L2 = {torch.bfloat16: 1.5e-2, torch.float16: 3e-3}
DEPTH_MAXREL = {torch.bfloat16: 3e-2, torch.float16: 5e-3}


def _dt(dtype):
    return "bf16" if dtype == torch.bfloat16 else "fp16"


def _load_model(sd, dtype, name="depth_anything_v2_synth.pth"):
    from muggled_dpt_b200 import make_dpt_from_state_dict

    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, name)
        torch.save(sd, path)
        cfg, model = make_dpt_from_state_dict(path)
    model.to(device="cuda", dtype=dtype, memory_format=torch.channels_last)
    return cfg, model


def _err(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    rel_l2 = ((a - b).norm() / b.norm().clamp_min(1e-12)).item()
    max_abs = (a - b).abs().max().item()
    # max error relative to the largest reference magnitude (depth maps contain values near 0 after the final ReLU,
    # where an element-wise ratio is meaningless)
    max_rel = max_abs / (b.abs().max().item() + 1e-12)
    return rel_l2, max_abs, max_rel


def _fixture(name):
    from oracle import dpt_oracle as O

    fix = torch.load(os.path.join(GOLDEN, name))
    sd = fix.get("state_dict") or O.make_synthetic_state_dict(fix["sd_name"], fix["sd_seed"], fix["sd_base_grid"])
    return fix, sd


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("name", ["tiny_a.pt", "tiny_b.pt"])
def test_tiny_stagewise_against_reference_golden(name, dtype):
    fix, sd = _fixture(name)
    cfg, model = _load_model(sd, dtype)
    assert cfg["features_per_token"] == fix["config"]["features_per_token"]
    img = fix["img"].to("cuda", dtype)
    report = {}
    with torch.inference_mode():
        tokens, grid_hw = model.patch_embed(img)
        assert tuple(grid_hw) == tuple(fix["grid_hw"])
        report["tokens"] = _err(tokens, fix["tokens"])
        # each stage is fed the reference's own (rounded) input so errors do not compound
        taps = model.imgencoder(fix["tokens"].to("cuda", dtype), grid_hw)
        for i in range(4):
            report[f"tap{i}"] = _err(taps[i], fix["taps"][i])
        maps = model.reassemble(*[t.to("cuda", dtype) for t in fix["taps"]], grid_hw)
        for i in range(4):
            assert tuple(maps[i].shape) == tuple(fix["maps"][i].shape)
            report[f"map{i}"] = _err(maps[i], fix["maps"][i])
        fused = model.fusion(*[t.to("cuda", dtype) for t in fix["maps"]])
        assert tuple(fused.shape) == tuple(fix["fused"].shape)
        report["fused"] = _err(fused, fix["fused"])
        depth = model.head(fix["fused"].to("cuda", dtype))
        report["head"] = _err(depth, fix["depth"])
        full = model(img)
        report["depth_e2e"] = _err(full, fix["depth"])
    for k, v in report.items():
        print(f"{name} {dtype} {k}: rel_l2={v[0]:.3e} max_abs={v[1]:.3e} max_rel={v[2]:.3e}")
    for k, v in report.items():
        gate(f"dav2.{name}.{_dt(dtype)}.{k}.rel_l2", v[0], REL_L2[dtype] * (2.0 if k == "depth_e2e" else 1.0))
    gate(f"dav2.{name}.{_dt(dtype)}.depth_e2e.max_rel", report["depth_e2e"][2], DEPTH_MAXREL[dtype])


@pytest.mark.parametrize("name", ["vits_a.pt", "vits_b.pt"])
def test_vits_against_reference_golden(name):
    from oracle import dpt_oracle as O
    from oracle.make_golden import state_dict_checksum, sub

    fix, sd = _fixture(name)
    assert state_dict_checksum(sd) == fix["sd_checksum"], "seeded checkpoint differs from the one the fixture was made with"
    dtype = torch.bfloat16
    cfg, model = _load_model(sd, dtype)
    img = fix["img"].to("cuda", dtype)
    with torch.inference_mode():
        depth = model(img)
        tokens, grid_hw = model.patch_embed(img)
        taps = model.imgencoder(tokens, grid_hw)
    e = _err(depth, fix["depth"])
    print(f"{name} depth: rel_l2={e[0]:.3e} max_abs={e[1]:.3e} max_rel={e[2]:.3e}")
    et = _err(sub(tokens.float().cpu()), fix["tokens_sub"])
    print(f"{name} tokens(sub): rel_l2={et[0]:.3e}")
    for i in range(4):
        ei = _err(sub(taps[i].float().cpu()), fix["taps_sub"][i])
        print(f"{name} tap{i}(sub): rel_l2={ei[0]:.3e}")
        gate(f"dav2.{name}.bf16.tap{i}_sub.rel_l2", ei[0], 3e-2)
    gate(f"dav2.{name}.bf16.tokens_sub.rel_l2", et[0], REL_L2[dtype])
    gate(f"dav2.{name}.bf16.depth.rel_l2", e[0], 2 * REL_L2[dtype])
    gate(f"dav2.{name}.bf16.depth.max_rel", e[2], DEPTH_MAXREL[dtype])


def test_vits_504_against_oracle():
    """config S (BASELINE.json configs[0]): ViT-S, 1x3x504x504 (the reference's effective '518'), vs the fp32 oracle"""
    from oracle import dpt_oracle as O

    sd = O.make_synthetic_state_dict("vits", seed=11)
    img = O.make_input(1, 504, 504, seed=2)
    ref = O.forward(sd, img)
    for dtype in (torch.bfloat16, torch.float16):
        cfg, model = _load_model(sd, dtype)
        with torch.inference_mode():
            depth = model(img.to("cuda", dtype))
        e = _err(depth, ref)
        print(f"vits504 {dtype}: rel_l2={e[0]:.3e} max_abs={e[1]:.3e} max_rel={e[2]:.3e}")
        assert tuple(depth.shape) == (1, 504, 504)
        gate(f"dav2.vits504.{_dt(dtype)}.depth.rel_l2", e[0], 2 * REL_L2[dtype])
        gate(f"dav2.vits504.{_dt(dtype)}.depth.max_rel", e[2], DEPTH_MAXREL[dtype])


def test_small_batch_wide_pair_tiles_against_oracle():
    """B = 4 at 504^2 with ViT-L widths (M = 5188 rows, the per-GPU shard of the 8-GPU strong-scaling run): the fp32
    residual GEMMs (proj, fc2: N = 1024) run as 256 x 384 CTA-pair tiles in one round (gemm_tc.cuh BLOCK_N = 384, with
    and without the residual prefetch), writing the row statistics and the 16-bit copy the next GEMM reads."""
    from oracle import dpt_oracle as O

    sd = O.make_synthetic_state_dict("vitl_4blk", seed=17)
    img = O.make_input(4, 504, 504, seed=5)
    ref = O.forward(sd, img, return_stages=True)
    for dtype in (torch.bfloat16, torch.float16):
        cfg, model = _load_model(sd, dtype)
        with torch.inference_mode():
            x = img.to("cuda", dtype)
            tokens, grid = model.patch_embed(x)
            taps = model.imgencoder(tokens, grid)
            model.enable_profiling(True)
            depth = model(x)
            torch.cuda.synchronize()
            labels = [r[0] for r in model.read_profile()]
            model.enable_profiling(False)
        assert sum(l.startswith("gemm384x2:") and l.endswith(".proj") for l in labels) == 4, labels[:12]
        assert sum(l.startswith("gemm384x2:") and l.endswith(".fc2") for l in labels) == 4
        for i in range(4):
            e = _err(taps[i], ref["taps"][i])
            gate(f"dav2.vitl_4blk_B4.{_dt(dtype)}.tap{i}.rel_l2", e[0], REL_L2[dtype])
        e = _err(depth, ref["depth"])
        gate(f"dav2.vitl_4blk_B4.{_dt(dtype)}.depth.rel_l2", e[0], 2 * REL_L2[dtype])
        gate(f"dav2.vitl_4blk_B4.{_dt(dtype)}.depth.max_rel", e[2], DEPTH_MAXREL[dtype])


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_wide_reassembly_merged_conv_transpose_against_oracle(dtype):
    """256 reassembly channels (the ViT-L case): ConvTranspose2d k=s=4 / k=s=2 run as ONE pixel-shuffling GEMM launch
    each (gemm_tc.cuh shuffle_n) instead of s*s launches; checked per map against the fp32 oracle."""
    from oracle import dpt_oracle as O

    sd = O.make_synthetic_state_dict("tiny_r256", seed=13, base_grid=5)
    img = O.make_input(2, 56, 84, seed=3)
    ref = O.forward(sd, img, return_stages=True)
    cfg, model = _load_model(sd, dtype)
    assert cfg["reassembly_features_list"] == [256, 256, 64, 128]
    with torch.inference_mode():
        grid = tuple(ref["grid_hw"])
        maps = model.reassemble(*[t.to("cuda", dtype) for t in ref["taps"]], grid)
        depth = model(img.to("cuda", dtype))
    for i in range(4):
        assert tuple(maps[i].shape) == tuple(ref["maps"][i].shape)
        e = _err(maps[i], ref["maps"][i])
        print(f"tiny_r256 {dtype} map{i}: rel_l2={e[0]:.3e}")
        gate(f"dav2.tiny_r256.{_dt(dtype)}.map{i}.rel_l2", e[0], REL_L2[dtype])
    e = _err(depth, ref["depth"])
    gate(f"dav2.tiny_r256.{_dt(dtype)}.depth.rel_l2", e[0], 2 * REL_L2[dtype])


def test_head_variants_fused_resize_and_generic_conv_agree_with_default():
    """The head's second convolution has three implementations selected by environment switches read once per
    process: the default TMA-fed halo kernel, DPT_HALO_FUSE=1 (halo tiles interpolated inside the kernel, the up-sampled
    map never materialised) and DPT_HALO=0 (generic nine-load spatial GEMM). All must match the oracle."""
    import subprocess
    import sys

    code = (
        "import os, sys, tempfile, torch\n"
        "sys.path.insert(0, os.getcwd())\n"
        "from muggled_dpt_b200 import make_dpt_from_state_dict\n"
        "from oracle import dpt_oracle as O\n"
        "for name, seed, hw in (('tiny', 3, (56, 84)), ('vits', 11, (112, 140))):\n"
        "    sd = O.make_synthetic_state_dict(name, seed=seed, base_grid=5 if name == 'tiny' else 37)\n"
        "    img = O.make_input(2, hw[0], hw[1], seed=5)\n"
        "    ref = O.forward(sd, img)\n"
        "    with tempfile.TemporaryDirectory() as td:\n"
        "        p = os.path.join(td, 'depth_anything_v2_x.pth'); torch.save(sd, p)\n"
        "        _, m = make_dpt_from_state_dict(p)\n"
        "    m.to(device='cuda', dtype=torch.float16)\n"
        "    out = m(img.to('cuda', torch.float16)).float().cpu()\n"
        "    e = ((out - ref).norm() / ref.norm()).item()\n"
        "    print(name, 'rel_l2', e)\n"
        "    assert e < 3e-3, e\n"
    )
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for env_extra in ({"DPT_HALO_FUSE": "1"}, {"DPT_HALO": "0"}):
        env = dict(os.environ, **env_extra)
        r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=600)
        print(env_extra, r.stdout.strip().replace("\n", " | "))
        assert r.returncode == 0, (env_extra, r.stdout[-2000:], r.stderr[-2000:])


def test_properties_batch_independence_and_determinism():
    """size-independent properties: frames of a batch do not interact (dpt_model.py has no cross-batch op), and the
    path is deterministic run to run"""
    from oracle import dpt_oracle as O

    sd = O.make_synthetic_state_dict("vits", seed=11)
    cfg, model = _load_model(sd, torch.bfloat16)
    img = O.make_input(3, 252, 196, seed=9).to("cuda", torch.bfloat16)
    with torch.inference_mode():
        d_all = model(img)
        d_again = model(img)
        d_one = model(img[1:2].contiguous())
    assert torch.equal(d_all, d_again)
    assert torch.equal(d_all[1:2], d_one)


def test_error_behaviour():
    from oracle import dpt_oracle as O

    sd = O.make_synthetic_state_dict("tiny", seed=3, base_grid=5)
    cfg, model = _load_model(sd, torch.bfloat16)
    with pytest.raises(ValueError):  # 42x42 -> 3x3 grid (odd): the reference raises RuntimeError inside fusion
        model(torch.zeros(1, 3, 42, 42, device="cuda", dtype=torch.bfloat16))
    with pytest.raises(RuntimeError):  # dtype mismatch (dpt_model.py:151-155)
        model(torch.zeros(1, 3, 56, 56, device="cuda", dtype=torch.float16))
    with pytest.raises(AssertionError):
        model.verify_input(torch.zeros(1, 3, 57, 56, device="cuda", dtype=torch.bfloat16))
    assert model.verify_input(torch.zeros(1, 3, 56, 56, device="cuda", dtype=torch.bfloat16))
    with pytest.raises(RuntimeError):
        model.to("cpu")


def test_inference_entry_point_bgr_image():
    import numpy as np
    from oracle import dpt_oracle as O

    sd = O.make_synthetic_state_dict("tiny", seed=3, base_grid=5)
    cfg, model = _load_model(sd, torch.bfloat16)
    rng = np.random.default_rng(0)
    bgr = rng.integers(0, 255, size=(90, 120, 3), dtype=np.uint8)
    out = model.inference(bgr, max_side_length=112, use_square_sizing=True)
    assert tuple(out.shape) == (1, 112, 112) and out.dtype == torch.bfloat16
    assert torch.isfinite(out.float()).all()


# ------------------------------------------------------------------------------------------------ MiDaS v3.1 BEiT


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("name", ["beit_tiny_a.pt", "beit_tiny_b.pt"])
def test_beit_tiny_stagewise_against_reference_golden(name, dtype):
    from oracle import dpt_oracle as O

    fix = torch.load(os.path.join(GOLDEN, name))
    sd = O.make_synthetic_state_dict_beit(fix["sd_name"], fix["sd_seed"])
    cfg, model = _load_model(sd, dtype, name="dpt_beit_synth.pt")
    assert model.model_type == "beit"
    img = fix["img"].to("cuda", dtype)
    report = {}
    with torch.inference_mode():
        tokens, grid_hw = model.patch_embed(img)
        assert tuple(grid_hw) == tuple(fix["grid_hw"])
        report["tokens"] = _err(tokens, fix["tokens"])
        taps = model.imgencoder(fix["tokens"].to("cuda", dtype), grid_hw)
        for i in range(4):
            report[f"tap{i}"] = _err(taps[i], fix["taps"][i])
        maps = model.reassemble(*[t.to("cuda", dtype) for t in fix["taps"]], grid_hw)
        for i in range(4):
            assert tuple(maps[i].shape) == tuple(fix["maps"][i].shape)
            report[f"map{i}"] = _err(maps[i], fix["maps"][i])
        fused = model.fusion(*[t.to("cuda", dtype) for t in fix["maps"]])
        report["fused"] = _err(fused, fix["fused"])
        depth = model.head(fix["fused"].to("cuda", dtype))
        report["head"] = _err(depth, fix["depth"])
        full = model(img)
        assert tuple(full.shape) == tuple(fix["depth"].shape)
        report["depth_e2e"] = _err(full, fix["depth"])
    for k, v in report.items():
        print(f"{name} {dtype} {k}: rel_l2={v[0]:.3e} max_abs={v[1]:.3e} max_rel={v[2]:.3e}")
    for k, v in report.items():
        gate(f"beit.{name}.{_dt(dtype)}.{k}.rel_l2", v[0], REL_L2[dtype] * (2.0 if k == "depth_e2e" else 1.0))
    gate(f"beit.{name}.{_dt(dtype)}.depth_e2e.max_rel", report["depth_e2e"][2], DEPTH_MAXREL[dtype])


def test_beit_base_384_against_oracle():
    """config E shape family (BASELINE.json configs[4]): BEiT at 384x384 (24x24 grid, 577 tokens, bias attention),
    here the 12-block base model with batch 2, vs the fp32 oracle"""
    from oracle import dpt_oracle as O

    sd = O.make_synthetic_state_dict_beit("beit_base_384", seed=9)
    img = O.make_input(2, 384, 384, seed=4)
    ref = O.forward_beit(sd, img)
    cfg, model = _load_model(sd, torch.bfloat16, name="dpt_beit_base_384.pt")
    with torch.inference_mode():
        depth = model(img.to("cuda", torch.bfloat16))
    e = _err(depth, ref)
    print(f"beit_base_384 bf16: rel_l2={e[0]:.3e} max_abs={e[1]:.3e} max_rel={e[2]:.3e}")
    assert tuple(depth.shape) == (2, 384, 384)
    gate("beit.base384.bf16.depth.rel_l2", e[0], 2 * REL_L2[torch.bfloat16])
    gate("beit.base384.bf16.depth.max_rel", e[2], DEPTH_MAXREL[torch.bfloat16])


# ------------------------------------------------------------------------------------------------ MiDaS v3.1 SwinV2


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("name", ["swinv2_micro_a.pt", "swinv2_micro_b.pt"])
def test_swinv2_micro_stagewise_against_reference_golden(name, dtype):
    from oracle import dpt_oracle as O

    fix = torch.load(os.path.join(GOLDEN, name))
    sd = O.make_synthetic_state_dict_swinv2(fix["sd_name"], fix["sd_seed"])
    cfg, model = _load_model(sd, dtype, name="dpt_swin2_synth.pt")
    assert model.model_type == "swinv2"
    img = fix["img"].to("cuda", dtype)
    report = {}
    with torch.inference_mode():
        tokens, grid_hw = model.patch_embed(img)
        assert tuple(grid_hw) == tuple(fix["grid_hw"])
        report["tokens"] = _err(tokens, fix["tokens"])
        taps = model.imgencoder(fix["tokens"].to("cuda", dtype), grid_hw)
        for i in range(4):
            assert tuple(taps[i].shape) == tuple(fix["taps"][i].shape)
            report[f"tap{i}"] = _err(taps[i], fix["taps"][i])
        maps = model.reassemble(*[t.to("cuda", dtype) for t in fix["taps"]], grid_hw)
        for i in range(4):
            assert tuple(maps[i].shape) == tuple(fix["maps"][i].shape)
            report[f"map{i}"] = _err(maps[i], fix["maps"][i])
        fused = model.fusion(*[t.to("cuda", dtype) for t in fix["maps"]])
        report["fused"] = _err(fused, fix["fused"])
        depth = model.head(fix["fused"].to("cuda", dtype))
        report["head"] = _err(depth, fix["depth"])
        full = model(img)
        assert tuple(full.shape) == tuple(fix["depth"].shape)
        report["depth_e2e"] = _err(full, fix["depth"])
    for k, v in report.items():
        print(f"{name} {dtype} {k}: rel_l2={v[0]:.3e} max_abs={v[1]:.3e} max_rel={v[2]:.3e}")
    # The fixture's logit scales reach the ln(100) clamp: logits of +-100 make the softmax so peaked that the 16-bit
    # rounding of the normalised q/k dominates. The reference's OWN 16-bit CPU forward on these weights is off by
    # 2.2e-2 .. 1.1e-1 (bf16) / 0.9e-2 .. 5e-2 (fp16) on the four taps and 7.2e-2 / 2.5e-2 on the depth map
    # (measured with the reference in the build container); the encoder gates are set below those figures.
    enc_tol = {torch.bfloat16: 0.12, torch.float16: 0.02}[dtype]
    e2e_tol = {torch.bfloat16: 0.08, torch.float16: 0.02}[dtype]
    for k, v in report.items():
        tol = enc_tol if k.startswith("tap") else (e2e_tol if k == "depth_e2e" else REL_L2[dtype])
        gate(f"swin.{name}.{_dt(dtype)}.{k}.rel_l2", v[0], tol)


def test_swinv2_micro_mild_logit_scale_tight_tolerance():
    """same architecture with logit scales around 10 (no clamp): the usual per-stage tolerances apply, which is what
    would catch a wrong window / shift / mask / bias index"""
    from oracle import dpt_oracle as O

    sd = O.make_synthetic_state_dict_swinv2("swinv2_micro", seed=21, logit_std=0.3)
    for shape in ((2, 128, 128), (1, 128, 160)):
        img = O.make_input(*shape, seed=5)
        st = O.forward_swinv2(sd, img, return_stages=True)
        cfg, model = _load_model(sd, torch.float16, name="dpt_swin2_mild.pt")
        with torch.inference_mode():
            tokens, grid_hw = model.patch_embed(img.to("cuda", torch.float16))
            taps = model.imgencoder(st["tokens"].to("cuda", torch.float16), grid_hw)
            depth = model(img.to("cuda", torch.float16))
        for i in range(4):
            e = _err(taps[i], st["taps"][i])
            print(f"swinv2 mild {shape} tap{i}: rel_l2={e[0]:.3e}")
            gate(f"swin.mild.{shape[1]}x{shape[2]}.fp16.tap{i}.rel_l2", e[0], 2 * REL_L2[torch.float16])
        e = _err(depth, st["depth"])
        print(f"swinv2 mild {shape} depth: rel_l2={e[0]:.3e} max_rel={e[2]:.3e}")
        gate(f"swin.mild.{shape[1]}x{shape[2]}.fp16.depth.rel_l2", e[0], 2 * REL_L2[torch.float16])
        gate(f"swin.mild.{shape[1]}x{shape[2]}.fp16.depth.max_rel", e[2], DEPTH_MAXREL[torch.float16])


def test_swinv2_tiny_256_against_oracle():
    """config W shape family (BASELINE.json configs[3]): windowed cosine attention, here the SwinV2-T architecture
    (window 16, heads 3/6/12/24 incl. an odd head count) at 256x256, batch 2, fp16, vs the fp32 oracle"""
    from oracle import dpt_oracle as O

    sd = O.make_synthetic_state_dict_swinv2("swinv2_tiny_256", seed=6)
    img = O.make_input(2, 256, 256, seed=8)
    ref = O.forward_swinv2(sd, img)
    cfg, model = _load_model(sd, torch.float16, name="dpt_swin2_tiny_256.pt")
    with torch.inference_mode():
        depth = model(img.to("cuda", torch.float16))
    e = _err(depth, ref)
    print(f"swinv2_tiny_256 fp16: rel_l2={e[0]:.3e} max_abs={e[1]:.3e} max_rel={e[2]:.3e}")
    assert tuple(depth.shape) == (2, 256, 256)
    gate("swin.tiny256.fp16.depth.rel_l2", e[0], 0.03)  # clamp-reaching logit scales, see the note in the micro test
    gate("swin.tiny256.fp16.depth.max_rel", e[2], 0.06)


def test_pipelined_host_forward_matches_synchronous_one():
    """dpt_forward_host_async on two alternating device buffer pairs (copies overlap the other slot's forward): every
    step's result equals the synchronous dpt_forward_host result of the same input, bit for bit"""
    from oracle import dpt_oracle as O

    sd = O.make_synthetic_state_dict("vits", seed=11)
    cfg, model = _load_model(sd, torch.bfloat16)
    B, H, W = 2, 252, 196
    imgs = [O.make_input(B, H, W, seed=40 + i).to(torch.bfloat16).pin_memory() for i in range(6)]
    sync_out = []
    for x in imgs:
        o = torch.empty(B, H, W, dtype=torch.bfloat16).pin_memory()
        model.forward_host(x, o)
        sync_out.append(o.clone())
    outs = [torch.empty(B, H, W, dtype=torch.bfloat16).pin_memory() for _ in imgs]
    pending = [None, None]
    for i, x in enumerate(imgs):
        k = i & 1
        if pending[k] is not None:
            pending[k].synchronize()
        pending[k] = model.forward_host_async(x, outs[i], slot=k)
    for ev in pending:
        ev.synchronize()
    for i in range(len(imgs)):
        assert torch.equal(outs[i], sync_out[i]), i
    with pytest.raises(ValueError):
        model.forward_host_async(imgs[0], torch.empty(B, H, W + 1, dtype=torch.bfloat16))
