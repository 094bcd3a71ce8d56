"""-m gpu: the five BASELINE.json configurations at their REAL architectures and image sizes (S: ViT-S 504^2, B: ViT-B
504^2, L: ViT-L 504^2, W: SwinV2-L 384^2 - window 24, 576-token windows, 48 heads of 32 - and E: BEiT-L 384^2), every
stage of the CUDA path (run end to end, through the reference-shaped Python surface and the C ABI) against the fp32 CPU
oracle: relative L2, and max-abs / max|ref| ("max-rel") per stage as the north star words it. Batch is reduced so the
oracle finishes in seconds (frames of a batch never interact; batch independence is tested in test_model_gpu.py).
A sixth case, G, is the Depth-Anything-V2 ViT-Giant at its real dimensions (1536 features, 40 blocks, 24 heads, SwiGLU
FFN, 1536-channel reassembly, 384 fusion channels - make_depthanythingv2_dpt.py:88-95) on a 224x308 frame.

Two bars per (configuration, dtype):
  * every stage stays below its pinned gate (<= 1.5 x the value measured on B200, tests/golden/parity_gates.json);
  * the depth map is at least as close to the fp32 result as the REFERENCE'S OWN 16-bit forward is
    (tests/golden/reference_16bit_error.json, produced by oracle/make_ref16_errors.py with the unmodified reference).
"""
import json
import os
import tempfile

import pytest
import torch

from gpu_util import gate

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CONFIGS = {
    "S": ("vits", 1, 504, "dav2"), "B": ("vitb", 2, 504, "dav2"), "L": ("vitl", 1, 504, "dav2"),
    "W": ("swinv2_large_384", 1, 384, "swin"), "E": ("beit_large_384", 1, 384, "beit"),
    "G": ("vitg", 1, (224, 308), "dav2"),
}
# loose defaults for gates that are not pinned yet (rel_l2 per stage, max-rel per stage)
DEFAULT = {"dav2": {torch.bfloat16: (5e-2, 6e-2), torch.float16: (6e-3, 8e-3)},
           "beit": {torch.bfloat16: (1.5e-2, 3e-2), torch.float16: (3e-3, 5e-3)},
           "swin": {torch.bfloat16: (1.2e-1, 1.5e-1), torch.float16: (2e-2, 3e-2)}}


def _err(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    d = (a - b).abs()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item(), d.max().item() / (b.abs().max().item() + 1e-12)


@pytest.mark.parametrize("key", ["S", "B", "L", "W", "E", "G"])
def test_baseline_config_every_stage_against_oracle(key):
    from muggled_dpt_b200 import make_dpt_from_state_dict
    from oracle import dpt_oracle as O

    name, B, S, fam = CONFIGS[key]
    H, W = S if isinstance(S, tuple) else (S, S)
    if fam == "dav2":
        sd, fwd, fname = O.make_synthetic_state_dict(name, seed=11), O.forward, f"depth_anything_v2_{name}.pth"
        if name == "vitg":
            sd = O.giantify(sd, seed=11)
    elif fam == "beit":
        sd, fwd, fname = O.make_synthetic_state_dict_beit(name, seed=11), O.forward_beit, f"dpt_{name}.pt"
    else:
        sd, fwd, fname = O.make_synthetic_state_dict_swinv2(name, seed=11), O.forward_swinv2, f"dpt_{name}.pt"
    img = O.make_input(B, H, W, seed=2)
    ref = fwd(sd, img, return_stages=True)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, fname)
        torch.save(sd, path)
        del sd
        cfg, model = make_dpt_from_state_dict(path)
    ref16_path = os.path.join(GOLDEN, "reference_16bit_error.json")
    ref16 = json.load(open(ref16_path)) if os.path.exists(ref16_path) else {}
    for dtype in (torch.bfloat16, torch.float16):
        dt = "bf16" if dtype == torch.bfloat16 else "fp16"
        model.to(device="cuda", dtype=dtype)
        with torch.inference_mode():
            x = img.to("cuda", dtype)
            tokens, grid = model.patch_embed(x)
            taps = model.imgencoder(tokens, grid)
            maps = model.reassemble(*taps, grid)
            fused = model.fusion(*maps)
            depth = model.head(fused)
            whole = model(x)
        assert tuple(depth.shape) == (B, H, W) and torch.equal(depth, whole)
        stages = {"tokens": (tokens, ref["tokens"]), "fused": (fused, ref["fused"]), "depth": (depth, ref["depth"])}
        for i in range(4):
            stages[f"tap{i}"] = (taps[i], ref["taps"][i])
            stages[f"map{i}"] = (maps[i], ref["maps"][i])
        d_l2, d_mr = DEFAULT[fam][dtype]
        errs = {}
        for sname, (a, b) in stages.items():
            assert tuple(a.shape) == tuple(b.shape), (sname, a.shape, b.shape)
            errs[sname] = _err(a, b)
            gate(f"config{key}.{name}.{dt}.{sname}.rel_l2", errs[sname][0], d_l2)
            gate(f"config{key}.{name}.{dt}.{sname}.max_rel", errs[sname][1], d_mr)
        r16 = ref16.get(f"{key}:{name}_B{B}_{S}_{'bfloat16' if dtype == torch.bfloat16 else 'float16'}")
        if r16 is not None:
            print(f"config {key} {dt}: depth rel_l2 {errs['depth'][0]:.2e} / max-rel {errs['depth'][1]:.2e}; the reference's own "
                  f"{dt} CPU forward: {r16['rel_l2']:.2e} / {r16['max_rel']:.2e}")
            assert errs["depth"][0] <= r16["rel_l2"], (errs["depth"], r16)
            assert errs["depth"][1] <= r16["max_rel"], (errs["depth"], r16)
