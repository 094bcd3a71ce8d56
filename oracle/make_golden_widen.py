#!/usr/bin/env python3
"""
Golden fixtures for the rows added after the core path, produced by the UNMODIFIED reference
imported from /root/reference (build container only). Run:  python oracle/make_golden_widen.py

  da_v1_tiny8.pt  - Depth-Anything V1 tap rule: an 8-block synthetic checkpoint loaded through the reference's own
                    factory under a "v1" file name (make_dpt.py:98-104), every stage tensor
  da_v2_giant_tiny.pt - the ViT-G block structure (SwiGLU FFN, detected from the mlp.w12 keys), taps + depth
  da_v2_metric.pt - the metric head (Sigmoid) through the "metric" file-name switch (make_dpt.py:56-66), depth only
  prepare_image.pt - PatchEmbed.prepare_image of the reference on seeded uint8 BGR images (three sizes / models)
  postprocess.pt  - demo_helpers/postprocess.py scale_prediction + convert_to_uint8 on seeded predictions
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import dpt_oracle as O  # noqa: E402
from oracle.make_golden import state_dict_checksum  # noqa: E402


def load_reference(sd, file_name):
    from muggled_dpt.make_dpt import make_dpt_from_state_dict

    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, file_name)
        torch.save(sd, path)
        import time
        sleep, time.sleep = time.sleep, lambda s: None  # the metric warning sleeps 1.5 s
        try:
            cfg, model = make_dpt_from_state_dict(path, enable_cache=False, enable_optimizations=True)
        finally:
            time.sleep = sleep
    return cfg, model.float().eval()


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")

    # ---- Depth-Anything V1 taps
    sd = O.make_synthetic_state_dict("tiny8", seed=21, base_grid=5)
    img = O.make_input(2, 56, 84, seed=9)
    cfg, model = load_reference(sd, "depth_anything_v1_synthetic.pth")
    with torch.inference_mode():
        tokens, grid_hw = model.patch_embed(img)
        taps = model.imgencoder(tokens, grid_hw)
        maps = model.reassemble(*taps, grid_hw)
        fused = model.fusion(*maps)
        depth = model.head(fused)
    torch.save({"model_type": "depthanythingv1", "config": {k: (list(v) if isinstance(v, (list, tuple)) else v) for k, v in cfg.items()},
                "sd_seed": 21, "sd_name": "tiny8", "sd_base_grid": 5, "sd_checksum": state_dict_checksum(sd), "img": img,
                "tokens": tokens, "taps": list(taps), "maps": list(maps), "fused": fused, "depth": depth,
                "grid_hw": tuple(grid_hw)}, os.path.join(out_dir, "da_v1_tiny8.pt"))
    print("da_v1_tiny8: depth std", depth.std().item(), "config keys", list(cfg.keys()))

    # ---- metric head
    sd = O.make_synthetic_state_dict("tiny", seed=4, base_grid=5)
    img = O.make_input(1, 56, 56, seed=2)
    cfg, model = load_reference(dict(sd), "depth_anything_v2_metric_synthetic.pth")
    with torch.inference_mode():
        depth = model(img)
    torch.save({"config": {k: (list(v) if isinstance(v, (list, tuple)) else v) for k, v in cfg.items()}, "sd_seed": 4,
                "sd_name": "tiny", "sd_base_grid": 5, "sd_checksum": state_dict_checksum(sd), "img": img, "depth": depth},
               os.path.join(out_dir, "da_v2_metric.pt"))
    print("da_v2_metric: depth range", depth.min().item(), depth.max().item(), "is_metric", cfg.get("is_metric"))

    # ---- ViT-G structure (SwiGLU FFN): detected from the mlp.w12 keys
    sd = O.giantify(O.make_synthetic_state_dict("tiny", seed=6, base_grid=5), seed=6)
    img = O.make_input(2, 84, 56, seed=4)
    cfg, model = load_reference(sd, "depth_anything_v2_vitg_synthetic.pth")
    assert cfg["is_giant"] is True
    with torch.inference_mode():
        tokens, grid_hw = model.patch_embed(img)
        taps = model.imgencoder(tokens, grid_hw)
        depth = model(img)
    torch.save({"config": {k: (list(v) if isinstance(v, (list, tuple)) else v) for k, v in cfg.items()}, "sd_seed": 6,
                "sd_name": "tiny", "sd_base_grid": 5, "sd_checksum": state_dict_checksum(sd), "img": img,
                "tokens": tokens, "taps": list(taps), "depth": depth, "grid_hw": tuple(grid_hw)},
               os.path.join(out_dir, "da_v2_giant_tiny.pt"))
    print("da_v2_giant_tiny: depth std", depth.std().item(), "taps std", [round(t.std().item(), 3) for t in taps])

    # ---- prepare_image: the reference's own patch-embed modules
    from muggled_dpt.v2_depthanything.patch_embed import PatchEmbed as PE2
    from muggled_dpt.v31_beit.patch_embed import PatchEmbed as PEB
    from muggled_dpt.v31_swinv2.patch_embed import PatchEmbed as PES

    rng = np.random.default_rng(5)
    cases = []
    for name, (h, w), kw in [("da_down", (183, 260), dict(max_side_length=112, use_square_sizing=True)),
                             ("da_aspect", (150, 97), dict(max_side_length=140, use_square_sizing=False)),
                             ("da_up", (40, 52), dict(max_side_length=84, use_square_sizing=True)),
                             ("beit", (120, 200), dict(max_side_length=96, use_square_sizing=True)),
                             ("swinv2", (211, 160), dict(max_side_length=128, use_square_sizing=False))]:
        # smooth image + noise so that both the antialias footprint and sharp pixels matter
        yy, xx = np.mgrid[0:h, 0:w]
        base = np.stack([127 + 100 * np.sin(xx / 9.0 + c) * np.cos(yy / 7.0 - c) for c in range(3)], axis=-1)
        bgr = np.clip(base + rng.normal(0, 25, size=(h, w, 3)), 0, 255).astype(np.uint8)
        if name.startswith("da"):
            pe, mt, patch, grid = PE2(64, 14, 5), "depthanythingv2", 14, 5
        elif name == "beit":
            pe, mt, patch, grid = PEB(64, 16, 6), "beit", 16, 6
        else:
            pe, mt, patch, grid = PES(64, 4, 64), "swinv2", 4, 64
        with torch.inference_mode():
            out = pe.float().prepare_image(bgr, **kw)
        mine = O.prepare_image(bgr, patch, grid, mt, **kw)
        print(f"prepare_image {name}: {bgr.shape} -> {tuple(out.shape)}; oracle max abs diff {(out - mine).abs().max().item():.2e}")
        cases.append({"name": name, "bgr": torch.from_numpy(bgr), "model_type": mt, "patch": patch, "base_grid": grid,
                      "kwargs": kw, "out": out})
    torch.save(cases, os.path.join(out_dir, "prepare_image.pt"))

    # ---- postprocess
    from muggled_dpt.demo_helpers.postprocess import convert_to_uint8, scale_prediction

    g = torch.Generator().manual_seed(3)
    cases = []
    for (B, H, W), wh in [((1, 56, 56), (200, 150)), ((2, 84, 112), (97, 61)), ((1, 42, 70), (70, 42))]:
        pred = (torch.rand(B, H, W, generator=g) * 7.0 + torch.linspace(0, 3, W)[None, None, :]).to(torch.bfloat16).float()
        out = convert_to_uint8(scale_prediction(pred, wh))
        mine = O.postprocess_u8(pred, wh)
        print(f"postprocess {pred.shape} -> {tuple(out.shape)}; oracle equal: {torch.equal(out, mine)}")
        cases.append({"pred": pred.to(torch.bfloat16), "target_wh": wh, "out": out})
    torch.save(cases, os.path.join(out_dir, "postprocess.pt"))


if __name__ == "__main__":
    main()
