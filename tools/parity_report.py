#!/usr/bin/env python3
"""Per-stage parity of the CUDA path against the fp32 CPU oracle at the benchmark's image size (504x504, B=1) for the
Depth-Anything-V2 ViT-S / ViT-B / ViT-L architectures (synthetic seeded checkpoints), in bf16 and fp16: relative L2,
max-abs and max-abs / max|ref| per stage (stages run end to end, so errors compound as they do in a real forward).
Writes one JSON document. usage (GPU box): python tools/parity_report.py [out.json] [models...]"""
import json
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from muggled_dpt_b200 import make_dpt_from_state_dict  # noqa: E402
from oracle import dpt_oracle as O  # noqa: E402  (the checker)

out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "parity_report.json")
models = sys.argv[2:] or ["vits", "vitb", "vitl"]


def err(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    d = (a - b).abs()
    return {"rel_l2": ((a - b).norm() / b.norm().clamp_min(1e-12)).item(), "max_abs": d.max().item(),
            "max_rel": d.max().item() / (b.abs().max().item() + 1e-12)}


report = {}
for name in models:
    sd = O.make_synthetic_state_dict(name, seed=11)
    img = O.make_input(1, 504, 504, seed=2)
    ref = O.forward(sd, img, return_stages=True)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, f"depth_anything_v2_{name}.pth")
        torch.save(sd, path)
        _, model = make_dpt_from_state_dict(path)
    for dtype in (torch.bfloat16, torch.float16):
        model.to(device="cuda", dtype=dtype)
        with torch.inference_mode():
            x = img.to("cuda", dtype)
            tokens, grid = model.patch_embed(x)
            taps = model.imgencoder(tokens, grid)
            maps = model.reassemble(*taps, grid)
            fused = model.fusion(*maps)
            depth = model.head(fused)
            whole = model(x)
        assert torch.equal(depth, whole)
        r = {"tokens": err(tokens, ref["tokens"])}
        for i in range(4):
            r[f"tap{i}"] = err(taps[i], ref["taps"][i])
        for i in range(4):
            r[f"map{i}"] = err(maps[i], ref["maps"][i])
        r["fused"] = err(fused, ref["fused"])
        r["depth"] = err(depth, ref["depth"])
        report[f"{name}_{str(dtype).split('.')[-1]}"] = r
        print(name, dtype, "depth rel_l2 %.2e max_abs %.2e max_rel %.2e | worst stage rel_l2 %.2e" % (
            r["depth"]["rel_l2"], r["depth"]["max_abs"], r["depth"]["max_rel"], max(v["rel_l2"] for v in r.values())), flush=True)
os.makedirs(os.path.dirname(out_path), exist_ok=True)
json.dump({"image": "1x3x504x504 N(0,1), seed 2", "reference": "oracle/dpt_oracle.py fp32 CPU (bit-exact to the reference on the golden fixtures)",
           "stages": report}, open(out_path, "w"), indent=1)
