#!/usr/bin/env python3
"""Per-stage parity of the CUDA path against the fp32 CPU oracle on the five BASELINE.json configurations at their real
architectures and image sizes (batch reduced so the CPU oracle finishes in seconds; frames of a batch never interact):

  S  Depth-Anything-V2 ViT-S  1x3x504x504  bf16 + fp16        B  ViT-B  2x3x504x504  bf16 + fp16
  L  ViT-L  1x3x504x504  bf16 + fp16                          W  SwinV2-L  1x3x384x384  fp16 (+ bf16)
  E  BEiT-L  1x3x384x384  bf16 (+ fp16)

Stages run end to end (errors compound as in a real forward). Per stage: relative L2, max-abs, max-abs / max|ref|
("max_rel", the north star's wording applied to a map whose small values sit next to a ReLU). Writes one JSON document.
usage (GPU box): python tools/parity_report.py [out.json] [config letters, default SBLWE]"""
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
which = sys.argv[2] if len(sys.argv) > 2 else "SBLWE"

CONFIGS = {
    "S": ("vits", 1, 504, "dav2"), "B": ("vitb", 2, 504, "dav2"), "L": ("vitl", 1, 504, "dav2"),
    "W": ("swinv2_large_384", 1, 384, "swin"), "E": ("beit_large_384", 1, 384, "beit"),
}


def err(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    d = (a - b).abs()
    return {"rel_l2": ((a - b).norm() / b.norm().clamp_min(1e-12)).item(), "max_abs": d.max().item(),
            "max_rel": d.max().item() / (b.abs().max().item() + 1e-12)}


report = {}
for key in which:
    name, B, S, fam = CONFIGS[key]
    if fam == "dav2":
        sd, fwd, fname = O.make_synthetic_state_dict(name, seed=11), O.forward, f"depth_anything_v2_{name}.pth"
    elif fam == "beit":
        sd, fwd, fname = O.make_synthetic_state_dict_beit(name, seed=11), O.forward_beit, f"dpt_{name}.pt"
    else:
        sd, fwd, fname = O.make_synthetic_state_dict_swinv2(name, seed=11), O.forward_swinv2, f"dpt_{name}.pt"
    img = O.make_input(B, S, S, seed=2)
    ref = fwd(sd, img, return_stages=True)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, fname)
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
        report[f"{key}:{name}_B{B}_{S}_{str(dtype).split('.')[-1]}"] = r
        print(key, name, dtype, "depth rel_l2 %.2e max_abs %.2e max_rel %.2e | worst stage rel_l2 %.2e" % (
            r["depth"]["rel_l2"], r["depth"]["max_abs"], r["depth"]["max_rel"], max(v["rel_l2"] for v in r.values())), flush=True)
    del model, ref, sd
os.makedirs(os.path.dirname(out_path), exist_ok=True)
json.dump({"image": "N(0,1) inputs, seed 2; synthetic seeded checkpoints, seed 11",
           "reference": "oracle/dpt_oracle.py fp32 CPU (bit-exact to the reference on the golden fixtures)",
           "stages": report}, open(out_path, "w"), indent=1)
