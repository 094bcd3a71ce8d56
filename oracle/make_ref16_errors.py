#!/usr/bin/env python3
"""The reference's OWN 16-bit error: the unmodified reference (imported from /root/reference in the build container)
run on the CPU in bf16 / fp16 against its fp32 forward, same synthetic checkpoints and inputs as tools/parity_report.py
(the five BASELINE.json configurations). Writes tests/golden/reference_16bit_error.json, which the -m gpu parity tests
use as a ceiling: the CUDA path must be at least as close to fp32 as the reference's own 16-bit path is.

TEST INFRASTRUCTURE ONLY. usage: python oracle/make_ref16_errors.py [config letters, default SBLWE]"""
import json
import os
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
from muggled_dpt.make_dpt import make_dpt_from_state_dict  # noqa: E402  (the real reference)
from oracle import dpt_oracle as O  # noqa: E402

CONFIGS = {
    "S": ("vits", 1, 504, "dav2"), "B": ("vitb", 2, 504, "dav2"), "L": ("vitl", 1, 504, "dav2"),
    "W": ("swinv2_large_384", 1, 384, "swin"), "E": ("beit_large_384", 1, 384, "beit"),
}
OUT = os.path.join(ROOT, "tests", "golden", "reference_16bit_error.json")


def err(a, b):
    a, b = a.float(), b.float()
    d = (a - b).abs()
    return {"rel_l2": ((a - b).norm() / b.norm().clamp_min(1e-12)).item(), "max_abs": d.max().item(),
            "max_rel": d.max().item() / (b.abs().max().item() + 1e-12)}


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "SBLWE"
    report = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for key in which:
        name, B, S, fam = CONFIGS[key]
        if fam == "dav2":
            sd, fname = O.make_synthetic_state_dict(name, seed=11), f"depth_anything_v2_{name}.pth"
        elif fam == "beit":
            sd, fname = O.make_synthetic_state_dict_beit(name, seed=11), f"dpt_{name}.pt"
        else:
            sd, fname = O.make_synthetic_state_dict_swinv2(name, seed=11), f"dpt_{name}.pt"
        img = O.make_input(B, S, S, seed=2)
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, fname)
            torch.save(sd, path)
            _, model = make_dpt_from_state_dict(path)
        with torch.inference_mode():
            ref = model(img)
            for dtype in (torch.bfloat16, torch.float16):
                t0 = time.time()
                model.to(dtype=dtype)
                out = model(img.to(dtype))
                e = err(out, ref)
                e["seconds"] = time.time() - t0
                report[f"{key}:{name}_B{B}_{S}_{str(dtype).split('.')[-1]}"] = e
                print(key, name, dtype, e, flush=True)
                json.dump(report, open(OUT, "w"), indent=1)
            model.to(dtype=torch.float32)


if __name__ == "__main__":
    main()
