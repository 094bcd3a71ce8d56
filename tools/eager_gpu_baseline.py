#!/usr/bin/env python3
"""What stock PyTorch eager gives on the same B200 (SURVEY.md section 8d "secondary on-box baseline"): the oracle port
of the reference path (identical ATen calls: conv2d / linear / scaled_dot_product_attention / layer_norm / interpolate)
moved to the GPU in bf16, ViT-L, 504x504. The reference itself cannot travel to the GPU box; the oracle reproduces it
bit-exactly on CPU. A measurement aid for DESIGN.md, not part of the product or of bench.py.
usage (GPU box): python tools/eager_gpu_baseline.py [model] [batch]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dpt_oracle as O  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "vitl"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dt = torch.bfloat16
sd = {k: (v.to("cuda", dt) if v.is_floating_point() else v.cuda()) for k, v in O.make_synthetic_state_dict(name, seed=11).items()}
cfg = O.infer_config(sd)
img = O.make_input(B, 504, 504, seed=2).to("cuda", dt)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
with torch.inference_mode():
    for _ in range(3):
        out = O.forward(sd, img, cfg=cfg)
    torch.cuda.synchronize()
    ms = []
    for _ in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = O.forward(sd, img, cfg=cfg)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
ms.sort()
med = ms[len(ms) // 2]
print(json.dumps({"what": "oracle port through stock torch eager on the GPU (bf16)", "model": name, "batch": B, "size": 504,
                  "ms_per_step": med, "frames_per_s": B / med * 1e3, "torch": torch.__version__}))
