#!/usr/bin/env python3
"""Timing of the pre- and post-processing kernels at the reference demo's sizes: a
1920x1080 BGR frame -> 504x504 model input (dpt_prepare_image) and a 504x504 prediction -> 1920x1080 uint8 display map
(dpt_postprocess_u8). CUDA events, L2 flushed between iterations; algorithmic bytes = input once + output once.
Prints one JSON line. usage (GPU box): python tools/bench_prepost.py"""
import json
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from muggled_dpt_b200 import make_dpt_from_state_dict  # noqa: E402
from muggled_dpt_b200.postprocess import scale_normalize_to_uint8  # noqa: E402
from oracle import dpt_oracle as O  # noqa: E402  (synthetic checkpoint generator only)

sd = O.make_synthetic_state_dict("tiny", seed=3, base_grid=36)
with tempfile.TemporaryDirectory() as td:
    path = os.path.join(td, "depth_anything_v2_tiny.pth")
    torch.save(sd, path)
    _, model = make_dpt_from_state_dict(path)
model.to(device="cuda", dtype=torch.bfloat16)
rng = np.random.default_rng(0)
bgr = rng.integers(0, 255, size=(1080, 1920, 3), dtype=np.uint8)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
pred = (torch.rand(1, 504, 504, device="cuda") * 9).to(torch.bfloat16)


def timed(fn, n=20):
    for _ in range(3):
        fn()
    ms = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    return ms[len(ms) // 2]


# prepare_image includes the H2D copy of the raw frame (pageable numpy memory), as the demo's per-frame call does
t_pre = timed(lambda: model.prepare_image_bgr(bgr, max_side_length=504, use_square_sizing=True))
out = model.prepare_image_bgr(bgr, max_side_length=504, use_square_sizing=True)
t_post = timed(lambda: scale_normalize_to_uint8(pred, (1920, 1080)))
pre_bytes = bgr.size + out.numel() * 2
post_bytes = pred.numel() * 2 + 1920 * 1080
print(json.dumps({
    "prepare_image": {"ms": t_pre, "in": list(bgr.shape), "out": list(out.shape), "algorithmic_bytes": pre_bytes,
                      "gbs": pre_bytes / t_pre / 1e6, "includes": "H2D of the 6.2 MB uint8 frame + kernel"},
    "postprocess_u8": {"ms": t_post, "in": list(pred.shape), "out": [1, 1080, 1920], "algorithmic_bytes": post_bytes,
                       "gbs": post_bytes / t_post / 1e6, "includes": "min/max pass + uint8 pass"},
}))
