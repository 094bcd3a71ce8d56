"""
Device post-processing of a depth prediction - the reference's demo_helpers/postprocess.py:22-102 chain
(scale_prediction -> normalize_01 -> convert_to_uint8) as it is used right after the model in run_image.py:185-195 and
run_video.py:347-350, in two small CUDA kernels behind dpt_postprocess_u8 (include/dpt_b200.h). CUDA-only, like the
model: there is no CPU path.
"""

from __future__ import annotations

import ctypes as C

import torch

from . import _native as N

_TORCH_TO_DPT = {torch.float16: N.DPT_F16, torch.bfloat16: N.DPT_BF16}


def scale_normalize_to_uint8(prediction_bhw: torch.Tensor, target_wh: tuple[int, int]) -> torch.Tensor:
    """convert_to_uint8(scale_prediction(prediction, target_wh)) of demo_helpers/postprocess.py: bilinear resize of the
    BxHxW prediction to (w, h), min/max normalisation over the whole scaled tensor, 0..255 uint8 (on the device)."""
    if not prediction_bhw.is_cuda:
        raise RuntimeError("muggled_dpt_b200.postprocess has no CPU fallback: pass the model's device tensor")
    if prediction_bhw.dtype not in _TORCH_TO_DPT or prediction_bhw.dim() != 3:
        raise ValueError("expected a BxHxW bf16 / fp16 prediction (the model's output)")
    pred = prediction_bhw.contiguous()
    B, H, W = pred.shape
    OW, OH = int(target_wh[0]), int(target_wh[1])
    with torch.cuda.device(pred.device):
        out = torch.empty((B, OH, OW), device=pred.device, dtype=torch.uint8)
        minmax = torch.empty(2, device=pred.device, dtype=torch.float32)
        stream = C.c_void_p(torch.cuda.current_stream(pred.device).cuda_stream)
        N.check(
            N.lib().dpt_postprocess_u8(C.c_void_p(pred.data_ptr()), B, H, W, C.c_void_p(out.data_ptr()), OH, OW,
                                       C.c_void_p(minmax.data_ptr()), _TORCH_TO_DPT[pred.dtype], stream),
            None, "dpt_postprocess_u8",
        )
    return out
