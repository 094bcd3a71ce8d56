"""
make_dpt_from_state_dict - same signature and return value as the reference factory
(muggled_dpt/make_dpt.py:21-72): loads an upstream checkpoint, sniffs the model type, infers the config from the
tensor shapes and returns (config_dict, model). The model it returns is the B200-native DPTModel; move it to the GPU
with `model.to(device="cuda", dtype=torch.bfloat16)` exactly like the reference demos do (run_image.py:158).
"""

from __future__ import annotations

from time import sleep

import torch

from .dpt_model import DPTModel
from .weights import (
    determine_model_type_from_state_dict,
    get_model_config_from_midas_beit_state_dict,
    get_model_config_from_midas_swinv2_state_dict,
    get_model_config_from_state_dict,
    get_model_config_from_v1_state_dict,
)


def make_dpt_from_state_dict(
    path_to_state_dict: str,
    enable_cache: bool = False,
    enable_optimizations: bool = True,
    strict_load: bool = True,
    model_type: str | None = None,
) -> tuple[dict, DPTModel]:
    # weights are packed on the host first, so always load to CPU (reference: cuda load with cpu fallback, :38-41)
    state_dict = torch.load(path_to_state_dict, map_location="cpu")

    if model_type is None:
        model_type = determine_model_type_from_state_dict(path_to_state_dict, state_dict)

    known_model_types = ["swinv2", "beit", "depthanythingv1", "depthanythingv2"]
    if model_type not in known_model_types:
        print("Accepted model types:", *known_model_types, sep="\n")
        raise NotImplementedError(f"Bad model type: {model_type}, no support for this yet!")
    if model_type == "beit":
        return make_beit_dpt_from_midas_v31_state_dict(state_dict, enable_cache, enable_optimizations, strict_load)
    if model_type == "swinv2":
        return make_swinv2_dpt_from_midas_v31_state_dict(state_dict, enable_cache, enable_optimizations, strict_load)
    if model_type == "depthanythingv1":
        return make_depthanythingv1_dpt_from_original_state_dict(state_dict, enable_cache, enable_optimizations, strict_load)

    # metric models are indistinguishable by weights; the reference keys off the file name (make_dpt.py:56-66)
    if model_type == "depthanythingv2" and "metric" in path_to_state_dict:
        state_dict["is_metric"] = torch.tensor((1), dtype=torch.float32)
        print("", "Warning: Metric Depth-Anything V2 model detected!", "  These models are not officially supported,",
              "  model outputs may be incorrect...", sep="\n", flush=True)
        sleep(1.5)

    return make_depthanythingv2_dpt_from_original_state_dict(state_dict, enable_cache, enable_optimizations, strict_load)


def make_depthanythingv2_dpt_from_original_state_dict(
    state_dict: dict,
    enable_cache: bool = False,
    enable_optimizations: bool = True,
    strict_load: bool = True,
) -> tuple[dict, DPTModel]:
    """make_depthanythingv2_dpt.py:24-61. enable_cache / enable_optimizations are accepted for signature parity: the
    position table is always precomputed per grid and the fused attention kernel is the only attention path."""
    if not strict_load:
        print("", "WARNING:", "  Loading model weights without 'strict' mode enabled!",
              "  Some weights may be missing or unused!", sep="\n", flush=True)
    config_dict = get_model_config_from_state_dict(state_dict, enable_cache, enable_optimizations)
    model = DPTModel(config_dict, state_dict, strict_load=strict_load)
    return config_dict, model


def make_depthanythingv1_dpt_from_original_state_dict(
    state_dict: dict,
    enable_cache: bool = False,
    enable_optimizations: bool = True,
    strict_load: bool = True,
) -> tuple[dict, DPTModel]:
    """make_depthanythingv1_dpt.py:24-61. Same weights schema and kernels as V2; the encoder taps are the outputs of
    the last four blocks (v1_depthanything/image_encoder_model.py:92-103) and there is no metric / giant variant."""
    if not strict_load:
        print("", "WARNING:", "  Loading model weights without 'strict' mode enabled!",
              "  Some weights may be missing or unused!", sep="\n", flush=True)
    config_dict = get_model_config_from_v1_state_dict(state_dict, enable_cache, enable_optimizations)
    model = DPTModel(config_dict, state_dict, strict_load=strict_load, model_type="depthanythingv1")
    return config_dict, model


def make_beit_dpt_from_midas_v31_state_dict(
    midas_v31_state_dict: dict,
    enable_cache: bool = False,
    enable_optimizations: bool = True,
    strict_load: bool = True,
) -> tuple[dict, DPTModel]:
    """make_beit_dpt.py:24-58. The relative position bias is always rebuilt per grid size on the device (the
    reference's optional cache, v31_beit/image_encoder_model.py:93-119, has no observable effect on results)."""
    if not strict_load:
        print("", "WARNING:", "  Loading model weights without 'strict' mode enabled!",
              "  Some weights may be missing or unused!", sep="\n", flush=True)
    config_dict = get_model_config_from_midas_beit_state_dict(midas_v31_state_dict, enable_cache, enable_optimizations)
    model = DPTModel(config_dict, midas_v31_state_dict, strict_load=strict_load, model_type="beit")
    return config_dict, model


def make_swinv2_dpt_from_midas_v31_state_dict(
    midas_v31_state_dict: dict,
    enable_cache: bool = False,
    enable_optimizations: bool = True,
    strict_load: bool = True,
) -> tuple[dict, DPTModel]:
    """make_swinv2_dpt.py:24-58. Bias tables / shift masks are rebuilt on the device for every layer and grid size."""
    if not strict_load:
        print("", "WARNING:", "  Loading model weights without 'strict' mode enabled!",
              "  Some weights may be missing or unused!", sep="\n", flush=True)
    config_dict = get_model_config_from_midas_swinv2_state_dict(midas_v31_state_dict, enable_cache, enable_optimizations)
    model = DPTModel(config_dict, midas_v31_state_dict, strict_load=strict_load, model_type="swinv2")
    return config_dict, model
