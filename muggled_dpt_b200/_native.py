"""ctypes binding of libdpt_b200.so (C ABI: include/dpt_b200.h). There is deliberately no fallback: if the shared
library is missing or fails to load, importing the compute path raises."""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DPT_B200_LIB: an alternative build of the same library (kernel A/B experiments under tools/), never a fallback
LIB_PATH = os.environ.get("DPT_B200_LIB") or os.path.join(_HERE, "lib", "libdpt_b200.so")

DPT_F16, DPT_BF16, DPT_F32 = 0, 1, 2
VARIANT_DINOV2, VARIANT_BEIT, VARIANT_SWINV2 = 0, 1, 2
ACT_NONE, ACT_GELU, ACT_RELU, ACT_SIGMOID = 0, 1, 2, 3

STATUS_NAMES = {0: "DPT_OK", -1: "DPT_ERR_INVALID", -2: "DPT_ERR_MISSING", -3: "DPT_ERR_WORKSPACE",
                -4: "DPT_ERR_CUDA", -5: "DPT_ERR_UNSUPPORTED"}


class DptConfig(C.Structure):
    _fields_ = [
        ("variant", C.c_int),
        ("dtype", C.c_int),
        ("features_per_token", C.c_int),
        ("num_heads", C.c_int),
        ("num_blocks", C.c_int),
        ("reassembly_features", C.c_int * 4),
        ("fusion_channels", C.c_int),
        ("patch_size_px", C.c_int),
        ("base_grid_h", C.c_int),
        ("base_grid_w", C.c_int),
        ("is_metric", C.c_int),
        ("ln_eps", C.c_float),
        ("heads_per_stage", C.c_int * 4),
        ("layers_per_stage", C.c_int * 4),
        ("window_h", C.c_int),
        ("window_w", C.c_int),
        ("pretrained_window", C.c_int * 4),
        ("taps_last4", C.c_int),
        ("mlp_swiglu", C.c_int),
    ]


class NativeLibraryMissing(RuntimeError):
    pass


_lib = None

# every symbol include/dpt_b200.h declares: (name, restype, argtypes)
_VP, _I, _SZ, _F = C.c_void_p, C.c_int, C.c_size_t, C.c_float
_PP4 = C.POINTER(C.c_void_p * 4)
SYMBOLS = [
    ("dpt_create", _I, [C.POINTER(DptConfig), C.POINTER(_VP)]),
    ("dpt_destroy", None, [_VP]),
    ("dpt_last_error", C.c_char_p, [_VP]),
    ("dpt_version", C.c_char_p, []),
    ("dpt_config_size", _I, []),
    ("dpt_set_weight", _I, [_VP, C.c_char_p, _VP, C.POINTER(C.c_int64), _I, _I]),
    ("dpt_workspace_bytes", _I, [_VP, _I, _I, _I, C.POINTER(_SZ)]),
    ("dpt_forward", _I, [_VP, _VP, _VP, _VP, _SZ, _I, _I, _I, _VP]),
    ("dpt_forward_host", _I, [_VP, _VP, _VP, _VP, _VP, _VP, _SZ, _I, _I, _I, _VP]),
    ("dpt_forward_host_async", _I, [_VP, _VP, _VP, _VP, _VP, _VP, _SZ, _I, _I, _I, _VP, _VP, _VP]),
    ("dpt_patch_embed", _I, [_VP, _VP, _VP, _VP, _SZ, _I, _I, _I, _VP]),
    ("dpt_encoder", _I, [_VP, _VP, _PP4, _VP, _SZ, _I, _I, _I, _VP]),
    ("dpt_reassemble", _I, [_VP, _PP4, _PP4, _VP, _SZ, _I, _I, _I, _VP]),
    ("dpt_fusion", _I, [_VP, _PP4, _VP, _VP, _SZ, _I, _I, _I, _VP]),
    ("dpt_head", _I, [_VP, _VP, _VP, _VP, _SZ, _I, _I, _I, _VP]),
    ("dpt_fusion_block", _I, [_VP, _I, _VP, _VP, _VP, _VP, _SZ, _I, _I, _I, _VP]),
    ("dpt_encoder_capture", _I, [_VP, _VP, _PP4, _VP, _VP, _I, _VP, _SZ, _I, _I, _I, _VP]),
    ("dpt_op_conv_gemm", _I, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _VP]),
    ("dpt_op_attention", _I, [_VP, _VP, C.c_int64, _I, _VP, _I, _I, _I, _I, _F, _I, _VP]),
    ("dpt_op_layernorm", _I, [_VP, _VP, _VP, _VP, C.c_int64, _I, _F, _I, _VP]),
    ("dpt_op_resize_bilinear", _I, [_VP, _VP, _I, _I, _I, _I, _I, _I, _I, _VP]),
    ("dpt_prepare_image", _I, [_VP, _I, _I, _VP, _I, _I, C.POINTER(_F * 3), C.POINTER(_F * 3), _I, _VP]),
    ("dpt_postprocess_u8", _I, [_VP, _I, _I, _I, _VP, _I, _I, _VP, _I, _VP]),
    ("dpt_allgather_depth", _I, [_VP, _VP, _VP, _SZ, _I, _VP]),
    ("dpt_op_last_error", C.c_char_p, []),
    ("dpt_last_launch_count", _I, [_VP]),
    ("dpt_profile_enable", _I, [_VP, _I]),
    ("dpt_profile_count", _I, [_VP]),
    ("dpt_profile_get", _I, [_VP, _I, C.c_char_p, _I, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
]


def lib() -> C.CDLL:
    """Loads libdpt_b200.so (once). Raises NativeLibraryMissing - never falls back."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryMissing(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or muggled_dpt_b200/csrc/build.sh). There is no CPU / PyTorch fallback for the depth path."
            )
        handle = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            fn = getattr(handle, name)  # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, handle=None, what: str = "") -> None:
    if rc == 0:
        return
    L = lib()
    msg = (L.dpt_last_error(handle) if handle else L.dpt_op_last_error()) or b""
    text = f"{what}: {STATUS_NAMES.get(rc, rc)}: {msg.decode(errors='replace')}"
    if rc == -1:
        raise ValueError(text)
    raise RuntimeError(text)


def ptr4(tensors) -> C.Array:
    arr = (C.c_void_p * 4)()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr
