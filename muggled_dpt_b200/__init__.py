"""muggled_dpt_b200 - B200-native (sm_100a) implementation of muggled_dpt's single-image depth inference hot path."""

from .make_dpt import (  # noqa: F401
    make_beit_dpt_from_midas_v31_state_dict,
    make_depthanythingv1_dpt_from_original_state_dict,
    make_depthanythingv2_dpt_from_original_state_dict,
    make_dpt_from_state_dict,
    make_swinv2_dpt_from_midas_v31_state_dict,
)
from .dpt_model import DPTModel  # noqa: F401

__all__ = ["make_dpt_from_state_dict", "make_depthanythingv2_dpt_from_original_state_dict",
           "make_depthanythingv1_dpt_from_original_state_dict",
           "make_beit_dpt_from_midas_v31_state_dict", "make_swinv2_dpt_from_midas_v31_state_dict", "DPTModel"]
