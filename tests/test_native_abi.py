"""CPU: the C-ABI library loads and exports every symbol include/dpt_b200.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "dpt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dpt_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from muggled_dpt_b200 import _native as N

    names = _declared_functions()
    assert len(names) >= 18
    lib = ctypes.CDLL(N.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in dpt_b200.h but not exported"
    bound = {s[0] for s in N.SYMBOLS}
    assert set(names) == bound, (set(names) ^ bound)


def test_version_and_error_strings_without_gpu():
    from muggled_dpt_b200 import _native as N

    L = N.lib()
    assert b"sm_100a" in L.dpt_version()
    assert L.dpt_last_error(None) is not None


def test_config_struct_layout_matches_header():
    from muggled_dpt_b200 import _native as N

    # 14 ints + 1 float + 14 ints (SwinV2 block) + taps_last4 + mlp_swiglu, no padding
    assert ctypes.sizeof(N.DptConfig) == 31 * 4


def test_allgather_entry_rejects_bad_arguments_without_gpu():
    from muggled_dpt_b200 import _native as N

    L = N.lib()
    assert L.dpt_allgather_depth(None, None, None, 0, N.DPT_BF16, None) == -1  # DPT_ERR_INVALID
    assert b"dpt_allgather_depth" in L.dpt_op_last_error()
