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

    # the struct the library was compiled with (include/dpt_b200.h) == the ctypes mirror, field for field
    assert ctypes.sizeof(N.DptConfig) == N.lib().dpt_config_size()
    header = open(os.path.join(ROOT, "include", "dpt_b200.h")).read()
    body = re.sub(r"/\*.*?\*/", "", header[header.index("typedef struct dpt_config"):header.index("} dpt_config;")], flags=re.S)
    declared = re.findall(r"\b(?:int|float)\s+([^;]+);", body)
    names = [n.split("[")[0].strip() for d in declared for n in d.split(",")]
    assert names == [f[0] for f in N.DptConfig._fields_]
    # the binding sketch a maintainer would copy from INTEGRATION.md carries the same fields
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    sketch = doc[doc.index("class _Cfg(C.Structure)"):doc.index("_lib.dpt_create.argtypes")]
    assert re.findall(r'\("([a-z0-9_]+)", C\.c_', sketch) == names


def test_allgather_entry_rejects_bad_arguments_without_gpu():
    from muggled_dpt_b200 import _native as N

    L = N.lib()
    assert L.dpt_allgather_depth(None, None, None, 0, N.DPT_BF16, None) == -1  # DPT_ERR_INVALID
    assert b"dpt_allgather_depth" in L.dpt_op_last_error()
