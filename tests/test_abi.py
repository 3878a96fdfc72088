"""The C-ABI library loads and exports every symbol include/cantucci_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "cantucci_b200.h")).read()
    return sorted(set(re.findall(r"CTC_API[^;(]*?\b(ctc_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    fns = header_functions()
    for must in ("ctc_ctx_create", "ctc_de_batch", "ctc_sample_grids", "ctc_mesh_spans",
                 "ctc_mesh_spans_device", "ctc_mesh_result"):
        assert must in fns


def test_library_exports_every_declared_symbol():
    from cantucci_b200 import _lib
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in header_functions():
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.EXPORTS) == header_functions()


def test_struct_layouts_match_the_reference_records():
    from cantucci_b200 import _lib
    assert ctypes.sizeof(_lib.CtcSpan) == 24          # 6 x f32
    assert _lib.VERTEX_DTYPE.itemsize == 28           # mesh/mod.rs:255-261
    assert _lib.VERTEX_DTYPE.fields["normal"][1] == 12
    assert _lib.VERTEX_DTYPE.fields["distance_from_surface"][1] == 24
    assert _lib.lib().ctc_version() == 100


def test_no_cpu_fallback_without_a_device():
    import cantucci_b200 as cb
    if cb.lib().ctc_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(cb.CantucciError) as e:
        cb.Context(0)
    assert e.value.code == 5   # CTC_ERR_NO_DEVICE


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "cantucci_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower() or f == "build.py" and False, f"{f} mentions the oracle"
