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


def test_host_side_quad_widening_matches_the_record_definition():
    """ctc_expand_quads_host (what ctc_mesh_spans' host-thread pool runs): a packed 8-byte quad record is four
    16-bit span-local ids v0 < v1 < v2 < v3 with the first two swapped for a flipped winding; widened it is
    [v0,v1,v2, v1,v3,v2] or [v0,v2,v1, v1,v2,v3] (buffer.rs:310-323).  Every destination alignment, so that the
    vector body (aligned non-temporal stores), its scalar head and its tail are all exercised.  No GPU involved."""
    import ctypes as C
    import numpy as np
    from cantucci_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(7)
    n = 10_007
    ids = np.sort(rng.choice(65536, size=(n, 4), replace=True).astype(np.uint32), axis=1)
    ids[:, 1] += (ids[:, 0] == ids[:, 1])            # v0 < v1 strictly (the record's winding bit needs it)
    ids = np.sort(np.minimum(ids, 65535), axis=1)
    ids[ids[:, 0] == ids[:, 1], 0] = 0
    ids[ids[:, 1] == 0, 1] = 1
    flip = rng.integers(0, 2, size=n).astype(bool)
    a = np.where(flip, ids[:, 1], ids[:, 0]); b = np.where(flip, ids[:, 0], ids[:, 1])
    rec = np.empty((n, 2), dtype=np.uint32)
    rec[:, 0] = a | (b << 16); rec[:, 1] = ids[:, 2] | (ids[:, 3] << 16)
    v0, v1, v2, v3 = ids.T
    want = np.where(flip[:, None], np.stack([v0, v2, v1, v1, v2, v3], 1), np.stack([v0, v1, v2, v1, v3, v2], 1)).astype(np.uint32)
    for off in range(0, 9):                          # destination offsets in u32 elements: every 32-byte phase
        for cnt in (0, 1, 3, 4, 5, 64, n):
            buf = np.full(6 * n + 32, 0xDEADBEEF, dtype=np.uint32)
            out = buf[off: off + 6 * cnt]
            assert L.ctc_expand_quads_host(rec.ctypes.data, cnt, buf.ctypes.data + 4 * off) == _lib.CTC_OK
            assert np.array_equal(out, want[:cnt].reshape(-1)), (off, cnt)
            assert np.all(buf[:off] == 0xDEADBEEF) and np.all(buf[off + 6 * cnt:] == 0xDEADBEEF), (off, cnt)   # nothing written outside
