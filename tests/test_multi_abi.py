"""ctc_mesh_spans_multi: the span scheduler behind the C ABI (one process, one worker thread per GPU).

CPU: the sharding / region tables (ctc_multi_shard_plan touches no GPU).  GPU: the sharded result must
equal the single-context result span by span, byte by byte -- with one device and with every device of
the box; overflow must report a capacity that makes the retry succeed."""
import ctypes as C

import numpy as np
import pytest

import cantucci_b200 as cb
from cantucci_b200 import _lib
from conftest import startup_leaves


# ------------------------------------------------------------------ CPU: sharding tables ---------------
@pytest.mark.parametrize("nspans,ngpus", [(0, 1), (1, 4), (7, 2), (64, 8), (4096, 8), (1000, 3)])
def test_shard_plan_tables(nspans, ngpus):
    vcap, icap = 1_000_003, 6_000_011
    first_v, first_i, owner = cb.shard_plan(nspans, ngpus, vcap, icap)
    assert first_v[0] == 0 and first_i[0] == 0
    assert np.all(np.diff(first_v.astype(np.int64)) >= 0) and first_v[-1] <= vcap and first_i[-1] <= icap
    assert np.all(first_v % 64 == 0) and np.all(first_i % 64 == 0)          # 16-byte aligned regions
    assert np.array_equal(owner, np.arange(nspans) % ngpus)                 # round-robin deal
    if nspans:
        counts = np.bincount(owner, minlength=ngpus)
        share_v = np.diff(first_v.astype(np.int64))
        # regions are proportional to span counts (up to the 64-element alignment)
        assert np.all(np.abs(share_v - vcap * counts / nspans) <= 64)
        # every span-owning device has room when the capacity is not tiny
        assert np.all(share_v[counts > 0] > 0)


def test_shard_plan_rejects_bad_arguments():
    with pytest.raises(cb.CantucciError):
        cb.shard_plan(10, 0, 100, 100)


def test_multi_symbols_are_exported():
    L = cb.lib()
    for name in ("ctc_multi_create", "ctc_multi_destroy", "ctc_mesh_spans_multi", "ctc_mesh_spans_multi_device",
                 "ctc_multi_shard_plan", "ctc_multi_ngpus", "ctc_multi_ctx", "ctc_multi_last_error"):
        assert hasattr(L, name)


# ------------------------------------------------------------------ GPU --------------------------------
def _assert_same(multi_batch, ref):
    assert len(multi_batch) == len(ref)
    for k in range(len(ref)):
        a, b = multi_batch.mesh(k), ref.mesh(k)
        assert np.array_equal(a.indices, b.indices), k
        assert np.array_equal(a.vertices.view(np.uint32), b.vertices.view(np.uint32)), k


@pytest.mark.gpu
@pytest.mark.parametrize("fast", [False, True])
def test_multi_on_one_device_equals_single_context(ctx, fast):
    spans = startup_leaves()
    shape = cb.Mandelbulb.classic(6, 2.5, fast=fast)
    ref, tr = cb.generate_for_boxes(spans, shape, 32, ctx)
    m = cb.MultiContext(ngpus=1)
    try:
        got, t = cb.generate_for_boxes_multi(spans, shape, 32, m)
        _assert_same(got, ref)
        assert t.vertices == tr.vertices and t.faces == tr.faces
        # empty call, and a call smaller than the device count
        empty, _ = cb.generate_for_boxes_multi(np.zeros((0, 6), np.float32), shape, 32, m)
        assert len(empty) == 0
    finally:
        m.close()


@pytest.mark.gpu
def test_multi_on_every_device_equals_single_context(ctx):
    n = cb.lib().ctc_device_count()
    shape = cb.Mandelbulb.classic(6, 2.5)
    spans = cb.tile_volume(shape.bounding_box(), 6)           # 216 spans
    ref, _ = cb.generate_for_boxes(spans, shape, 32, ctx)
    m = cb.MultiContext(ngpus=n)
    try:
        assert m.ngpus == n
        got, t = cb.generate_for_boxes_multi(spans, shape, 32, m)
        _assert_same(got, ref)
        assert t.vertices == len(ref.vertices)
        # fewer spans than devices: some devices get nothing
        few, _ = cb.generate_for_boxes_multi(spans[100:101], shape, 32, m)
        _assert_same(few, cb.generate_for_boxes(spans[100:101], shape, 32, ctx)[0])
    finally:
        m.close()


@pytest.mark.gpu
def test_multi_device_destination_and_overflow_retry(ctx):
    """Destination in device memory of devices[0] (the NVLink gather target), and the overflow contract."""
    import torch
    n = cb.lib().ctc_device_count()
    shape = cb.Mandelbulb.classic(6, 2.5)
    sh = shape._ctc_shape()
    spans = cb.tile_volume(shape.bounding_box(), 4)           # 64 spans
    ref, _ = cb.generate_for_boxes(spans, shape, 32, ctx)
    ns = len(spans)
    m = cb.MultiContext(ngpus=n)
    try:
        span_v = np.zeros((ns, 2), np.uint64); span_i = np.zeros((ns, 2), np.uint64)
        need = (C.c_uint64 * 2)()
        dv = torch.empty((64, 7), dtype=torch.float32, device="cuda:0")
        di = torch.empty((384,), dtype=torch.int32, device="cuda:0")
        rc = cb.lib().ctc_mesh_spans_multi_device(m.handle, C.byref(sh), spans.ctypes.data, ns, 32, dv.data_ptr(), 64,
                                                  di.data_ptr(), 384, span_v.ctypes.data, span_i.ctypes.data, need, None)
        assert rc == _lib.CTC_ERR_OVERFLOW and need[0] >= len(ref.vertices) and need[1] >= len(ref.indices)
        dv = torch.empty((int(need[0]), 7), dtype=torch.float32, device="cuda:0")
        di = torch.empty((int(need[1]),), dtype=torch.int32, device="cuda:0")
        t = _lib.CtcTimings()
        rc = cb.lib().ctc_mesh_spans_multi_device(m.handle, C.byref(sh), spans.ctypes.data, ns, 32, dv.data_ptr(), int(need[0]),
                                                  di.data_ptr(), int(need[1]), span_v.ctypes.data, span_i.ctypes.data, need,
                                                  C.byref(t))
        m.check(rc)
        torch.cuda.synchronize()
        got = cb.MultiMeshBatch(dv.cpu().numpy().view(cb.VERTEX_DTYPE).reshape(-1), di.cpu().numpy().view(np.uint32), span_v, span_i)
        _assert_same(got, ref)
        assert t.vertices == len(ref.vertices) and t.faces * 6 == len(ref.indices)
    finally:
        m.close()
