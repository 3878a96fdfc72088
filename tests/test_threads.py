"""Shape: Sync + Send -- the reference calls generate_for_box from num_cpus worker threads at once
(mesh/mod.rs:61-62,141).  One shared context (calls serialise on its mutex) and one context per
thread must both give the single-threaded results."""
import threading

import numpy as np
import pytest

from conftest import startup_leaves

pytestmark = pytest.mark.gpu


def test_concurrent_callers_get_identical_meshes(ctx):
    import cantucci_b200 as cb
    spans = startup_leaves()[:16]
    shape = cb.Mandelbulb.classic(6, 2.5)
    want = [cb.MeshBuffer.generate_for_box(cb.Span(tuple(r[:3]), tuple(r[3:])), shape, 32, ctx)[0] for r in spans]
    for shared in (True, False):
        out, errs = [None] * len(spans), []

        def work(k):
            try:
                c = ctx if shared else cb.Context(0)
                out[k] = cb.MeshBuffer.generate_for_box(cb.Span(tuple(spans[k][:3]), tuple(spans[k][3:])), shape, 32, c)[0]
            except Exception as e:       # noqa: BLE001
                errs.append(e)

        th = [threading.Thread(target=work, args=(k,)) for k in range(len(spans))]
        [t.start() for t in th]
        [t.join() for t in th]
        assert not errs, errs
        for a, b in zip(out, want):
            assert np.array_equal(a.indices, b.indices)
            assert np.array_equal(a.vertices.view(np.uint32), b.vertices.view(np.uint32))


def test_concurrent_single_span_calls_are_coalesced(ctx):
    """SURVEY 8b threading bullet: the reference calls generate_for_box from num_cpus worker threads at once
    (mesh/mod.rs:141-148).  On ONE context those calls are coalesced into batched launches: results stay
    those of separate calls, and 16 concurrent single-span calls cost about as much as ONE 16-span call
    plus one single-span call -- not 16 of them."""
    import ctypes as C
    import time
    import cantucci_b200 as cb
    from cantucci_b200 import _lib
    L = _lib.lib()
    spans = startup_leaves()[16:32]
    sh = cb.Mandelbulb.classic(6, 2.5, fast=True)._ctc_shape()
    n = len(spans)
    bufs = [(np.empty(40_000, dtype=cb.VERTEX_DTYPE), np.empty(240_000, dtype=np.uint32), np.zeros(2, np.uint64), np.zeros(2, np.uint64))
            for _ in range(n)]

    def single(k):
        v, i, vo, io = bufs[k]
        rc = L.ctc_mesh_spans(ctx.handle, C.byref(sh), spans[k:k + 1].ctypes.data, 1, 64, v.ctypes.data, len(v), i.ctypes.data, len(i),
                              vo.ctypes.data, io.ctypes.data, None)
        assert rc == 0, ctx.last_error()

    def batched():
        v = np.empty(640_000, dtype=cb.VERTEX_DTYPE); i = np.empty(3_840_000, dtype=np.uint32)
        vo = np.zeros(n + 1, np.uint64); io = np.zeros(n + 1, np.uint64)
        rc = L.ctc_mesh_spans(ctx.handle, C.byref(sh), spans.ctypes.data, n, 64, v.ctypes.data, len(v), i.ctypes.data, len(i),
                              vo.ctypes.data, io.ctypes.data, None)
        assert rc == 0
        return v, i, vo, io

    def concurrent():
        th = [threading.Thread(target=single, args=(k,)) for k in range(n)]
        t0 = time.perf_counter()
        [t.start() for t in th]
        [t.join() for t in th]
        return time.perf_counter() - t0

    ref = batched()
    concurrent()                                           # warm-up (staging buffers)
    b0, r0 = ctx.coalescing_stats()
    t_conc = min(concurrent() for _ in range(5))
    b1, r1 = ctx.coalescing_stats()
    assert b1 > b0 and (r1 - r0) > (b1 - b0)              # batched launches served several requests each
    for k in range(n):                                     # every caller got exactly its span's mesh
        v, i, vo, io = bufs[k]
        a, b = int(ref[2][k]), int(ref[2][k + 1]); c, d = int(ref[3][k]), int(ref[3][k + 1])
        assert int(vo[1]) == b - a and int(io[1]) == d - c
        assert np.array_equal(v[: b - a].view(np.uint32), ref[0][a:b].view(np.uint32))
        assert np.array_equal(i[: d - c], ref[1][c:d])
    t0 = time.perf_counter(); batched(); t_batched = time.perf_counter() - t0
    t0 = time.perf_counter(); single(0); t_single = time.perf_counter() - t0
    # serialised, the 16 calls would cost 16 x t_single; coalesced they cost about one single call + one batch
    # (thread start-up included, hence the slack)
    assert t_conc < 3.0 * (t_batched + t_single) + 2e-3, (t_conc, t_batched, t_single)
    # per-call error contract survives the batching: a caller with too-small buffers gets ITS overflow
    tiny = (np.empty(4, dtype=cb.VERTEX_DTYPE), np.empty(24, dtype=np.uint32), np.zeros(2, np.uint64), np.zeros(2, np.uint64))
    bufs[3] = tiny
    rcs = [None] * n

    def single_rc(k):
        v, i, vo, io = bufs[k]
        rcs[k] = L.ctc_mesh_spans(ctx.handle, C.byref(sh), spans[k:k + 1].ctypes.data, 1, 64, v.ctypes.data, len(v), i.ctypes.data, len(i),
                                  vo.ctypes.data, io.ctypes.data, None)
    th = [threading.Thread(target=single_rc, args=(k,)) for k in range(n)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert rcs[3] == _lib.CTC_ERR_OVERFLOW and int(tiny[2][1]) == int(ref[2][4] - ref[2][3])
    assert all(rc == 0 for k, rc in enumerate(rcs) if k != 3)


@pytest.mark.gpu
def test_hybrid_index_wire_delivers_the_same_buffers(ctx):
    """ctc_ctx_set_host_wire_share: whichever launch groups ship packed quad records (widened by host threads) and
    whichever ship six u32 per quad, a large host-buffer call into page-locked memory delivers identical buffers."""
    import ctypes as C
    import torch
    import cantucci_b200 as cb
    L = cb.lib()
    shape = cb.Mandelbulb.classic(6, 2.5, fast=True)
    sh = shape._ctc_shape()
    spans = cb.tile_volume(shape.bounding_box(), 8)           # 512 spans at R = 32: ramped launch groups
    ns = len(spans)
    vcap, icap = 600_000, 3_600_000
    v = torch.empty((vcap, 7), dtype=torch.float32).pin_memory()
    i = torch.empty((icap,), dtype=torch.int32).pin_memory()
    v_off = np.zeros(ns + 1, np.uint64); i_off = np.zeros(ns + 1, np.uint64)
    ctx.set_group_spans(48)                                   # 11 launch groups
    ctx.set_host_index_wire(2)                                # packed records for every call size
    got = {}
    try:
        for num, den in ((1, 1), (1, 2), (1, 3), (0, 1)):
            ctx.set_host_wire_share(num, den)
            i.zero_()
            ctx.check(L.ctc_mesh_spans(ctx.handle, C.byref(sh), spans.ctypes.data, ns, 32, v.data_ptr(), vcap, i.data_ptr(), icap,
                                       v_off.ctypes.data, i_off.ctypes.data, None))
            nv, ni = int(v_off[ns]), int(i_off[ns])
            got[(num, den)] = (v.numpy()[:nv].copy(), i.numpy()[:ni].copy(), v_off.copy(), i_off.copy(), ctx.mesh_d2h_bytes())
    finally:
        ctx.set_host_wire_share(1, 1); ctx.set_group_spans(0); ctx.set_host_index_wire(1)
    ref = got[(1, 1)]
    assert ref[1].size > 0
    for key, g in got.items():
        assert np.array_equal(g[0].view(np.uint32), ref[0].view(np.uint32)) and np.array_equal(g[1], ref[1]), key
        assert np.array_equal(g[2], ref[2]) and np.array_equal(g[3], ref[3]), key
    nq = ref[1].size // 6
    assert got[(1, 1)][4] == ref[0].shape[0] * 28 + nq * 8 and got[(0, 1)][4] == ref[0].shape[0] * 28 + nq * 24
    assert got[(1, 1)][4] < got[(1, 2)][4] < got[(0, 1)][4]
