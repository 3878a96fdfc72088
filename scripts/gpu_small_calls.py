"""Latency of small ctc_mesh_spans calls (1 / 8 / 64 startup leaves), pageable and pinned destinations."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cantucci_b200 as cb
from cantucci_b200 import _lib
L = _lib.lib(); ctx = cb.Context(0)
shape = cb.Mandelbulb.classic(6, 2.5, fast=True); sh = shape._ctc_shape()
startup = cb.spans_array([n.span for n in cb.startup_tree(shape.bounding_box()).leaves()])
for pinned in (False, True):
    if pinned:
        vt = torch.empty((700_000, 7), dtype=torch.float32).pin_memory(); it = torch.empty((4_200_000,), dtype=torch.int32).pin_memory()
        pv, pi = vt.data_ptr(), it.data_ptr()
    else:
        v = np.empty(700_000, dtype=cb.VERTEX_DTYPE); idx = np.empty(4_200_000, dtype=np.uint32); pv, pi = v.ctypes.data, idx.ctypes.data
    for n in (1, 8, 64):
        sp = np.ascontiguousarray(startup[20:20 + n] if n < 64 else startup)
        v_off = np.zeros(n + 1, dtype=np.uint64); i_off = np.zeros(n + 1, dtype=np.uint64); tt = _lib.CtcTimings()
        def call():
            ctx.check(L.ctc_mesh_spans(ctx.handle, C.byref(sh), sp.ctypes.data, n, 64, pv, 700_000, pi, 4_200_000, v_off.ctypes.data, i_off.ctypes.data, C.byref(tt)))
        for _ in range(5): call()
        t0 = time.perf_counter()
        for _ in range(100): call()
        wall = (time.perf_counter() - t0) / 100
        print(f"pinned={pinned} spans={n:2d}: wall {wall*1e6:7.1f} us, device kernels {(tt.first_ms+tt.second_ms+tt.third_ms)*1e3:6.1f} us, {int(v_off[n])} vertices", flush=True)
