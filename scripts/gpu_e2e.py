"""e2e latency of ctc_mesh_spans on the benched volume: host index wire on/off, pinned/pageable destination."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cantucci_b200 as cb
from cantucci_b200 import _lib
L = _lib.lib(); ctx = cb.Context(0)
spans = cb.tile_volume(cb.Span((-1.2,) * 3, (1.2,) * 3), int(sys.argv[1]) if len(sys.argv) > 1 else 16)
sh = cb.Mandelbulb.classic(6, 2.5, fast=True)._ctc_shape()
ns = len(spans); vcap, icap = 14_000_000, 84_000_000
v_off = np.zeros(ns + 1, np.uint64); i_off = np.zeros(ns + 1, np.uint64)
for pinned in (True, False):
    if pinned:
        v = torch.empty((vcap, 7), dtype=torch.float32).pin_memory(); i = torch.empty((icap,), dtype=torch.int32).pin_memory()
        pv, pi = v.data_ptr(), i.data_ptr()
    else:
        v = np.empty((vcap, 7), np.float32); i = np.empty(icap, np.uint32); pv, pi = v.ctypes.data, i.ctypes.data
    for wire in (True, False):
        ctx.set_host_index_wire(wire)
        def call():
            ctx.check(L.ctc_mesh_spans(ctx.handle, C.byref(sh), spans.ctypes.data, ns, 64, pv, vcap, pi, icap, v_off.ctypes.data, i_off.ctypes.data, None))
        for _ in range(3): call()
        t0 = time.perf_counter()
        for _ in range(10): call()
        dt = (time.perf_counter() - t0) / 10
        print(f"pinned={pinned} host_wire={wire} threads={ctx.host_index_wire_stats()[2]}: {dt*1e3:.2f} ms  ({int(v_off[ns])} v, {int(i_off[ns])} i)", flush=True)
