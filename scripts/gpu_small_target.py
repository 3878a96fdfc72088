"""Short workload for ncu launch lists: 1-, 8- and 64-span host-buffer calls (fast mode)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cantucci_b200 as cb
from cantucci_b200 import _lib
ctx = cb.Context(0)
L = _lib.lib()
bbox = cb.Span((-1.2,) * 3, (1.2,) * 3)
startup = cb.spans_array([n.span for n in cb.startup_tree(bbox).leaves()])
sh = cb.Mandelbulb.classic(6, 2.5, fast=True)._ctc_shape()
v = np.empty(700_000, dtype=cb.VERTEX_DTYPE); idx = np.empty(4_200_000, dtype=np.uint32)
for n in (1, 8, 64):
    sp = np.ascontiguousarray(startup[20:20 + n] if n < 64 else startup)
    v_off = np.zeros(n + 1, dtype=np.uint64); i_off = np.zeros(n + 1, dtype=np.uint64)
    for _ in range(3):
        ctx.check(L.ctc_mesh_spans(ctx.handle, C.byref(sh), sp.ctypes.data, n, 64, v.ctypes.data, len(v), idx.ctypes.data, len(idx),
                                   v_off.ctypes.data, i_off.ctypes.data, None))
    print(n, int(v_off[n]))
