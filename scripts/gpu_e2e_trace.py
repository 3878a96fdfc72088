"""Host-side milestones of the e2e call (CANTUCCI_B200_TRACE) for a few host-thread counts."""
import os, subprocess, sys
here = os.path.dirname(os.path.abspath(__file__))
code = r'''
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(sys.argv[1]))
import numpy as np, torch
import cantucci_b200 as cb
from cantucci_b200 import _lib
L = _lib.lib(); ctx = cb.Context(0)
spans = cb.tile_volume(cb.Span((-1.2,) * 3, (1.2,) * 3), 16)
sh = cb.Mandelbulb.classic(6, 2.5, fast=True)._ctc_shape()
ns = len(spans); vcap, icap = 14_000_000, 84_000_000
v_off = np.zeros(ns + 1, np.uint64); i_off = np.zeros(ns + 1, np.uint64)
v = torch.empty((vcap, 7), dtype=torch.float32).pin_memory(); i = torch.empty((icap,), dtype=torch.int32).pin_memory()
for k in range(6):
    t0 = time.perf_counter()
    ctx.check(L.ctc_mesh_spans(ctx.handle, C.byref(sh), spans.ctypes.data, ns, 64, v.data_ptr(), vcap, i.data_ptr(), icap, v_off.ctypes.data, i_off.ctypes.data, None))
    print(f"call {k}: {(time.perf_counter() - t0) * 1e3:.2f} ms", file=sys.stderr)
'''
for threads in ("4", "8", "16"):
    env = dict(os.environ, CANTUCCI_B200_TRACE="1", CANTUCCI_B200_EXPAND_THREADS=threads)
    print(f"== {threads} host threads", flush=True)
    r = subprocess.run([sys.executable, "-c", code, here], env=env, capture_output=True, text=True)
    print("\n".join(r.stderr.strip().splitlines()[-6:]), flush=True)
