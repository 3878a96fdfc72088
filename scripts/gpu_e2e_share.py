"""e2e call on the benched volume for several shares of packed launch groups (ctc_ctx_set_host_wire_share)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cantucci_b200 as cb
from cantucci_b200 import _lib
L = _lib.lib(); ctx = cb.Context(0)
spans = cb.tile_volume(cb.Span((-1.2,) * 3, (1.2,) * 3), 16)
bulb = cb.Mandelbulb.classic(6, 2.5, fast=True); sh = bulb._ctc_shape()
ns = len(spans); vcap, icap = 14_000_000, 84_000_000
spans = np.ascontiguousarray(spans[cb.order_spans(spans, bulb, 64, ctx)])
v_off = np.zeros(ns + 1, np.uint64); i_off = np.zeros(ns + 1, np.uint64)
v = torch.empty((vcap, 7), dtype=torch.float32).pin_memory(); i = torch.empty((icap,), dtype=torch.int32).pin_memory()
ref = None
for num, den in ((4, 4), (3, 4), (2, 3), (2, 4), (1, 3), (1, 4), (0, 4), (2, 4), (4, 4)):
    ctx.set_host_wire_share(num, den)
    def call():
        ctx.check(L.ctc_mesh_spans(ctx.handle, C.byref(sh), spans.ctypes.data, ns, 64, v.data_ptr(), vcap, i.data_ptr(), icap, v_off.ctypes.data, i_off.ctypes.data, None))
    for _ in range(3): call()
    ts = []
    for _ in range(12):
        t0 = time.perf_counter(); call(); ts.append(time.perf_counter() - t0)
    ts = np.array(ts) * 1e3
    ni = int(i_off[ns])
    digest = hash(i.numpy()[:ni].tobytes())
    if ref is None: ref = digest
    print(f"packed share {num}/{den}: median {np.median(ts):6.2f} ms  min {ts.min():6.2f}  d2h {ctx.mesh_d2h_bytes() / 1e6:.0f} MB  indices equal: {digest == ref}", flush=True)
