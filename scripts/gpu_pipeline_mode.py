"""Device-resident step vs the pipelined (copy-out) step with a destination on the SAME GPU: what the launch-group
ramp and the copy pipeline cost when no link is in the way."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cantucci_b200 as cb
from cantucci_b200 import _lib
from cantucci_b200.scheduler import DeviceMesher
L = _lib.lib(); ctx = cb.Context(0); dev = torch.device("cuda", 0)
spans = cb.tile_volume(cb.Span((-1.2,) * 3, (1.2,) * 3), 16)
sh = cb.Mandelbulb.classic(6, 2.5, fast=True)._ctc_shape()
ns = len(spans); vcap, icap = 14_000_000, 84_000_000
m = DeviceMesher(ctx, torch, dev, vcap, icap, ns)
def dev_step():
    m.launch(sh, spans, 64); m.result()
v = torch.empty((vcap, 7), dtype=torch.float32, device=dev); i = torch.empty((icap,), dtype=torch.int32, device=dev)
vo = torch.zeros(ns + 1, dtype=torch.int64, device=dev); io = torch.zeros(ns + 1, dtype=torch.int64, device=dev)
def pipe_step():
    ctx.check(L.ctc_mesh_spans(ctx.handle, C.byref(sh), spans.ctypes.data, ns, 64, v.data_ptr(), vcap, i.data_ptr(), icap, vo.data_ptr(), io.data_ptr(), None))
for name, fn in (("device-resident", dev_step), ("pipelined, local device destination", pipe_step)):
    for gs in (0, 466, 256, 128):
        ctx.set_group_spans(gs)
        for _ in range(3): fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(20): fn()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
        print(f"{name:40s} group_spans={gs:4d}: {dt*1e3:.3f} ms", flush=True)
for wire in (0, 1):
    L.ctc_ctx_set_index_wire(ctx.handle, wire); ctx.set_group_spans(0)
    for _ in range(3): pipe_step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): pipe_step()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
    print(f"pipelined, packed wire={wire}: {dt*1e3:.3f} ms", flush=True)
