"""ctc_order_spans on the benched volume: what the call costs, and what meshing in its order changes -- the
device-resident step, the e2e call (pinned host buffers), and per-span equality of the meshes."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cantucci_b200 as cb
from cantucci_b200 import _lib
from cantucci_b200.scheduler import DeviceMesher
L = _lib.lib(); ctx = cb.Context(0); dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
spans = cb.tile_volume(cb.Span((-1.2,) * 3, (1.2,) * 3), 16)
bulb = cb.Mandelbulb.classic(6, 2.5, fast=True); sh = bulb._ctc_shape()
ns = len(spans)
order = cb.order_spans(spans, bulb, 64, ctx)
o32 = np.zeros(ns, np.uint32)
t0 = time.perf_counter()
for _ in range(50):
    ctx.check(L.ctc_order_spans(ctx.handle, C.byref(sh), spans.ctypes.data, ns, 64, o32.ctypes.data))
print(f"ctc_order_spans, {ns} spans: {(time.perf_counter() - t0) / 50 * 1e6:.0f} us per call", flush=True)
sorted_spans = np.ascontiguousarray(spans[order])
m = DeviceMesher(ctx, torch, dev, 14_000_000, 84_000_000, ns)

def dev_step(sp, reps=20):
    for _ in range(3):
        m.launch(sh, sp, 64); m.result()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        m.launch(sh, sp, 64); m.result()
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for rep in range(2):
    print(f"device-resident step: caller order {dev_step(spans):.3f} ms, surface-first {dev_step(sorted_spans):.3f} ms", flush=True)

vcap, icap = 14_000_000, 84_000_000
v_off = np.zeros(ns + 1, np.uint64); i_off = np.zeros(ns + 1, np.uint64)
v = torch.empty((vcap, 7), dtype=torch.float32).pin_memory(); i = torch.empty((icap,), dtype=torch.int32).pin_memory()

def e2e(sp, with_order, reps=10):
    def call():
        if with_order:
            ctx.check(L.ctc_order_spans(ctx.handle, C.byref(sh), spans.ctypes.data, ns, 64, o32.ctypes.data))
            s2 = np.ascontiguousarray(spans[o32])
        else:
            s2 = sp
        ctx.check(L.ctc_mesh_spans(ctx.handle, C.byref(sh), s2.ctypes.data, ns, 64, v.data_ptr(), vcap, i.data_ptr(), icap, v_off.ctypes.data, i_off.ctypes.data, None))
    for _ in range(3): call()
    t0 = time.perf_counter()
    for _ in range(reps): call()
    return (time.perf_counter() - t0) / reps * 1e3

for rep in range(2):
    print(f"e2e: caller order {e2e(spans, False):.2f} ms, surface-first (order call inside) {e2e(None, True):.2f} ms, "
          f"surface-first (pre-ordered) {e2e(sorted_spans, False):.2f} ms", flush=True)
# per-span equality
a, _ = cb.generate_for_boxes(spans[:512], bulb, 64, ctx)
b, _ = cb.generate_for_boxes(spans[:512], bulb, 64, ctx, surface_first=True)
same = all(np.array_equal(a.mesh(k).indices, b.mesh(k).indices) and
           np.array_equal(a.mesh(k).vertices.view(np.uint32), b.mesh(k).vertices.view(np.uint32)) for k in range(512))
print("surface_first meshes equal the caller-order meshes span by span:", same)
