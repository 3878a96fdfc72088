"""How much of the e2e call (ctc_mesh_spans, pinned host buffers) is the ORDER of the spans?  The device->host copy
pipeline idles while the first launch groups hold only empty spans (the corner of the bounding box) and has a tail
when the last groups hold surface.  Compares the caller's lexicographic order with surface-first orders, beside the
floor: the same bytes as plain device->host copies."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cantucci_b200 as cb
from cantucci_b200 import _lib
L = _lib.lib(); ctx = cb.Context(0)
T = int(sys.argv[1]) if len(sys.argv) > 1 else 16
spans = cb.tile_volume(cb.Span((-1.2,) * 3, (1.2,) * 3), T)
bulb = cb.Mandelbulb.classic(6, 2.5, fast=True)
sh = bulb._ctc_shape()
ns = len(spans); vcap, icap = 14_000_000 * (T // 16) ** 2, 84_000_000 * (T // 16) ** 2
v_off = np.zeros(ns + 1, np.uint64); i_off = np.zeros(ns + 1, np.uint64)
v = torch.empty((vcap, 7), dtype=torch.float32).pin_memory(); i = torch.empty((icap,), dtype=torch.int32).pin_memory()
pv, pi = v.data_ptr(), i.data_ptr()

def run(sp, label, reps=10):
    sp = np.ascontiguousarray(sp)
    def call():
        ctx.check(L.ctc_mesh_spans(ctx.handle, C.byref(sh), sp.ctypes.data, ns, 64, pv, vcap, pi, icap, v_off.ctypes.data, i_off.ctypes.data, None))
    for _ in range(3): call()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); call(); ts.append(time.perf_counter() - t0)
    ts = np.array(ts) * 1e3
    print(f"{label:44s}: median {np.median(ts):6.2f} ms  min {ts.min():6.2f}  ({int(v_off[ns])} v, {int(i_off[ns])} i)", flush=True)
    return np.diff(v_off.astype(np.int64)).copy()

counts = run(spans, "caller order (lexicographic tiles)")
nv, ni = int(v_off[ns]), int(i_off[ns])
# floor: the same bytes, two streams
dv = torch.empty((nv, 7), dtype=torch.float32, device="cuda"); dq = torch.empty((ni // 6, 2), dtype=torch.int32, device="cuda")
hq = torch.empty((ni // 6, 2), dtype=torch.int32).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with torch.cuda.stream(s1): v[:nv].copy_(dv, non_blocking=True)
    with torch.cuda.stream(s2): hq.copy_(dq, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"floor: plain D2H of {(nv * 28 + ni // 6 * 8) / 1e6:.0f} MB on two streams: {dt * 1e3:.2f} ms = {(nv * 28 + ni // 6 * 8) / dt / 1e9:.1f} GB/s", flush=True)

order = np.argsort(-counts, kind="stable")
run(spans[order], "ideal: most vertices first")
run(spans[order[::-1]], "worst: fewest vertices first")
# practical: DE at the span centre against the span's reach (what ctc_cull_spans computes)
sp = spans.view(np.float32).reshape(ns, 6)
centres = np.ascontiguousarray(0.5 * (sp[:, :3] + sp[:, 3:]))
t0 = time.perf_counter()
d = cb.Mandelbulb.classic(6, 2.5).batch_min_distance_from(centres, ctx)
key = np.where(np.isnan(d), 0.0, np.abs(d))
order2 = np.argsort(key, kind="stable")
t1 = time.perf_counter()
print(f"ordering by |DE(centre)|: {1e3 * (t1 - t0):.2f} ms host+device", flush=True)
run(spans[order2], "practical: smallest |DE(centre)| first")
# two-class: spans that may hold surface first (DE(centre) <= reach), caller order inside each class
reach = 0.5 * np.sqrt(3.0) * (sp[:, 3] - sp[:, 0]) * (1 + 2 / 64)
maybe = ~(d > reach)
order3 = np.concatenate([np.nonzero(maybe)[0], np.nonzero(~maybe)[0]])
run(spans[order3], f"two classes: {int(maybe.sum())} possible-surface spans first")
ctx.set_host_index_wire(False)
run(spans, "caller order, u32 indices over PCIe")
run(spans[order], "ideal order, u32 indices over PCIe")
