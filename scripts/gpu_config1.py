"""Device-resident latency of small batches (config 1 = 64 startup leaves; 16 / 32 / 128 / 176 spans too)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cantucci_b200 as cb
from cantucci_b200 import refine
from cantucci_b200.scheduler import DeviceMesher
ctx = cb.Context(0); dev = torch.device("cuda", 0)
shape = cb.Mandelbulb.classic(6, 2.5, fast=True); sh = shape._ctc_shape()
startup = cb.spans_array([n.span for n in cb.startup_tree(shape.bounding_box()).leaves()])
leaves, _ = refine.config3_spans(cb.Mandelbulb.classic(6, 2.5), 6, ctx)
tiles = cb.tile_volume(cb.Span((-1.2,) * 3, (1.2,) * 3), 16)
out = []
for name, sp in (("16", startup[:16]), ("32", startup[:32]), ("64 (config 1)", startup), ("128", np.ascontiguousarray(tiles[1000:1128])), ("176 (config 3)", leaves)):
    sp = np.ascontiguousarray(sp)
    m = DeviceMesher(ctx, torch, dev, 3_000_000, 18_000_000, len(sp))
    for _ in range(5): m.launch(sh, sp, 64); m.result()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(50): m.launch(sh, sp, 64); m.result()
    b.record(); torch.cuda.synchronize()
    out.append(f"{name}: {a.elapsed_time(b) / 50 * 1e3:.0f} us")
print(os.path.basename(os.environ.get("CANTUCCI_B200_LIB", "default")), " | ".join(out), flush=True)
