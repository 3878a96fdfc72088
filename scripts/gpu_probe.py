"""Quick on-GPU probe: pass timings of the main configs (device time via ctc_timings)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cantucci_b200 as cb

ctx = cb.default_context(0)
bbox = cb.Span((-1.2,)*3, (1.2,)*3)
tree = cb.startup_tree(bbox)
startup = cb.spans_array([n.span for n in tree.leaves()])

def run(name, spans, R, shape, reps=3):
    out = None
    for r in range(reps):
        t0 = time.time()
        batch, t = cb.generate_for_boxes(spans, shape, R, ctx)
        wall = time.time() - t0
        out = (t, wall)
    t, wall = out
    ns = len(cb.spans_array(spans))
    samples = ns * (R + 1) ** 3
    print(json.dumps({"cfg": name, "fast": shape.math_fast, "spans": ns, "R": R, "first_ms": round(t.first, 4),
                      "second_ms": round(t.second, 4), "third_ms": round(t.third, 4), "verts": t.vertices,
                      "faces": t.faces, "Gsamples_per_s_pass1": round(samples / t.first / 1e6, 2),
                      "wall_ms": round(wall * 1e3, 2)}), flush=True)

for fast in (False, True):
    run("config1_startup", startup, 64, cb.Mandelbulb.classic(6, 2.5, fast=fast))
    run("vol512_as_8^3_spans", cb.tile_volume(bbox, 8), 64, cb.Mandelbulb.classic(6, 2.5, fast=fast))
    run("vol1024_as_16^3_spans", cb.tile_volume(bbox, 16), 64, cb.Mandelbulb.classic(6, 2.5, fast=fast), reps=2)
for p in (2, 4, 16):
    for fast in (False, True):
        run(f"vol256_p{p}_i32", cb.tile_volume(bbox, 4), 64, cb.Mandelbulb(p, 32, 2.5, fast=fast), reps=2)
print("launches", ctx.kernel_launches())
