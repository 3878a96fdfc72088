"""Debug aid: exact-mode parity of one config-3 leaf against the oracle, field by field."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cantucci_b200 as cb
from cantucci_b200 import refine
from oracle import oracle as O

ctx = cb.Context(0)
bulb = cb.Mandelbulb.classic(6, 2.5)
spans, _ = refine.config3_spans(bulb, 6, ctx)
sh = O.mandelbulb(8, 6, 2.5)
for k in (0, 1, 2, 5):
    row = spans[k:k + 1]
    g = cb.sample_grids(row, bulb, 64, ctx)[0]
    want = O.sample_grid(sh, O.make_span(row[0, :3], row[0, 3:]), 64)
    bad = np.nonzero(g.view(np.uint32) != want.view(np.uint32))[0]
    print("span", k, row, "grid mismatches", bad.size, bad[:8], g[bad[:4]], want[bad[:4]])
    batch, _ = cb.generate_for_boxes(row, bulb, 64, ctx)
    v, i, _ = O.generate_for_box(sh, O.make_span(row[0, :3], row[0, 3:]), 64)
    got = batch.mesh(0)
    print("   verts", len(got.vertices), len(v), "idx equal", np.array_equal(got.indices, i))
    if len(v) == len(got.vertices) and len(v):
        a = got.vertices.view(np.uint32).reshape(-1, 7); b = v.view(np.uint32).reshape(-1, 7)
        print("   per-field mismatching rows:", [(int((a[:, c] != b[:, c]).sum())) for c in range(7)])
        rows = np.nonzero((a != b).any(axis=1))[0][:3]
        for r in rows:
            print("   row", r, got.vertices[r], v[r])
