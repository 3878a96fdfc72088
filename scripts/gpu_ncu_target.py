"""Short workload for ncu launch lists: config 1 (exact, fast) and the 512^3 volume as spans."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cantucci_b200 as cb
ctx = cb.default_context(0)
bbox = cb.Span((-1.2,)*3, (1.2,)*3)
startup = cb.spans_array([n.span for n in cb.startup_tree(bbox).leaves()])
which = sys.argv[1] if len(sys.argv) > 1 else "all"
for fast in (False, True):
    if which in ("all", "config1"):
        cb.generate_for_boxes(startup, cb.Mandelbulb.classic(6, 2.5, fast=fast), 64, ctx)
    if which in ("all", "tiles8"):
        cb.generate_for_boxes(cb.tile_volume(bbox, 8), cb.Mandelbulb.classic(6, 2.5, fast=fast), 64, ctx)
