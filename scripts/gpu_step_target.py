"""Short workload for ncu: `reps` device-resident steps of the benched 1024^3 volume (fast mode unless `exact`)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import cantucci_b200 as cb
from cantucci_b200 import _lib
from cantucci_b200.scheduler import DeviceMesher

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
exact = len(sys.argv) > 2 and sys.argv[2] == "exact"
overlap = not (len(sys.argv) > 3 and sys.argv[3] == "serial")
power = int(sys.argv[4]) if len(sys.argv) > 4 else 8
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 6
ctx = cb.Context(0)
ctx.set_overlap(overlap)
dev = torch.device("cuda", 0)
spans = cb.tile_volume(cb.Span((-1.2,) * 3, (1.2,) * 3), 16)
sh = cb.Mandelbulb(power, iters, 2.5, fast=not exact)._ctc_shape()
m = DeviceMesher(ctx, torch, dev, 40_000_000, 240_000_000, len(spans))
for _ in range(reps):
    m.launch(sh, spans, 64)
    print(m.result(allow_lerp_assert=True)[:2], ctx.mesh_fixups())
