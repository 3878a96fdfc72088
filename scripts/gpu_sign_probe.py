"""Calibration of fast mode's sign-trust band (run on the B200): ctc_fast_sign_probe over the benched
1024^3 volume, config 1 and a max_iters sweep.  Writes profiles/sign_probe_r2.md-ready JSON lines."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cantucci_b200 as cb
from cantucci_b200 import _lib

ctx = cb.Context(0)
L = _lib.lib()
bbox = cb.Span((-1.2, -1.2, -1.2), (1.2, 1.2, 1.2))


def probe(name, spans, iters, bail, R=64):
    sh = cb.Mandelbulb.classic(iters, bail, fast=True)._ctc_shape()
    out = (C.c_uint64 * 489)()
    t0 = time.time()
    ctx.check(L.ctc_fast_sign_probe(ctx.handle, C.byref(sh), spans.ctypes.data, spans.shape[0], R, out, 489))
    o = [int(x) for x in out]
    raw = np.frombuffer(out, dtype=np.uint64)
    ndump = min(int(raw[104] & 0xFFFFFFFF), 64)
    dump = raw[105:].view(np.float32).reshape(64, 12)[:ndump]
    f = lambda b: float(np.uint32(b).view(np.float32))
    rec = {"name": name, "spans": int(spans.shape[0]), "max_iters": iters, "bailout": bail, "samples": o[0],
           "raw_sign_mismatches": o[1], "mismatches_in_axis_band": o[2], "both_inside": o[3], "escape_status_differs": o[4],
           "max_err_over_dr": f(o[5]), "max_err_over_amp": f(o[6]), "mismatches_fast_escaped": o[7],
           "uncovered_at_2^-17_amp": [[float(x) for x in r] for r in dump],
           "secs": round(time.time() - t0, 2), "kappa": []}
    for q in range(24):
        s_max, u_max, s_pol, u_pol = o[8 + 4 * q: 12 + 4 * q]
        rec["kappa"].append({"log2": -(8 + q), "suspects_dr": s_max, "uncovered_dr": u_max,
                             "suspects_amp": s_pol, "uncovered_amp": u_pol})
    print(json.dumps(rec), flush=True)
    return rec


vol = cb.tile_volume(bbox, 16)
probe("bbox_1024cube_16x16x16_spans_R64", vol, 6, 2.5)
tree = cb.startup_tree(bbox)
start = cb.spans_array([n.span for n in tree.leaves()])
probe("config1_startup_64_spans", start, 6, 2.5)
probe("config1_i8_b5", start, 8, 5.0)
sub = np.ascontiguousarray(vol[::8])
for it in (2, 3, 4, 12, 32, 128):
    probe(f"bbox_1024cube_every_8th_span_i{it}", sub, it, 2.5)
