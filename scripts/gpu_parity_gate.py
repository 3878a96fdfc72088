"""Runs the parity gate of bench.py on the benched volume (or a subset) in fast and exact mode."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import cantucci_b200 as cb

tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 16
stride = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = cb.Context(0)
spans = np.ascontiguousarray(bench.workload_spans(tiles)[::stride])
t0 = time.time()
ora = bench.oracle_volume(spans)
print("oracle", ora["spans"], "spans", round(ora["secs"], 2), "s pool,", round(time.time() - t0, 2), "s total,", ora["threads"], "threads", flush=True)
for fast in (True, False):
    shape = cb.Mandelbulb.classic(bench.MAX_ITERS, bench.BAILOUT, fast=fast)
    t0 = time.time()
    gv, gi, gvo, gio, gpl = bench.gpu_volume_host(ctx, shape, spans, vcap=int(ora["v_off"][-1] * 1.05) + 1024, icap=int(ora["i_off"][-1] * 1.05) + 6144)
    t1 = time.time()
    par = bench.parity_gate(gv, gi, gvo, gio, gpl, ora, spans, exact=not fast)
    par["mode"] = "fast" if fast else "exact"
    par["fixups"] = ctx.mesh_fixups()
    par["gpu_secs"] = round(t1 - t0, 2); par["gate_secs"] = round(time.time() - t1, 2)
    print(json.dumps(par), flush=True)
