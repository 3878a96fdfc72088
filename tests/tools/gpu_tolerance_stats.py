"""Relative error of the tolerance modes against the oracle, bucketed by completed iterations."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import cantucci_b200 as cb
from oracle import oracle as O

ctx = cb.default_context(0)
rng = np.random.default_rng(13)
pts = rng.uniform(-1.3, 1.3, size=(60000, 3)).astype(np.float32)
for power, iters, fast in [(8, 6, True), (8, 32, True), (2, 32, False), (2, 32, True), (4, 32, False), (4, 32, True),
                           (16, 32, False), (16, 32, True)]:
    sh = O.mandelbulb(power, iters, 2.5)
    infos = [O.min_distance_from_info(sh, p) for p in pts]
    want = np.array([i[0] for i in infos], dtype=np.float32)
    k = np.array([i[1].iters for i in infos]); bailed = np.array([i[1].bailed for i in infos]).astype(bool)
    margin = np.array([i[1].min_margin for i in infos])
    got = cb.Mandelbulb(power, iters, 2.5, fast=fast).batch_min_distance_from(pts, ctx)
    rel = np.abs(got.astype(np.float64) - want) / np.maximum(np.abs(want), 1e-30)
    row = {"power": power, "iters": iters, "fast": fast, "sign_mismatch": float(np.mean((got.view(np.uint32) >> 31) != (want.view(np.uint32) >> 31)))}
    for kk in range(1, 9):
        sel = bailed & (k == kk) & (margin > 1e-4)
        if sel.sum() > 20:
            row[f"k{kk}"] = [int(sel.sum())] + [float(f"{q:.2e}") for q in np.quantile(rel[sel], [0.5, 0.99, 1.0])]
    sel = ~bailed & np.isfinite(want)
    row["interior"] = [int(sel.sum())] + [float(f"{q:.2e}") for q in np.quantile(rel[sel], [0.5, 0.99])]
    print(json.dumps(row), flush=True)
