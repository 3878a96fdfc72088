"""Generates the golden fixtures in this directory FROM THE ORACLE (the reference ships no
expected outputs and cannot be built here -- parity unpinned, see oracle/cantucci_oracle.h).
Run:  python tests/golden/make_golden.py"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from conftest import BENCH_POINTS, startup_leaves  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    gold = {}
    for power, iters, bail in [(8, 8, 5.0), (8, 6, 2.5), (8, 32, 2.5), (2, 32, 2.5), (4, 32, 2.5), (16, 32, 2.5)]:
        sh = O.mandelbulb(power, iters, bail)
        got = O.batch_min_distance_from(sh, BENCH_POINTS)
        gold[f"p{power}_i{iters}_b{bail}"] = [f"{b:08x}" for b in got.view(np.uint32)]
    json.dump(gold, open(os.path.join(HERE, "bench_points_de.json"), "w"), indent=1)

    spans = startup_leaves()
    sh = O.mandelbulb(8, 6, 2.5)
    meshes, _ = O.generate_for_boxes_mt(sh, spans, 64)
    hv, hi = hashlib.sha256(), hashlib.sha256()
    for v, i, _ in meshes:
        hv.update(v.tobytes()); hi.update(i.tobytes())
    json.dump({
        "spans_sha256": hashlib.sha256(spans.tobytes()).hexdigest(),
        "vertices_per_span": [len(m[0]) for m in meshes],
        "quads_per_span": [len(m[1]) // 6 for m in meshes],
        "vertices_sha256": hv.hexdigest(),
        "indices_sha256": hi.hexdigest(),
    }, open(os.path.join(HERE, "config1_startup.json"), "w"), indent=1)

    v, i, _ = O.generate_for_box(sh, O.make_span((0.0, 0.0, 0.0), (0.6, 0.6, 0.6)), 16)
    sp = O.sphere((0.0, 0.0, 0.0), 0.9)
    sv, si, _ = O.generate_for_box(sp, O.make_span((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0)), 16)
    np.savez_compressed(os.path.join(HERE, "small_meshes.npz"),
                        bulb_v=v.view(np.uint32).reshape(-1, 7), bulb_i=i,
                        sphere_v=sv.view(np.uint32).reshape(-1, 7), sphere_i=si)
    # config 3: the leaf set of the refinement mirror (host logic), driven by the oracle's DE
    sys.path.insert(0, os.path.dirname(HERE))
    from test_configs import _OracleBulb
    from cantucci_b200 import refine
    spans3, cam = refine.config3_spans(_OracleBulb.classic(6, 2.5), 6)
    json.dump({"n_leaves": int(len(spans3)), "spans_sha256": hashlib.sha256(spans3.tobytes()).hexdigest(),
               "camera_position": [float(c) for c in cam.position]},
              open(os.path.join(HERE, "config3_leaves.json"), "w"), indent=1)
    print("golden fixtures written")


if __name__ == "__main__":
    main()
