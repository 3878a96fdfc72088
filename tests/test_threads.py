"""Shape: Sync + Send -- the reference calls generate_for_box from num_cpus worker threads at once
(mesh/mod.rs:61-62,141).  One shared context (calls serialise on its mutex) and one context per
thread must both give the single-threaded results."""
import threading

import numpy as np
import pytest

from conftest import startup_leaves

pytestmark = pytest.mark.gpu


def test_concurrent_callers_get_identical_meshes(ctx):
    import cantucci_b200 as cb
    spans = startup_leaves()[:16]
    shape = cb.Mandelbulb.classic(6, 2.5)
    want = [cb.MeshBuffer.generate_for_box(cb.Span(tuple(r[:3]), tuple(r[3:])), shape, 32, ctx)[0] for r in spans]
    for shared in (True, False):
        out, errs = [None] * len(spans), []

        def work(k):
            try:
                c = ctx if shared else cb.Context(0)
                out[k] = cb.MeshBuffer.generate_for_box(cb.Span(tuple(spans[k][:3]), tuple(spans[k][3:])), shape, 32, c)[0]
            except Exception as e:       # noqa: BLE001
                errs.append(e)

        th = [threading.Thread(target=work, args=(k,)) for k in range(len(spans))]
        [t.start() for t in th]
        [t.join() for t in th]
        assert not errs, errs
        for a, b in zip(out, want):
            assert np.array_equal(a.indices, b.indices)
            assert np.array_equal(a.vertices.view(np.uint32), b.vertices.view(np.uint32))
