"""ctc_order_spans: the cost-aware span order of the scheduler.  The order is a permutation, puts every
span that can hold surface before every provably empty one, and never changes a mesh."""
import numpy as np
import pytest

import cantucci_b200 as cb
from conftest import startup_leaves


@pytest.mark.gpu
def test_order_is_a_stable_permutation_with_the_empty_spans_last(ctx, oracle):
    shape = cb.Mandelbulb.classic(6, 2.5)
    spans = cb.tile_volume(shape.bounding_box(), 8)           # 512 spans
    order = cb.order_spans(spans, shape, 32, ctx)
    assert np.array_equal(np.sort(order), np.arange(512))
    # the key, recomputed with the oracle: |DE(centre)| / reach, ascending
    sp = spans.view(np.float32).reshape(-1, 6)
    centres = np.ascontiguousarray(((sp[:, :3].astype(np.float64) + sp[:, 3:]) * 0.5).astype(np.float32))
    d = oracle.batch_min_distance_from(oracle.mandelbulb(8, 6, 2.5), centres)
    ext = sp[:, 3:].astype(np.float64) - sp[:, :3]
    reach = (np.sqrt(((0.5 * ext + ext / 32) ** 2).sum(axis=1)) * 1.0001).astype(np.float32)
    key = np.abs(d) / reach
    key[np.isnan(key)] = 0.0
    assert np.array_equal(order, np.argsort(key, kind="stable"))
    # every span the DE bound (with the culling margin of 2: the Mandelbulb estimate is not a rigorous bound, and at
    # this span size a few spans with DE(centre) slightly above their reach do hold surface) proves empty comes after
    # every span that holds vertices
    batch, _ = cb.generate_for_boxes(spans, shape, 32, ctx)
    nv = np.diff(batch.v_off.astype(np.int64))
    place = np.empty(512, np.int64); place[order] = np.arange(512)
    proven_empty = d > 2.0 * reach
    assert proven_empty.any() and not nv[proven_empty].any()
    assert place[nv > 0].max() < place[proven_empty].min()


@pytest.mark.gpu
@pytest.mark.parametrize("fast", [False, True])
def test_surface_first_meshes_are_the_callers_meshes(ctx, fast):
    spans = startup_leaves()
    shape = cb.Mandelbulb.classic(6, 2.5, fast=fast)
    ref, tr = cb.generate_for_boxes(spans, shape, 32, ctx)
    got, tg = cb.generate_for_boxes(spans, shape, 32, ctx, surface_first=True)
    assert got.slot is not None and tg.vertices == tr.vertices and tg.faces == tr.faces
    # the buffers are front-loaded: the first half of the table entries holds more vertices than the second
    counts = np.diff(got.v_off.astype(np.int64))
    assert counts[:32].sum() > counts[32:].sum()
    for k in range(len(spans)):
        a, b = got.mesh(k), ref.mesh(k)
        assert np.array_equal(a.indices, b.indices), k
        assert np.array_equal(a.vertices.view(np.uint32), b.vertices.view(np.uint32)), k


@pytest.mark.gpu
def test_order_spans_rejects_what_the_mesh_call_rejects(ctx):
    shape = cb.Mandelbulb.classic(6, 2.5)
    bad = startup_leaves()[:2].copy()
    bad[1, 3] = bad[1, 0]                                     # start == end on x
    with pytest.raises(AssertionError):
        cb.order_spans(bad, shape, 32, ctx)
    assert cb.order_spans(startup_leaves()[:0], shape, 32, ctx).shape == (0,)


@pytest.mark.gpu
def test_culling_and_surface_first_compose(ctx):
    shape = cb.Mandelbulb.classic(6, 2.5, fast=True)
    spans = cb.tile_volume(shape.bounding_box(), 8)[::3]          # 171 spans, many of them empty space
    ref, tr = cb.generate_for_boxes(spans, shape, 32, ctx)
    got, tg = cb.generate_for_boxes(spans, shape, 32, ctx, cull=True, surface_first=True)
    assert tg.vertices == tr.vertices and tg.faces == tr.faces
    for k in range(len(spans)):
        a, b = got.mesh(k), ref.mesh(k)
        assert np.array_equal(a.indices, b.indices), k
        assert np.array_equal(a.vertices.view(np.uint32), b.vertices.view(np.uint32)), k
