"""BASELINE.json configs 3, 4, 5 as parity / property tests (configs 1 and 2 live in test_gpu_parity.py)."""
import ctypes as C

import numpy as np
import pytest

import cantucci_b200 as cb
from cantucci_b200 import refine


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


class _OracleBulb(cb.Mandelbulb):
    """TEST-ONLY: a Mandelbulb whose DE comes from the CPU oracle, to check the host-side refinement
    logic without a GPU."""

    def batch_min_distance_from(self, pts, ctx=None):
        from oracle import oracle as O
        return O.batch_min_distance_from(O.mandelbulb(self.power, self.max_iters, self.bailout),
                                         np.asarray(pts, dtype=np.float32))


# ------------------------------------------------------------------ config 3 (host logic, CPU) -----
def test_config3_refinement_host_logic():
    spans, cam = refine.config3_spans(_OracleBulb.classic(6, 2.5), 6)
    # the camera sits on the -x tip of the bulb; the 25 focus rays all land in one leaf chain per level
    assert abs(cam.position[0] + 1.105) < 2e-3 and cam.position[1] == 0 and cam.position[2] == 0
    widths, counts = np.unique(np.round(spans[:, 3] - spans[:, 0], 5), return_counts=True)
    assert len(spans) == 176
    assert np.allclose(widths, [0.0375, 0.075, 0.15, 0.3, 0.6], atol=1e-5)      # depths 6..2
    assert counts.tolist() == [32, 28, 28, 28, 60]
    # leaves tile the bounding box exactly: volumes add up
    vol = np.prod((spans[:, 3:] - spans[:, :3]).astype(np.float64), axis=1).sum()
    assert abs(vol - 2.4 ** 3) < 1e-4
    import hashlib, json, os
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "config3_leaves.json")))
    assert gold["n_leaves"] == len(spans) and gold["spans_sha256"] == hashlib.sha256(spans.tobytes()).hexdigest()


# ------------------------------------------------------------------ config 3 (GPU parity) ----------
@pytest.mark.gpu
def test_config3_every_leaf_matches_oracle(oracle, ctx):
    bulb = cb.Mandelbulb.classic(6, 2.5)
    spans, _ = refine.config3_spans(bulb, 6, ctx)
    ref_spans, _ = refine.config3_spans(_OracleBulb.classic(6, 2.5), 6)
    assert np.array_equal(spans, ref_spans)                 # GPU DE drives the same refinement as the CPU DE
    batch, t = cb.generate_for_boxes(spans, bulb, 64, ctx)
    meshes, _ = oracle.generate_for_boxes_mt(oracle.mandelbulb(8, 6, 2.5), spans, 64)
    counts = []
    for k, (v, i, _) in enumerate(meshes):
        got = batch.mesh(k)
        assert np.array_equal(got.indices, i), k
        assert np.array_equal(got.vertices.view(np.uint32), v.view(np.uint32)), k
        counts.append(len(v))
    # "wildly different vertex counts": empty leaves next to dense ones
    assert min(counts) == 0 and max(counts) > 15000
    fast, _ = cb.generate_for_boxes(spans, cb.Mandelbulb.classic(6, 2.5, fast=True), 64, ctx)
    assert abs(len(fast.vertices) - len(batch.vertices)) <= len(batch.vertices) // 2000 + 8


@pytest.mark.gpu
def test_ray_march_kernel_matches_host_loop(oracle, ctx):
    """ctc_ray_march (get_focii's sphere tracing, mesh/mod.rs:229-241) against the same loop on the host
    with the oracle's DE: exact mode is bit-identical, misses are reported as None."""
    bulb = cb.Mandelbulb.classic(6, 2.5)
    cam = refine.Camera.default_orbit()
    got = refine.get_focii(bulb, cam, 5, ctx)
    want = refine.get_focii(_OracleBulb.classic(6, 2.5), cam, 5)
    assert len(got) == len(want) and 5 <= len(got) <= 25      # the outer rays of the 1-rad frustum miss the bulb
    assert np.array_equal(bits(got), bits(want))
    away = refine.Camera(np.array([-3.0, 0.0, 0.0]), np.array([-1.0, 0.0, 0.0]))     # looking away: no hits
    assert len(refine.get_focii(bulb, away, 3, ctx)) == 0


# ------------------------------------------------------------------ config 4 (power sweep) ---------
def _tiles16():
    return cb.tile_volume(cb.Span((-1.2, -1.2, -1.2), (1.2, 1.2, 1.2)), 16)


@pytest.mark.gpu
@pytest.mark.parametrize("power", [2, 4, 8, 16])
@pytest.mark.parametrize("max_iters", [32, 128])
def test_config4_power_sweep_subset_against_oracle(oracle, ctx, power, max_iters):
    """Eight tiles of the 1024^3 volume per (P, max_iters): grids against the oracle (bit-exact for
    P=8, tolerance otherwise), topology wherever the sign field agrees."""
    tiles = _tiles16()
    pick = tiles[np.random.default_rng(power * 1000 + max_iters).choice(len(tiles), 8, replace=False)]
    pick[0] = tiles[(8 * 16 + 8) * 16 + 3]      # one tile that certainly cuts the surface
    sh = oracle.mandelbulb(power, max_iters, 2.5)
    exact = cb.Mandelbulb(power, max_iters, 2.5)
    g_exact = cb.sample_grids(pick, exact, 64, ctx)
    g_fast = cb.sample_grids(pick, cb.Mandelbulb(power, max_iters, 2.5, fast=True), 64, ctx)
    flips_exact = flips_fast = total = 0
    for k, row in enumerate(pick):
        want, iters = oracle.sample_grid_iters(sh, oracle.make_span(row[:3], row[3:]), 64)
        if power == 8:
            assert np.array_equal(bits(g_exact[k]), bits(want)), k
        # Samples that escape within 6 iterations only: later escapes and interior samples sit on
        # chaotic orbits (a 1-ulp libm difference grows by ~P per iteration), so neither their escape
        # time nor the sign of ln(r) after 32..128 iterations can agree between two libm's (P != 8).
        esc = iters <= 6
        flips_exact += int(np.sum(((bits(g_exact[k]) >> 31) != (bits(want) >> 31)) & esc))
        flips_fast += int(np.sum(((bits(g_fast[k]) >> 31) != (bits(want) >> 31)) & esc))
        total += int(esc.sum())
        # far-field samples (escape at once) agree tightly in every mode
        far = want > 0.5
        if far.any():
            assert np.max(np.abs(g_fast[k][far] - want[far]) / want[far]) < 1e-5
            assert np.max(np.abs(g_exact[k][far] - want[far]) / want[far]) < 1e-5
    # samples whose escape iteration sits on the bailout boundary may still flip
    assert flips_exact / total < (0 if power == 8 else 2e-3) + 1e-12
    assert flips_fast / total < 2e-3
    if power == 8:
        batch, _ = cb.generate_for_boxes(pick, exact, 64, ctx)
        meshes, _ = oracle.generate_for_boxes_mt(sh, pick, 64)
        for k, m in enumerate(meshes):
            if m is None:        # the reference panicked (lerp assert) -- cannot happen with status OK
                continue
            assert np.array_equal(batch.mesh(k).indices, m[1]) and np.array_equal(batch.mesh(k).vertices.view(np.uint32), m[0].view(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("power,max_iters", [(8, 32), (8, 128), (2, 32), (16, 128)])
def test_config4_full_1024_volume_properties(ctx, power, max_iters):
    """The whole 1024^3 volume (4096 spans) in fast mode: size-independent properties."""
    import torch
    from cantucci_b200 import _lib
    from cantucci_b200.scheduler import DeviceMesher
    tiles = _tiles16()
    sh = cb.Mandelbulb(power, max_iters, 2.5, fast=True)._ctc_shape()
    dev = torch.device("cuda", 0)
    m = DeviceMesher(ctx, torch, dev, 40_000_000, 240_000_000, len(tiles))
    runs = []
    for _ in range(2):
        m.launch(sh, tiles, 64)
        nv, ni, t = m.result(allow_lerp_assert=True)
        assert 0 < nv <= m.vcap and ni % 6 == 0 and ni <= m.icap
        v_off = m.v_off[: len(tiles) + 1].cpu().numpy(); i_off = m.i_off[: len(tiles) + 1].cpu().numpy()
        assert v_off[-1] == nv and i_off[-1] == ni and np.all(np.diff(v_off) >= 0) and np.all(np.diff(i_off) >= 0)
        # (a span whose lerp factor left [0,1] -- NaN distances next to the surface, where the reference panics --
        # carries NaN positions, in exact mode and, because suspects are re-evaluated exactly, in fast mode too)
        pos = m.v[:nv, :3]
        runs.append((nv, ni, int(m.i[:ni].to(torch.int64).sum()), int(pos.view(torch.int32).to(torch.int64).sum()),
                     float(torch.nan_to_num(pos, nan=0.0).abs().max())))
    assert runs[0] == runs[1]                                   # idempotent, deterministic
    assert runs[0][4] <= 1.2 + 2 * 0.15 / 64 + 1e-6             # vertices inside the skirted volume
    # span-local indices stay inside their span's vertex range
    cnt_v = torch.from_numpy(np.diff(v_off)).to(dev)
    span_of_index = torch.repeat_interleave(torch.arange(len(tiles), device=dev), torch.from_numpy(np.diff(i_off)).to(dev))
    assert bool((m.i[:ni].to(torch.int64) < cnt_v[span_of_index]).all())


# ------------------------------------------------------------------ config 5 (4096^3 as 64^3 spans) -
@pytest.mark.gpu
def test_config5_4096_volume_sharded_subset_and_oracle_sample(oracle, ctx):
    """Config 5 is 262 144 spans; the full volume is exercised by `bench.py --tiles 64`.  Here: a
    4096-span slab of it through the device API, plus 24 of its spans against the oracle."""
    tiles = cb.tile_volume(cb.Span((-1.2, -1.2, -1.2), (1.2, 1.2, 1.2)), 64)
    assert tiles.shape == (262144, 6)
    slab = np.ascontiguousarray(tiles[10 * 4096: 11 * 4096])            # an x-slab cutting the bulb, clear of the axis planes
    bulb = cb.Mandelbulb.classic(6, 2.5)
    batch, t = cb.generate_for_boxes(slab, bulb, 64, ctx)
    assert t.vertices == len(batch.vertices) > 1_000_000
    pick = np.random.default_rng(5).choice(len(slab), 24, replace=False)
    meshes, _ = oracle.generate_for_boxes_mt(oracle.mandelbulb(8, 6, 2.5), slab[pick], 64)
    for k, (v, i, _) in zip(pick, meshes):
        got = batch.mesh(int(k))
        assert np.array_equal(got.indices, i) and np.array_equal(got.vertices.view(np.uint32), v.view(np.uint32))


# ------------------------------------------------------------------ the caller: ShapeMesh::update ---
@pytest.mark.gpu
def test_shape_mesh_update_flies_in_and_matches_oracle(oracle, ctx):
    """ShapeMesh::update (mesh/mod.rs:82-178) driven by a camera that approaches the bulb: the first
    frame meshes the 64 startup leaves, later frames split the leaves around the focus points and mesh
    the new children.  Every Ready leaf must equal the oracle's generate_for_box of its span."""
    sm = cb.ShapeMesh(cb.Mandelbulb.classic(6, 2.5), ctx, resolution=32)
    cam = refine.Camera.default_orbit()
    assert sm.update(cam) == 64                         # frame 1: the startup octree
    assert sm.update(cam) == 0                          # camera 1.9 away from the surface: nothing splits (threshold 1.2)
    meshed = []
    for x in (-2.0, -1.5, -1.2):                        # fly in along the orbit ray
        cam = refine.Camera(np.array([x, 0.0, 0.0]), np.array([1.0, 0.0, 0.0]))
        meshed.append(sm.update(cam))
    assert sum(meshed) > 0 and all(m % 8 == 0 for m in meshed)      # splits create 8 children each
    ready = sm.ready_meshes()
    assert len(ready) == 64 + sum(meshed) - sum(meshed) // 8
    sh = oracle.mandelbulb(8, 6, 2.5)
    spans = cb.spans_array([s for s, _ in ready])
    want, _ = oracle.generate_for_boxes_mt(sh, spans, 32)
    for (span, got), (v, i, _) in zip(ready, want):
        assert np.array_equal(got.indices, i)
        assert np.array_equal(got.vertices.view(np.uint32), v.view(np.uint32))


# ------------------------------------------------------------------ maximum size the ABI accepts ----
@pytest.mark.gpu
def test_resolution_1024_single_span_grid(oracle, ctx):
    """R = 1024 is the largest resolution (1025^3 = 1.08 G samples, 4.3 GB): pass 1 into a device
    buffer, spot-checked against the oracle; R = 2048 is refused as an argument error."""
    import torch
    from cantucci_b200 import _lib
    bulb = cb.Mandelbulb.classic(6, 2.5)
    sh = bulb._ctc_shape()
    bbox = np.array([[-1.2, -1.2, -1.2, 1.2, 1.2, 1.2]], dtype=np.float32)
    n = 1025
    g = torch.empty((n ** 3,), dtype=torch.float32, device="cuda:0")
    ctx.check(_lib.lib().ctc_sample_grids_device(ctx.handle, C.byref(sh), bbox.ctypes.data, 1, 1024, g.data_ptr()))
    ctx.synchronize()
    rng = np.random.default_rng(7)
    ijk = np.concatenate([rng.integers(0, n, size=(3000, 3)), [[0, 0, 0], [1024, 1024, 1024], [512, 512, 512], [1024, 0, 513]]])
    flat = (ijk[:, 0] * n + ijk[:, 1]) * n + ijk[:, 2]
    got = g[torch.from_numpy(flat).to("cuda:0")].cpu().numpy()
    fr = np.float32(1024)
    ov = np.float32(np.float32(2.4) / fr)
    s0 = np.float32(np.float32(-1.2) + (-ov)); e0 = np.float32(np.float32(1.2) + ov)
    pts = (s0 + np.float32(e0 - s0) * (ijk.astype(np.float32) / fr)).astype(np.float32)
    want = oracle.batch_min_distance_from(oracle.mandelbulb(8, 6, 2.5), pts)
    assert np.array_equal(bits(got), bits(want))
    assert bits(got)[-2] == 0xFFC00000                      # the origin
    rc = _lib.lib().ctc_sample_grids_device(ctx.handle, C.byref(sh), bbox.ctypes.data, 1, 2048, g.data_ptr())
    assert rc == _lib.CTC_ERR_INVALID_ARGUMENT


@pytest.mark.gpu
def test_de_bound_culling_only_drops_spans_the_oracle_finds_empty(oracle, ctx):
    """cull_spans drops a span when the distance estimate at its centre exceeds twice the half-
    diagonal of its skirt-expanded box.  Every culled span must have an EMPTY mesh in the CPU oracle -- on the
    startup octree (config 1), the depth-6 refinement (config 3) and the 1024^3 volume -- and meshing with
    cull = True must return exactly the un-culled result.  Sphere: spans entirely inside are dropped too."""
    import cantucci_b200 as cb
    from cantucci_b200 import refine
    shape = cb.Mandelbulb.classic(6, 2.5)
    osh = oracle.mandelbulb(8, 6, 2.5)
    startup = cb.spans_array([n.span for n in cb.startup_tree(shape.bounding_box()).leaves()])
    leaves, _ = refine.config3_spans(shape, 6, ctx)
    volume = cb.tile_volume(cb.Span((-1.2,) * 3, (1.2,) * 3), 16)
    culled_total = 0
    for name, spans in (("config1", startup), ("config3", leaves), ("volume", volume)):
        keep = cb.cull_spans(spans, shape, 64, ctx)
        dropped = np.ascontiguousarray(spans[~keep])
        culled_total += len(dropped)
        if len(dropped):
            v, i, v_off, i_off, planes, panicked, _ = oracle.generate_for_boxes_flat_mt(osh, dropped, 64, oracle.hardware_threads(), signs=True)
            assert int(v_off[-1]) == 0 and int(i_off[-1]) == 0, (name, int(v_off[-1]))
            assert not planes.any(), name                      # every sample of a culled span is outside
    assert culled_total > 0
    keep = cb.cull_spans(volume, shape, 64, ctx)
    assert 0.05 < (~keep).mean() < 0.9
    sub = volume[::7]
    a, _ = cb.generate_for_boxes(sub, shape, 64, ctx)
    b, _ = cb.generate_for_boxes(sub, shape, 64, ctx, cull=True)
    assert np.array_equal(a.v_off, b.v_off) and np.array_equal(a.i_off, b.i_off)
    assert np.array_equal(a.indices, b.indices) and np.array_equal(a.vertices.view(np.uint32), b.vertices.view(np.uint32))
    # Sphere (two-sided bound): spans far outside AND spans deep inside go
    sph = cb.Sphere((0.0, 0.0, 0.0), 1.0)
    tiles = cb.tile_volume(cb.Span((-1.2,) * 3, (1.2,) * 3), 8)
    keep = cb.cull_spans(tiles, sph, 32, ctx)
    centre = 0.5 * (tiles[:, :3] + tiles[:, 3:])
    rad = np.linalg.norm(centre, axis=1)
    assert (~keep)[rad < 0.3].all() and keep[np.abs(rad - 1.0) < 0.1].all()
    a, _ = cb.generate_for_boxes(tiles, sph, 32, ctx)
    b, _ = cb.generate_for_boxes(tiles, sph, 32, ctx, cull=True)
    assert np.array_equal(a.v_off, b.v_off) and np.array_equal(a.indices, b.indices)
    assert np.array_equal(np.diff(a.v_off.astype(np.int64))[~keep], np.zeros((~keep).sum(), dtype=np.int64))


@pytest.mark.gpu
def test_render_pixels_equal_ray_march_and_the_oracle_loop(oracle, ctx):
    """ctc_render sphere-traces one ray per pixel with get_focii's loop (mesh/mod.rs:229-241).  In
    exact mode every pixel must equal ctc_ray_march on the bit-identical ray (formed on the host by pixel_rays), and
    a sample of pixels must equal the loop run with the CPU oracle's DE."""
    import cantucci_b200 as cb
    from cantucci_b200 import _lib
    W, H = 96, 64
    shape = cb.Mandelbulb.classic(6, 2.5)
    cam = cb.look_at_rays((1.6, 1.1, 2.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 45.0, W, H)
    img = cb.render(shape, cam, W, H, 100, 1e-4, ctx)
    origins, dirs = cb.pixel_rays(cam, W, H)
    pos = np.empty_like(origins); hit = np.zeros(len(origins), dtype=np.uint32)
    sh = shape._ctc_shape()
    ctx.check(_lib.lib().ctc_ray_march(ctx.handle, C.byref(sh), origins.ctypes.data, dirs.ctypes.data, len(origins), 100,
                                       C.c_float(1e-4), pos.ctypes.data, hit.ctypes.data))
    flat = img.reshape(-1, 4)
    got_hit = flat[:, 3] >= 0
    assert 0.1 < got_hit.mean() < 0.9                                   # the bulb fills part of the frame
    assert np.array_equal(got_hit, hit.astype(bool))
    assert np.array_equal(flat[got_hit, :3].view(np.uint32), pos[got_hit].view(np.uint32))
    # the loop itself with the CPU oracle's DE, on a few pixels (hits and misses)
    osh = oracle.mandelbulb(8, 6, 2.5)
    f = np.float32
    for k in list(np.nonzero(got_hit)[0][::97][:12]) + list(np.nonzero(~got_hit)[0][::211][:4]):
        p, d = origins[k].copy(), dirs[k]
        ok, t = False, f(0.0)
        for _ in range(100):
            dist = f(oracle.batch_min_distance_from(osh, p.reshape(1, 3))[0])
            p = (p + (d * dist).astype(np.float32)).astype(np.float32)
            t = f(t + dist)
            if dist < f(1e-4):
                ok = True; break
            if not (t < f(1e6)):
                break
        assert ok == bool(got_hit[k]), k
        if ok:
            assert np.array_equal(p.view(np.uint32), flat[k, :3].view(np.uint32)), k
            assert f(flat[k, 3]).view(np.uint32) == t.view(np.uint32), k
    # fast mode renders the same silhouette up to a handful of edge pixels
    fast = cb.render(cb.Mandelbulb.classic(6, 2.5, fast=True), cam, W, H, 100, 1e-4, ctx)
    assert np.mean((fast[..., 3] >= 0) != (img[..., 3] >= 0)) < 0.01
