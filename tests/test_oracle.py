"""CPU tests: the C oracle against an independent numpy restatement, analytic
Sphere facts, its own frozen golden vectors, and the reference's quirks.

PARITY UNPINNED by the reference itself: it ships no expected outputs
(SURVEY.md 8c).  These tests are what pins the oracle instead."""
import ctypes
import hashlib
import json
import os

import numpy as np
import pytest

import np_restatement as NP
from conftest import BENCH_POINTS, startup_leaves

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def bits(a):
    return np.asarray(a, dtype=np.float32).view(np.uint32)


def rand_points(n, seed, lo=-1.3, hi=1.3):
    return np.random.default_rng(seed).uniform(lo, hi, size=(n, 3)).astype(np.float32)


# ---------------------------------------------------------------- DE ------
@pytest.mark.parametrize("max_iters,bailout", [(6, 2.5), (8, 5.0), (32, 2.5)])
def test_p8_de_matches_numpy_restatement_bit_exact(oracle, max_iters, bailout):
    pts = np.concatenate([BENCH_POINTS, rand_points(4000, 1)])
    sh = oracle.mandelbulb(8, max_iters, bailout)
    got = oracle.batch_min_distance_from(sh, pts)
    want, iters, r, dr = NP.mandelbulb_de(pts, 8, max_iters, bailout)
    assert np.array_equal(bits(got), bits(want))
    info = [oracle.min_distance_from_info(sh, p)[1] for p in pts[:200]]
    assert [i.iters for i in info] == list(iters[:200])
    assert np.array_equal(bits([i.r for i in info]), bits(r[:200]))
    assert np.array_equal(bits([i.dr for i in info]), bits(dr[:200]))


@pytest.mark.parametrize("power", [2, 4, 16, 3])
def test_generic_power_de_matches_numpy_restatement_bit_exact(oracle, power):
    pts = np.concatenate([BENCH_POINTS, rand_points(600, 2)])
    sh = oracle.mandelbulb(power, 12, 2.5)
    got = oracle.batch_min_distance_from(sh, pts)
    want, *_ = NP.mandelbulb_de(pts, power, 12, 2.5)
    assert np.array_equal(bits(got), bits(want))


def test_z_axis_and_origin_special_cases(oracle):
    sh = oracle.mandelbulb(8, 6, 2.5)
    pts = np.array([[0, 0, 0], [0, 0, 0.5], [0, 0, -0.5], [-0.0, 0.0, 0.9], [0, -0.0, -1.1], [0, 0, 1.2]], dtype=np.float32)
    got = oracle.batch_min_distance_from(sh, pts)
    want, *_ = NP.mandelbulb_de(pts, 8, 6, 2.5)
    # the origin is 0/0 in rotate_on_z_axis: x86's default NaN has the sign bit set,
    # so is_sign_positive() is false and the sample classifies as inside
    assert bits(got)[0] == 0xFFC00000
    assert np.array_equal(bits(got)[1:], bits(want)[1:])


def test_polynomial_and_trig_power8_are_not_the_same_map(oracle):
    """The reference benches rotate_inner_generic and rotate_inner_p8_scalar side by side
    (mandelbulb.rs:479-508) as if equivalent.  They are not: the polynomial uses
    (cos 8t cos 8p, cos 8t sin 8p, sin 8t) and a '- y8' term, the trig form
    (sin 8t cos 8p, sin 8t sin 8p, cos 8t).  rotate::<8> only ever calls the polynomial,
    which is what this repo reproduces."""
    diffs = []
    for p in BENCH_POINTS:
        a = oracle.rotate(8, p, "p8_scalar")
        b = oracle.rotate(8, p, "generic")
        diffs.append(np.abs(a - b).max() / max(np.abs(a).max(), 1e-6))
    assert np.median(diffs) > 0.1
    # magnitudes agree (both are |p|^8 up to the y8 slip), the directions do not
    p = BENCH_POINTS[3]
    assert np.array_equal(bits(oracle.rotate(8, p, "dispatch")), bits(oracle.rotate(8, p, "p8_scalar")))


def test_generic_de_agrees_with_the_glsl_statement_of_the_math(oracle):
    """src/shape/mandelbulb.frag:1-39 states the generic-power DE a second time (GLSL, no z-axis special
    case).  Evaluated here in float64 it must agree with the oracle's f32 generic path to f32 accuracy on
    samples that escape early (later escapes amplify rounding)."""
    def glsl_de(p, power, max_iters, bailout):
        z = np.array(p, dtype=np.float64); dr = 1.0; r = 0.0
        for _ in range(max_iters):
            r = np.linalg.norm(z)
            if r > bailout:
                break
            theta = np.arccos(z[2] / r); phi = np.arctan2(z[1], z[0])
            dr = r ** (power - 1.0) * power * dr + 1.0
            zr = r ** power
            theta *= power; phi *= power
            z = zr * np.array([np.sin(theta) * np.cos(phi), np.sin(phi) * np.sin(theta), np.cos(theta)]) + p
        return 0.5 * np.log(r) * r / dr
    pts = rand_points(400, 9, 0.9, 1.3)         # outside the bulb: escape within a few iterations
    for power in (2, 4, 16):
        sh = oracle.mandelbulb(power, 12, 2.5)
        for p in pts:
            d, info = oracle.min_distance_from_info(sh, p)
            if info.bailed and info.iters <= 2 and info.min_margin > 1e-3:
                assert abs(d - glsl_de(p.astype(np.float64), power, 12, 2.5)) <= 2e-5 * abs(d) + 1e-7, (power, p)


def test_sphere_de_is_analytic(oracle):
    sh = oracle.sphere((0.1, -0.2, 0.3), 0.75)
    pts = rand_points(1000, 3)
    got = oracle.batch_min_distance_from(sh, pts)
    exact = np.linalg.norm(pts.astype(np.float64) - np.array([0.1, -0.2, 0.3]), axis=1) - 0.75
    assert np.allclose(got, exact, atol=2e-7)
    assert np.array_equal(bits(got), bits(NP.sphere_de(pts, (0.1, -0.2, 0.3), 0.75)))


def test_logf_restatement_equals_this_hosts_libm(oracle):
    """The CUDA exact mode ports glibc's logf (FMA build).  Check the restatement against the
    libm this host actually runs, so a non-FMA host is noticed."""
    rng = np.random.default_rng(4)
    xs = np.concatenate([
        rng.uniform(0, 2.6, 200000), rng.uniform(0.9, 1.1, 50000), np.exp(rng.uniform(-80, 80, 50000)),
        [0.0, 1.0, 2.5, np.inf, 1e-45, 1e-39],
    ]).astype(np.float32)
    libm = ctypes.CDLL("libm.so.6")
    libm.logf.restype = ctypes.c_float
    libm.logf.argtypes = [ctypes.c_float]
    L = oracle.lib()
    bad = sum(1 for x in xs
              if np.float32(L.orc_logf_glibc_fma(float(x))).view(np.uint32) != np.float32(libm.logf(float(x))).view(np.uint32))
    assert bad == 0


# ------------------------------------------------------------ mesher ------
def _cmp_mesh(oracle, shape, de_fn, start, end, R):
    span = oracle.make_span(start, end)
    v, idx, _ = oracle.generate_for_box(shape, span, R)
    V, I, dists = NP.naive_surface_nets(de_fn, start, end, R)
    grid = oracle.sample_grid(shape, span, R)
    assert np.array_equal(bits(grid), bits(dists.ravel()))
    assert len(v) == len(V) and np.array_equal(idx, I)
    got = np.concatenate([v["position"], v["normal"], v["distance_from_surface"][:, None]], axis=1)
    assert np.array_equal(bits(got), bits(V))
    return v, idx


def test_mesher_matches_numpy_restatement_on_a_mandelbulb_span(oracle):
    sh = oracle.mandelbulb(8, 6, 2.5)
    v, idx = _cmp_mesh(oracle, sh, lambda p: NP.mandelbulb_de(p, 8, 6, 2.5)[0], (0.0, 0.0, 0.0), (0.6, 0.6, 0.6), 8)
    assert len(v) > 50 and len(idx) > 300


def test_mesher_matches_numpy_restatement_on_a_sphere_and_is_analytic(oracle):
    c, rad = (0.05, -0.02, 0.01), 0.8
    sh = oracle.sphere(c, rad)
    v, idx = _cmp_mesh(oracle, sh, lambda p: NP.sphere_de(p, c, rad), (-1.0, -1.0, -1.0), (1.0, 1.0, 1.0), 8)
    # vertices sit within a cell of the sphere, normals point outward (exact DE => gradient = radial)
    pos = v["position"].astype(np.float64) - np.array(c)
    rr = np.linalg.norm(pos, axis=1)
    assert np.all(np.abs(rr - rad) < 2.25 / 8)
    cosang = np.sum(pos / rr[:, None] * v["normal"], axis=1)
    assert np.all(cosang > 0.999)
    assert np.allclose(np.linalg.norm(v["normal"], axis=1), 1.0, atol=1e-6)
    # closed surface: every edge of the triangle soup is shared by exactly two triangles
    tri = idx.reshape(-1, 3)
    e = np.sort(np.concatenate([tri[:, [0, 1]], tri[:, [1, 2]], tri[:, [2, 0]]]), axis=1)
    _, counts = np.unique(e, axis=0, return_counts=True)
    assert np.all(counts == 2)
    # triangles wind counter-clockwise seen from outside (view.rs:122: CCW front faces)
    P = v["position"].astype(np.float64)
    nrm = np.cross(P[tri[:, 1]] - P[tri[:, 0]], P[tri[:, 2]] - P[tri[:, 0]])
    cen = P[tri].mean(axis=1) - np.array(c)
    assert np.all(np.sum(nrm * cen, axis=1) > 0)


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_mesher_matches_numpy_restatement_on_random_shapes_and_spans(oracle, seed):
    """Randomised cross-check of the two independent restatements: random anisotropic spans, Sphere
    with random centre/radius or Mandelbulb with P in {2, 4, 8}, R in {4, 8}."""
    rng = np.random.default_rng(100 + seed)
    R = int(rng.choice([4, 8]))
    start = rng.uniform(-1.1, 0.2, 3).astype(np.float32)
    end = (start + rng.uniform(0.3, 0.9, 3)).astype(np.float32)
    if seed % 2:
        c = rng.uniform(-0.3, 0.3, 3).astype(np.float32); rad = float(np.float32(rng.uniform(0.5, 0.9)))
        sh, de = oracle.sphere(tuple(c), rad), (lambda p: NP.sphere_de(p, c, rad))
    else:
        P = int(rng.choice([2, 4, 8])); iters = int(rng.integers(3, 8))
        sh, de = oracle.mandelbulb(P, iters, 2.5), (lambda p: NP.mandelbulb_de(p, P, iters, 2.5)[0])
    try:
        V, I, dists = NP.naive_surface_nets(de, start, end, R)
    except AssertionError:          # the numpy restatement hit the lerp assert: the oracle must report the panic too
        with pytest.raises(AssertionError):
            oracle.generate_for_box(sh, oracle.make_span(start, end), R)
        return
    v, idx, _ = oracle.generate_for_box(sh, oracle.make_span(start, end), R)
    assert np.array_equal(bits(oracle.sample_grid(sh, oracle.make_span(start, end), R)), bits(dists.ravel()))
    assert np.array_equal(idx, I) and len(v) == len(V)
    got = np.concatenate([v["position"], v["normal"], v["distance_from_surface"][:, None]], axis=1)
    assert np.array_equal(bits(got), bits(V))


def test_argument_asserts(oracle):
    sh = oracle.mandelbulb(8, 6, 2.5)
    with pytest.raises(AssertionError):
        oracle.generate_for_box(sh, oracle.make_span((0, 0, 0), (1, 1, 0)), 8)      # start < end
    with pytest.raises(AssertionError):
        oracle.generate_for_box(sh, oracle.make_span((0, 0, 0), (1, 1, 1)), 12)     # power of two
    with pytest.raises(AssertionError):
        oracle.generate_for_box(sh, oracle.make_span((0, 0, 0), (1, 1, 1)), 1)      # GridTable size >= 2


def test_nan_weight_panics_like_the_reference_lerp_assert(oracle):
    """R=2 over (-2..2)^3 puts the lattice at -4, 0, 4: the origin is NaN (sign bit set =>
    'inside'), every other sample bails at once and is positive.  The edges leaving the origin
    get a NaN lerp factor, which the reference's lerp assert (math.rs:19) turns into a panic."""
    sh = oracle.mandelbulb(8, 6, 2.5)
    span = oracle.make_span((-2, -2, -2), (2, 2, 2))
    g = oracle.sample_grid(sh, span, 2)
    assert bits(g)[13] == 0xFFC00000 and np.all(np.delete(g, 13) > 0)
    with pytest.raises(AssertionError, match="lerp"):
        oracle.generate_for_box(sh, span, 2)


# ------------------------------------------------------------ golden ------
def test_golden_bench_points(oracle):
    gold = json.load(open(os.path.join(GOLD, "bench_points_de.json")))
    for key, want in gold.items():
        power, iters, bail = key.split("_")
        sh = oracle.mandelbulb(int(power[1:]), int(iters[1:]), float(bail[1:]))
        got = oracle.batch_min_distance_from(sh, BENCH_POINTS)
        assert [f"{b:08x}" for b in bits(got)] == want, key


def test_golden_startup_octree_counts_and_hashes(oracle):
    """Config 1: the 64 startup leaves at R=64 (mesh/mod.rs:52-56,133)."""
    gold = json.load(open(os.path.join(GOLD, "config1_startup.json")))
    spans = startup_leaves()
    assert hashlib.sha256(spans.tobytes()).hexdigest() == gold["spans_sha256"]
    sh = oracle.mandelbulb(8, 6, 2.5)
    meshes, _ = oracle.generate_for_boxes_mt(sh, spans, 64)
    assert [len(m[0]) for m in meshes] == gold["vertices_per_span"]
    assert [len(m[1]) // 6 for m in meshes] == gold["quads_per_span"]
    assert sum(gold["vertices_per_span"]) == 550428 and sum(gold["quads_per_span"]) == 559686
    hv, hi = hashlib.sha256(), hashlib.sha256()
    for v, i, _ in meshes:
        hv.update(v.tobytes()); hi.update(i.tobytes())
    assert hv.hexdigest() == gold["vertices_sha256"]
    assert hi.hexdigest() == gold["indices_sha256"]


def test_golden_small_meshes(oracle):
    z = np.load(os.path.join(GOLD, "small_meshes.npz"))
    sh = oracle.mandelbulb(8, 6, 2.5)
    v, i, _ = oracle.generate_for_box(sh, oracle.make_span((0.0, 0.0, 0.0), (0.6, 0.6, 0.6)), 16)
    assert np.array_equal(v.view(np.uint32).reshape(-1, 7), z["bulb_v"]) and np.array_equal(i, z["bulb_i"])
    sp = oracle.sphere((0.0, 0.0, 0.0), 0.9)
    v, i, _ = oracle.generate_for_box(sp, oracle.make_span((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0)), 16)
    assert np.array_equal(v.view(np.uint32).reshape(-1, 7), z["sphere_v"]) and np.array_equal(i, z["sphere_i"])


# ------------------------------------------------------ host octree -------
def test_host_octree_mirror_matches_oracle_span_maths(oracle):
    import cantucci_b200 as cb
    root = cb.Span((-1.2, -1.2, -1.2), (1.2, 1.2, 1.2))
    kids = cb.create_spans(root)
    okids = oracle.create_spans(oracle.make_span(root.start, root.end))
    for a, b in zip(kids, okids):
        assert np.array_equal(a.as_row(), np.array([*b.start, *b.end], dtype=np.float32))
    # child order = (x,y,z) bits with z lowest (octree/mod.rs:131-138)
    assert kids[1].start[2] == 0.0 and kids[1].start[0] == np.float32(-1.2)
    assert kids[4].start[0] == 0.0 and kids[4].start[2] == np.float32(-1.2)
    tree = cb.startup_tree(root)
    leaves = tree.leaves()
    assert len(leaves) == 64
    # iter_mut is a LIFO DFS: first leaf is child 7 of child 7 (octree/iter.rs:88-106)
    assert leaves[0].span == cb.create_spans(kids[7])[7]
    assert leaves[-1].span == cb.create_spans(kids[0])[0]
    assert tree.leaf_around((0.01, 0.01, 0.01)).span == cb.create_spans(kids[7])[0]
    assert tree.leaf_around((5, 0, 0)) is None
