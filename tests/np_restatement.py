"""An INDEPENDENT numpy-float32 restatement of the reference's DE and mesher.

Written from the Rust sources (not from oracle/cantucci_oracle.c) so that the C
oracle is pinned by a second implementation.  numpy f32 array arithmetic is
IEEE and never fused; libm's logf/acosf/atan2f/sinf/cosf are called through
ctypes so transcendental results are glibc's (numpy ships its own SIMD logf).
"""
import ctypes
import ctypes.util

import numpy as np

f32 = np.float32
_libm = ctypes.CDLL(ctypes.util.find_library("m"))
for _n in ("logf", "acosf", "sinf", "cosf"):
    getattr(_libm, _n).restype = ctypes.c_float
    getattr(_libm, _n).argtypes = [ctypes.c_float]
_libm.atan2f.restype = ctypes.c_float
_libm.atan2f.argtypes = [ctypes.c_float, ctypes.c_float]


def _m1(name, x):
    fn = getattr(_libm, name)
    return np.array([fn(float(v)) for v in np.asarray(x, dtype=f32).ravel()], dtype=f32).reshape(np.shape(x))


def _atan2(y, x):
    return np.array([_libm.atan2f(float(a), float(b)) for a, b in zip(np.ravel(y), np.ravel(x))],
                    dtype=f32).reshape(np.shape(x))


def powi(x, n):
    """llvm.powi with constant exponent: binary square-and-multiply (mandelbulb.rs:73,121,136)."""
    x = np.asarray(x, dtype=f32)
    if n == 0:
        return np.ones_like(x)
    res, cur = None, x
    while n:
        if n & 1:
            res = cur if res is None else (res * cur).astype(f32)
        n >>= 1
        if n:
            cur = (cur * cur).astype(f32)
    return res


def magnitude(x, y, z):
    # _mm_dp_ps(v, v, 0x71) + _mm_sqrt_ss (mandelbulb.rs:411-418)
    return np.sqrt(((x * x + y * y) + (z * z + f32(0.0))).astype(f32)).astype(f32)


def rotate_p8(x, y, z):
    # mandelbulb.rs:148-200
    x2 = x * x; x4 = x2 * x2; x6 = x4 * x2; x8 = x4 * x4
    y2 = y * y; y4 = y2 * y2; y6 = y4 * y2; y8 = y4 * y4
    z2 = z * z; z4 = z2 * z2; z6 = z4 * z2; z8 = z4 * z4
    rxy2 = x2 + y2; rxy4 = rxy2 * rxy2; rxy6 = rxy2 * rxy4; rxy8 = rxy4 * rxy4
    c28, c70, c7, c8, c6, one = f32(28.0), f32(70.0), f32(7.0), f32(8.0), f32(6.0), f32(1.0)
    a = one + (z8 - c28 * z6 * rxy2 + c70 * z4 * rxy4 - c28 * z2 * rxy6) / rxy8
    nx = a * (x8 - c28 * x6 * y2 + c70 * x4 * y4 - c28 * x2 * y6 - y8)
    ny = c8 * a * x * y * (x6 - c7 * x4 * y2 + c7 * x2 * y4 - y6)
    nz = c8 * z * np.sqrt(rxy2) * (z2 - rxy2) * (z4 - c6 * z2 * rxy2 + rxy4)
    return nx.astype(f32), ny.astype(f32), nz.astype(f32)


def rotate_generic(P, x, y, z):
    # mandelbulb.rs:128-146
    old_radius = magnitude(x, y, z)
    theta = _m1("acosf", z / old_radius)
    phi = _atan2(y, x)
    new_radius = powi(old_radius, P)
    theta = theta * f32(P)
    phi = phi * f32(P)
    vx = _m1("sinf", theta) * _m1("cosf", phi)
    vy = _m1("sinf", phi) * _m1("sinf", theta)
    vz = _m1("cosf", theta)
    return (vx * new_radius).astype(f32), (vy * new_radius).astype(f32), (vz * new_radius).astype(f32)


def rotate_on_z_axis(P, x, y, z):
    # mandelbulb.rs:114-126
    old_radius = magnitude(x, y, z)
    theta = _m1("acosf", z / old_radius)
    new_radius = powi(old_radius, P)
    theta = theta * f32(P)
    zero = np.zeros_like(z)
    return zero, zero, (new_radius * _m1("cosf", theta)).astype(f32)


def mandelbulb_de(points, P, max_iters, bailout):
    """Mandelbulb::<P>::min_distance_from (mandelbulb.rs:59-79), one point at a time
    vectorised with an 'alive' mask.  Returns (de, iters, r, dr)."""
    with np.errstate(all="ignore"):
        p = np.asarray(points, dtype=f32).reshape(-1, 3)
        n = p.shape[0]
        px, py, pz = p[:, 0].copy(), p[:, 1].copy(), p[:, 2].copy()
        zx, zy, zz = px.copy(), py.copy(), pz.copy()
        dr = np.ones(n, dtype=f32)
        r = np.zeros(n, dtype=f32)
        iters = np.zeros(n, dtype=np.int64)
        alive = np.ones(n, dtype=bool)
        bail = f32(bailout)
        for _ in range(max_iters):
            if not alive.any():
                break
            idx = np.nonzero(alive)[0]
            x, y, z = zx[idx], zy[idx], zz[idx]
            rr = magnitude(x, y, z)
            r[idx] = rr
            out = rr > bail
            alive[idx[out]] = False
            keep = ~out
            idx, x, y, z, rr = idx[keep], x[keep], y[keep], z[keep], rr[keep]
            if idx.size == 0:
                break
            dr[idx] = (powi(rr, P - 1) * f32(P) * dr[idx] + f32(1.0)).astype(f32)
            on_axis = (x == 0) & (y == 0)          # +-0.0 both compare equal to 0
            nx = np.empty_like(x); ny = np.empty_like(x); nz = np.empty_like(x)
            if on_axis.any():
                nx[on_axis], ny[on_axis], nz[on_axis] = rotate_on_z_axis(P, x[on_axis], y[on_axis], z[on_axis])
            off = ~on_axis
            if off.any():
                if P == 8:
                    nx[off], ny[off], nz[off] = rotate_p8(x[off], y[off], z[off])
                else:
                    nx[off], ny[off], nz[off] = rotate_generic(P, x[off], y[off], z[off])
            zx[idx] = nx + px[idx]; zy[idx] = ny + py[idx]; zz[idx] = nz + pz[idx]
            iters[idx] += 1
        ln_r = (_m1("logf", r) * r).astype(f32)
        de = (f32(0.5) * ln_r / dr).astype(f32)
        return de, iters, r, dr


def sphere_de(points, center, radius):
    # sphere.rs:33-35
    p = np.asarray(points, dtype=f32).reshape(-1, 3)
    d = np.asarray(center, dtype=f32)[None, :] - p
    return (np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]) - f32(radius)).astype(f32)


def sign_positive(v):
    return (np.asarray(v, dtype=f32).view(np.uint32) >> 31) == 0


def naive_surface_nets(de_fn, start, end, R):
    """MeshBuffer::naive_surface_nets (mesh/buffer.rs:58-391) with scalar Python loops.
    de_fn(points[n,3]) -> f32[n].  Returns (vertices [V,7] f32, indices u32)."""
    with np.errstate(all="ignore"):
        start = np.asarray(start, dtype=f32); end = np.asarray(end, dtype=f32)
        fr = f32(R)
        overflow = (end - start) / fr                      # :65
        s0 = start + (-overflow); e0 = end + overflow      # :66
        across = e0 - s0                                   # :77
        n = R + 1
        ii = np.arange(n, dtype=f32) / fr                  # :79
        X, Y, Z = np.meshgrid(ii, ii, ii, indexing="ij")
        pts = np.stack([s0[0] + across[0] * X, s0[1] + across[1] * Y, s0[2] + across[2] * Z], -1).astype(f32)
        dists = de_fn(pts.reshape(-1, 3)).reshape(n, n, n)  # x-major, z fastest (grid.rs:45-48)
        step = (e0 - s0) / fr                              # :101
        delta = (f32(0.7) * (e0 - s0)) / fr                # :257
        offs = [np.array([step[0] if i & 4 else 0, step[1] if i & 2 else 0, step[2] if i & 1 else 0], dtype=f32)
                for i in range(8)]
        EDGES = [(0, 4), (1, 5), (2, 6), (3, 7), (0, 2), (1, 3), (4, 6), (5, 7), (0, 1), (2, 3), (4, 5), (6, 7)]
        points = np.full((R, R, R), 0xFFFFFFFF, dtype=np.uint32)
        verts = []
        one = f32(1.0)
        for x in range(R):
            for y in range(R):
                for z in range(R):
                    d = [dists[x, y, z], dists[x, y, z + 1], dists[x, y + 1, z], dists[x, y + 1, z + 1],
                         dists[x + 1, y, z], dists[x + 1, y, z + 1], dists[x + 1, y + 1, z], dists[x + 1, y + 1, z + 1]]
                    sp = [bool(sign_positive(v)) for v in d]
                    if all(s == sp[0] for s in sp):
                        continue
                    p0 = s0 + np.array([x, y, z], dtype=f32) * step
                    count, tot = 0, np.zeros(3, dtype=f32)
                    for fr_, to in EDGES:
                        if sp[fr_] == sp[to]:
                            continue
                        if d[fr_] < 0:
                            d_from, d_to = d[fr_], d[to]
                        else:
                            d_from, d_to = -d[fr_], -d[to]
                        if d_to == d_from:
                            w = f32(0.5)
                        else:
                            dl = f32(d_to - d_from)
                            w = f32(f32(d_from + dl) / dl)
                        assert w >= 0 and w <= 1
                        a = (p0 + offs[fr_]).astype(f32); b = (p0 + offs[to]).astype(f32)
                        pt = (a * f32(one - w) + b * w).astype(f32)
                        tot = (tot + pt).astype(f32); count += 1
                    p = (f32(0.0) + tot / f32(count)).astype(f32)
                    q = []
                    for c in range(3):
                        for sg in (1, -1):
                            u = np.zeros(3, dtype=f32); u[c] = 1
                            q.append((p + u * f32(sg * delta[c])).astype(f32))
                    dd = de_fn(np.stack([p] + q))
                    nv = np.array([dd[1] - dd[2], dd[3] - dd[4], dd[5] - dd[6]], dtype=f32)
                    mag = np.sqrt(f32(f32(nv[0] * nv[0] + nv[1] * nv[1]) + nv[2] * nv[2]))
                    nv = (nv * f32(one / mag)).astype(f32)
                    points[x, y, z] = len(verts)
                    verts.append(np.concatenate([p, nv, [dd[0]]]).astype(f32))
        idx = []
        for x in range(R):
            for y in range(R):
                for z in range(R):
                    dv = dists[x, y, z]
                    base = bool(sign_positive(dv)); neg = dv < 0
                    if y > 0 and z > 0 and base != bool(sign_positive(dists[x + 1, y, z])):
                        v0, v1, v2, v3 = points[x, y - 1, z - 1], points[x, y - 1, z], points[x, y, z - 1], points[x, y, z]
                        idx += [v0, v2, v1, v1, v2, v3] if neg else [v0, v1, v2, v1, v3, v2]
                    if x > 0 and z > 0 and base != bool(sign_positive(dists[x, y + 1, z])):
                        v0, v1, v2, v3 = points[x - 1, y, z - 1], points[x - 1, y, z], points[x, y, z - 1], points[x, y, z]
                        idx += [v0, v1, v2, v1, v3, v2] if neg else [v0, v2, v1, v1, v2, v3]
                    if x > 0 and y > 0 and base != bool(sign_positive(dists[x, y, z + 1])):
                        v0, v1, v2, v3 = points[x - 1, y - 1, z], points[x - 1, y, z], points[x, y - 1, z], points[x, y, z]
                        idx += [v0, v2, v1, v1, v2, v3] if neg else [v0, v1, v2, v1, v3, v2]
        V = np.stack(verts) if verts else np.zeros((0, 7), dtype=f32)
        return V, np.array(idx, dtype=np.uint32), dists
