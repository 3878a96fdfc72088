// de_device.cuh -- device-side distance estimators (sm_100a).
//
// Implements the reference's Shape::min_distance_from for Mandelbulb<P>
// (/root/reference/src/shape/mandelbulb.rs:59-79, rotate :96-200) and Sphere
// (src/shape/sphere.rs:33-35) under two arithmetic policies:
//
//   MathExact -- every f32 op is a correctly-rounded IEEE op issued through
//                __fmul_rn/__fadd_rn/__fdiv_rn/__fsqrt_rn (never contracted to
//                FMA, like rustc), in the reference's evaluation order, and
//                ln() is glibc's logf algorithm.  Power-8 distances are
//                bit-identical to the reference's x86-64 CPU path.
//   MathFast  -- FMA-contracted, re-associated polynomial, MUFU rsqrt/rcp/lg2,
//                squared-radius bailout test, trig-free complex powers for the
//                generic-P path.  Tolerance mode.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace ctc {

struct ShapeDev {
    int32_t  kind;       // CTC_SHAPE_*
    uint32_t power;
    uint32_t max_iters;  // clamped to 2^32-1 on the host
    float    bailout;
    float    bail2;      // bailout * bailout (host, IEEE): the FAST path tests squared radii
    float    cx, cy, cz, radius;
};

// x86's default NaN (sign bit set): what every NaN the reference's DE can
// produce looks like, so f32::is_sign_positive() classifies it "inside".
// CUDA's canonical NaN is 0x7FFFFFFF and would classify "outside".
__device__ __forceinline__ float canonical_x86_nan(float v) {
    return (v != v) ? __int_as_float(0xFFC00000) : v;
}

// ---------------------------------------------------------------------------
// glibc 2.39 logf (sysdeps/ieee754/flt-32/e_logf.c, FMA build __logf_fma):
// 16-entry table + degree-3 polynomial in double.  Third-party algorithm, not
// part of /root/reference; Rust's f32::ln resolves to it on x86-64 Linux.
// ---------------------------------------------------------------------------
__device__ const double kLogfTab[32] = {
    0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2, 0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2,
    0x1.49539f0f010b0p+0, -0x1.01eae7f513a67p-2, 0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3,
    0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3, 0x1.25e227b0b8ea0p+0, -0x1.1aa2bc79c8100p-3,
    0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4, 0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4,
    0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5, 0x1.0000000000000p+0,  0x0.0p+0,
    0x1.e608cfd9a47acp-1,  0x1.aa5aa5df25984p-5, 0x1.ca4b31f026aa0p-1,  0x1.c5e53aa362eb4p-4,
    0x1.b2036576afce6p-1,  0x1.526e57720db08p-3, 0x1.9c2d163a1aa2dp-1,  0x1.bc2860d224770p-3,
    0x1.886e6037841edp-1,  0x1.1058bc8a07ee1p-2, 0x1.767dcf5534862p-1,  0x1.4043057b6ee09p-2,
};

__device__ __forceinline__ float logf_glibc(float x) {
    uint32_t ix = __float_as_uint(x);
    if (ix == 0x3f800000u) return 0.0f;
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
        if (ix * 2u == 0u) return __int_as_float(0xff800000);         // log(+-0) = -inf
        if (ix == 0x7f800000u) return x;                              // log(inf) = inf
        if ((ix & 0x80000000u) || ix * 2u >= 0xff000000u)             // negative or NaN
            return __int_as_float(0xFFC00000);
        ix = __float_as_uint(__fmul_rn(x, 0x1p23f));                  // subnormal: normalise
        ix -= 23u << 23;
    }
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (int)((tmp >> 19) & 15u);
    const int k = (int)tmp >> 23;
    const uint32_t iz = ix - (tmp & 0xff800000u);
    const double invc = kLogfTab[2 * i], logc = kLogfTab[2 * i + 1];
    const double z = (double)__uint_as_float(iz);
    const double r  = __fma_rn(z, invc, -1.0);
    const double y0 = __fma_rn((double)k, 0x1.62e42fefa39efp-1, logc);
    const double r2 = __dmul_rn(r, r);
    double y = __fma_rn(0x1.5575b0be00b6ap-2, r, -0x1.ffffef20a4123p-2);
    y = __fma_rn(-0x1.00ea348b88334p-2, r2, y);
    y = __fma_rn(y, r2, __dadd_rn(y0, r));
    return __double2float_rn(y);
}

// ---------------------------------------------------------------------------
// arithmetic policies
// ---------------------------------------------------------------------------
struct MathExact {
    static constexpr bool kExact = true;
    __device__ static __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    __device__ static __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    __device__ static __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    __device__ static __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    __device__ static __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
    __device__ static __forceinline__ float log(float a) { return logf_glibc(a); }
};

// f32::powi(n) with constant n, as LLVM expands it (binary square-and-multiply;
// SURVEY 8a, a7).  n is warp-uniform.
template <class M>
__device__ __forceinline__ float powi(float x, uint32_t n) {
    if (n == 0) return 1.0f;
    float res = 0.0f, cur = x;
    bool have = false;
    while (n) {
        if (n & 1u) { res = have ? M::mul(res, cur) : cur; have = true; }
        n >>= 1;
        if (n) cur = M::mul(cur, cur);
    }
    return res;
}

// ---------------------------------------------------------------------------
// EXACT: literal evaluation order of the reference
// ---------------------------------------------------------------------------

// rotate_inner_p8_scalar (mandelbulb.rs:148-200)
__device__ __forceinline__ void rotate_p8_exact(float x, float y, float z, float& ox, float& oy, float& oz) {
    using M = MathExact;
    const float x2 = M::mul(x, x), x4 = M::mul(x2, x2), x6 = M::mul(x4, x2), x8 = M::mul(x4, x4);
    const float y2 = M::mul(y, y), y4 = M::mul(y2, y2), y6 = M::mul(y4, y2), y8 = M::mul(y4, y4);
    const float z2 = M::mul(z, z), z4 = M::mul(z2, z2), z6 = M::mul(z4, z2), z8 = M::mul(z4, z4);
    const float w2 = M::add(x2, y2), w4 = M::mul(w2, w2), w6 = M::mul(w2, w4), w8 = M::mul(w4, w4);

    float t = M::sub(z8, M::mul(M::mul(28.0f, z6), w2));
    t = M::add(t, M::mul(M::mul(70.0f, z4), w4));
    t = M::sub(t, M::mul(M::mul(28.0f, z2), w6));
    const float a = M::add(1.0f, M::div(t, w8));

    float px = M::sub(x8, M::mul(M::mul(28.0f, x6), y2));
    px = M::add(px, M::mul(M::mul(70.0f, x4), y4));
    px = M::sub(px, M::mul(M::mul(28.0f, x2), y6));
    px = M::sub(px, y8);
    ox = M::mul(a, px);

    float py = M::sub(x6, M::mul(M::mul(7.0f, x4), y2));
    py = M::add(py, M::mul(M::mul(7.0f, x2), y4));
    py = M::sub(py, y6);
    oy = M::mul(M::mul(M::mul(M::mul(8.0f, a), x), y), py);

    const float pz = M::add(M::sub(z4, M::mul(M::mul(6.0f, z2), w2)), w4);
    oz = M::mul(M::mul(M::mul(M::mul(8.0f, z), M::sqrt(w2)), M::sub(z2, w2)), pz);
}

// rotate_inner_px_generic::<P> (mandelbulb.rs:128-146).  CUDA's accurate
// acosf/atan2f/sinf/cosf stand in for glibc's (<= 2 ulp apart): tolerance path.
__device__ __noinline__ void rotate_generic_exact(uint32_t P, float x, float y, float z, float r,
                                                  float& ox, float& oy, float& oz) {
    using M = MathExact;
    float theta = acosf(M::div(z, r));
    float phi = atan2f(y, x);
    const float new_radius = powi<M>(r, P);
    theta = M::mul(theta, (float)P);
    phi = M::mul(phi, (float)P);
    float st, ct, sp, cp;
    sincosf(theta, &st, &ct);
    sincosf(phi, &sp, &cp);
    ox = M::mul(M::mul(st, cp), new_radius);
    oy = M::mul(M::mul(sp, st), new_radius);
    oz = M::mul(ct, new_radius);
}

// rotate_on_z_axis::<P> (mandelbulb.rs:114-126), #[cold]
__device__ __noinline__ float rotate_on_z_axis_exact(uint32_t P, float z, float r) {
    using M = MathExact;
    float theta = acosf(M::div(z, r));      // 0/0 at the origin -> NaN, as in the reference
    const float new_radius = powi<M>(r, P);
    theta = M::mul(theta, (float)P);
    return M::mul(new_radius, cosf(theta));
}

// Mandelbulb::<P>::min_distance_from (mandelbulb.rs:59-79)
template <bool kP8>
__device__ __forceinline__ float mandelbulb_de_exact(const ShapeDev& s, float px, float py, float pz,
                                                     uint32_t* iters_out = nullptr) {
    using M = MathExact;
    const uint32_t P = kP8 ? 8u : s.power;
    const float fP = (float)P;
    float zx = px, zy = py, zz = pz;
    float dr = 1.0f, r = 0.0f;
    uint32_t it = 0;
    for (; it < s.max_iters; ++it) {
        // Vec3::magnitude (mandelbulb.rs:411-418): sqrt((x*x + y*y) + z*z)
        r = M::sqrt(M::add(M::add(M::mul(zx, zx), M::mul(zy, zy)), M::mul(zz, zz)));
        if (r > s.bailout) break;
        // dr = r.powi(P-1) * P * dr + 1.0
        dr = M::add(M::mul(M::mul(powi<M>(r, P - 1u), fP), dr), 1.0f);
        float nx, ny, nz;
        if (zx == 0.0f && zy == 0.0f) {            // is_on_z_axis: x, y are +-0.0
            nx = 0.0f; ny = 0.0f;
            nz = rotate_on_z_axis_exact(P, zz, r);
        } else if (kP8) {
            rotate_p8_exact(zx, zy, zz, nx, ny, nz);
        } else {
            rotate_generic_exact(P, zx, zy, zz, r, nx, ny, nz);
        }
        zx = M::add(nx, px); zy = M::add(ny, py); zz = M::add(nz, pz);
    }
    if (iters_out) *iters_out = it;
    const float ln_r = M::mul(M::log(r), r);
    return canonical_x86_nan(M::div(M::mul(0.5f, ln_r), dr));
}

// EXACT power-8 DE along a lattice COLUMN (K1 walks z with (px, py) fixed): every sub-expression of
// the FIRST iteration that depends on x and y only is computed once per column.  Pure common-
// subexpression hoisting: each value is produced by the same IEEE operation on the same operands as
// in rotate_p8_exact / mandelbulb_de_exact, so the result is bit-identical.
struct ColumnExactP8 {
    float w2, w4, w6, w8, sqrt_w2, poly_x, poly_y;
};

__device__ __forceinline__ ColumnExactP8 column_exact_p8(float x, float y) {
    using M = MathExact;
    ColumnExactP8 c;
    const float x2 = M::mul(x, x), x4 = M::mul(x2, x2), x6 = M::mul(x4, x2), x8 = M::mul(x4, x4);
    const float y2 = M::mul(y, y), y4 = M::mul(y2, y2), y6 = M::mul(y4, y2), y8 = M::mul(y4, y4);
    c.w2 = M::add(x2, y2); c.w4 = M::mul(c.w2, c.w2); c.w6 = M::mul(c.w2, c.w4); c.w8 = M::mul(c.w4, c.w4);
    c.sqrt_w2 = M::sqrt(c.w2);
    float px = M::sub(x8, M::mul(M::mul(28.0f, x6), y2));
    px = M::add(px, M::mul(M::mul(70.0f, x4), y4));
    px = M::sub(px, M::mul(M::mul(28.0f, x2), y6));
    c.poly_x = M::sub(px, y8);
    float py = M::sub(x6, M::mul(M::mul(7.0f, x4), y2));
    py = M::add(py, M::mul(M::mul(7.0f, x2), y4));
    c.poly_y = M::sub(py, y6);
    return c;
}

// Caller guarantees (px, py) != (+-0, +-0) (the z-axis special case is handled per warp in K1).
__device__ __forceinline__ float mandelbulb_de_exact_p8_column(const ShapeDev& s, float px, float py, float pz,
                                                               const ColumnExactP8& c) {
    using M = MathExact;
    float dr = 1.0f;
    const float z2 = M::mul(pz, pz);
    float r = M::sqrt(M::add(c.w2, z2));                     // sqrt((x*x + y*y) + z*z)
    if (!(r > s.bailout)) {
        // dr = r.powi(7) * 8 * 1.0 + 1.0   (the multiplication by dr = 1.0 is exact)
        dr = M::add(M::mul(powi<M>(r, 7u), 8.0f), 1.0f);
        const float z4 = M::mul(z2, z2), z6 = M::mul(z4, z2), z8 = M::mul(z4, z4);
        float t = M::sub(z8, M::mul(M::mul(28.0f, z6), c.w2));
        t = M::add(t, M::mul(M::mul(70.0f, z4), c.w4));
        t = M::sub(t, M::mul(M::mul(28.0f, z2), c.w6));
        const float a = M::add(1.0f, M::div(t, c.w8));
        const float ox = M::mul(a, c.poly_x);
        const float oy = M::mul(M::mul(M::mul(M::mul(8.0f, a), px), py), c.poly_y);
        const float pz4 = M::add(M::sub(z4, M::mul(M::mul(6.0f, z2), c.w2)), c.w4);
        const float oz = M::mul(M::mul(M::mul(M::mul(8.0f, pz), c.sqrt_w2), M::sub(z2, c.w2)), pz4);
        float zx = M::add(ox, px), zy = M::add(oy, py), zz = M::add(oz, pz);
        for (uint32_t it = 1; it < s.max_iters; ++it) {
            r = M::sqrt(M::add(M::add(M::mul(zx, zx), M::mul(zy, zy)), M::mul(zz, zz)));
            if (r > s.bailout) break;
            dr = M::add(M::mul(M::mul(powi<M>(r, 7u), 8.0f), dr), 1.0f);
            float nx, ny, nz;
            if (zx == 0.0f && zy == 0.0f) {
                nx = 0.0f; ny = 0.0f;
                nz = rotate_on_z_axis_exact(8u, zz, r);
            } else {
                rotate_p8_exact(zx, zy, zz, nx, ny, nz);
            }
            zx = M::add(nx, px); zy = M::add(ny, py); zz = M::add(nz, pz);
        }
    }
    const float ln_r = M::mul(M::log(r), r);
    return canonical_x86_nan(M::div(M::mul(0.5f, ln_r), dr));
}

// ---------------------------------------------------------------------------
// FAST: same recurrence, FMA-contracted and re-associated
// ---------------------------------------------------------------------------
__device__ __forceinline__ float fast_rcp(float a) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
__device__ __forceinline__ float fast_sqrt(float a) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
__device__ __forceinline__ float fast_rsqrt(float a) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
__device__ __forceinline__ float fast_lg2(float a) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }

// (a + i b)^n by binary exponentiation, n >= 1 warp-uniform
__device__ __forceinline__ void cpow(float a, float b, uint32_t n, float& re, float& im) {
    float cr = a, ci = b;       // running square
    float rr = 1.0f, ri = 0.0f; // result
    bool have = false;
    while (n) {
        if (n & 1u) {
            if (have) { const float t = rr * cr - ri * ci; ri = rr * ci + ri * cr; rr = t; }
            else { rr = cr; ri = ci; have = true; }
        }
        n >>= 1;
        if (n) { const float t = cr * cr - ci * ci; ci = 2.0f * cr * ci; cr = t; }
    }
    re = rr; im = ri;
}

// One power-8 step of the FAST path.  Same map as rotate_inner_p8_scalar:
//     X = a (x8 - 28 x6 y2 + 70 x4 y4 - 28 x2 y6 - y8)     a = 1 + (z8 - 28 z6 w2 + 70 z4 w4 - 28 z2 w6) / w8
//     Y = 8 a x y (x6 - 7 x4 y2 + 7 x2 y4 - y6)            w2 = x2 + y2
//     Z = 8 z w (z2 - w2)(z4 - 6 z2 w2 + w4)
// read as complex 8th powers, which three squarings evaluate:
//     a w8 = Re (z + i w)^8 =: A,   Z = Im (z + i w)^8,
//     X = A (cos 8phi - 2 v^8),  Y = A sin 8phi   with (u, v) = (x, y) / w = e^{i phi}
// (the "- 2 v^8" is the reference's "- y8" where the real part of (x + i y)^8 has "+ y8"; it is
// reproduced, not fixed).  Working on the unit vector (u, v) removes the division by w^8, which
// underflows near the z axis and injects inf/NaN into the reference's own arithmetic; w2 is clamped
// before the rsqrt, so the step is finite for every finite input.
// 28 FMA-pipe instructions + 1 MUFU per step (+ 4 and 1 MUFU for dr) against 75 algorithmic flops.
__device__ __forceinline__ void p8_azimuth(float x, float y, float w2, float& iw, float& c8, float& s8h) {
    iw = fast_rsqrt(fmaxf(w2, 1e-36f));
    const float u = x * iw, v = y * iw;
    const float v2 = v * v;
    const float c2 = fmaf(u, u, -v2);                       // cos 2phi
    const float s2 = 2.0f * (u * v);                        // sin 2phi
    const float c4 = fmaf(c2, c2, -(s2 * s2));              // cos 4phi
    const float q4 = 2.0f * (s2 * c2);                      // sin 4phi
    const float v4 = v2 * v2;
    c8 = fmaf(-2.0f, v4 * v4, fmaf(c4, c4, -(q4 * q4)));    // cos 8phi - 2 v^8
    s8h = c4 * q4;                                          // sin 8phi / 2
}

__device__ __forceinline__ void p8_elevation(float z, float z2, float w, float w2, float& A, float& Zh) {
    const float s2 = 2.0f * (z * w);
    const float c2 = z2 - w2;
    const float c4 = fmaf(c2, c2, -(s2 * s2));
    const float q4 = 2.0f * (s2 * c2);
    A = fmaf(c4, c4, -(q4 * q4));                           // Re (z + i w)^8
    Zh = c4 * q4;                                           // Im (z + i w)^8 / 2
}

__device__ __forceinline__ void rotate_p8_fast(float x, float y, float z, float z2, float w2,
                                               float px, float py, float pz, float& ox, float& oy, float& oz) {
    float iw, c8, s8h, A, Zh;
    p8_azimuth(x, y, w2, iw, c8, s8h);
    p8_elevation(z, z2, w2 * iw, w2, A, Zh);
    ox = fmaf(A, c8, px);
    oy = fmaf(2.0f * A, s8h, py);
    oz = fmaf(2.0f, Zh, pz);
}

// FAST path for a sample ON the z axis (px == py == 0 exactly): the orbit never leaves the axis
// (rotate_on_z_axis, mandelbulb.rs:114-126): z' = |z|^P cos(P theta) + pz with theta in {0, pi},
// i.e. z' = z^P + pz for even P and the same for odd P (cos(P pi) = -1 flips the sign back).
// The origin is 0/0 in the reference -> NaN.  Cold: one lattice column per dense grid at most.
__device__ __noinline__ float mandelbulb_de_fast_on_axis(uint32_t P, uint32_t max_iters, float bailout, float pz) {
    float zz = pz, dr = 1.0f, r = 0.0f;
    for (uint32_t it = 0; it < max_iters; ++it) {
        r = fabsf(zz);
        if (r > bailout) break;
        if (r == 0.0f) return __int_as_float(0xFFC00000);
        float rp1 = 1.0f;
        { float cur = r; uint32_t n = P - 1u; while (n) { if (n & 1u) rp1 *= cur; n >>= 1; if (n) cur *= cur; } }
        dr = fmaf((float)P * rp1, dr, 1.0f);
        const float rp = rp1 * r;
        zz = ((zz < 0.0f && (P & 1u)) ? -rp : rp) + pz;
    }
    return canonical_x86_nan(0.5f * __logf(r) * r / dr);
}

// kCheckAxis = false: the caller guarantees (px, py) != (0, 0) (K1 tests it once per warp).
template <bool kP8, bool kCheckAxis = true>
__device__ __forceinline__ float mandelbulb_de_fast(const ShapeDev& s, float px, float py, float pz) {
    const uint32_t P = kP8 ? 8u : s.power;
    if (kCheckAxis && px == 0.0f && py == 0.0f) return mandelbulb_de_fast_on_axis(P, s.max_iters, s.bailout, pz);
    const float bail2 = s.bail2;
    float zx = px, zy = py, zz = pz;
    float dr = 1.0f, r2;
    uint32_t left = s.max_iters;                 // >= 1 (checked on the host, mandelbulb.rs:20)
    do {
        const float z2 = zz * zz;
        const float w2 = fmaf(zx, zx, zy * zy);
        r2 = w2 + z2;
        if (r2 > bail2) break;                   // r > bailout, on squares
        if (kP8) {
            // dr = 8 r^7 dr + 1
            const float r6 = r2 * r2 * r2;
            dr = fmaf(8.0f * (r6 * fast_sqrt(r2)), dr, 1.0f);
            rotate_p8_fast(zx, zy, zz, z2, w2, px, py, pz, zx, zy, zz);
        } else {
            const float r = fast_sqrt(r2);
            // generic P without trig: (z + i w)^P = r^P (cos P.theta + i sin P.theta),
            // ((x + i y)/w)^P = cos P.phi + i sin P.phi
            float rp1 = 1.0f;                       // r^(P-1)
            { float cur = r; uint32_t n = P - 1u; while (n) { if (n & 1u) rp1 *= cur; n >>= 1; if (n) cur *= cur; } }
            dr = fmaf((float)P * rp1, dr, 1.0f);
            const float iw = fast_rsqrt(fmaxf(w2, 1e-36f));
            const float w = w2 * iw;
            float ct, st, cp, sp;
            cpow(zz, w, P, ct, st);                 // r^P cos(P theta), r^P sin(P theta)
            cpow(zx * iw, zy * iw, P, cp, sp);      // cos(P phi), sin(P phi)
            zx = fmaf(st, cp, px); zy = fmaf(st, sp, py); zz = ct + pz;
        }
    } while (--left);
    // 0.5 * ln(r) * r / dr with ln(r) = 0.5 * ln2 * lg2(r2)
    return (0.25f * 0.69314718056f) * fast_lg2(r2) * fast_sqrt(r2) * fast_rcp(dr);
}

// FAST power-8 DE for a sample of a lattice COLUMN: K1 walks 32 z-samples with (px, py) fixed, so the
// first iteration's azimuth factors (c8, s8h), w and w2 are computed once per column and only the
// elevation part of the first step is per sample.  Caller guarantees (px, py) != (0, 0).
__device__ __forceinline__ float mandelbulb_de_fast_p8_column(const ShapeDev& s, float px, float py, float pz,
                                                              float w2c, float wc, float c8, float s8h) {
    const float bail2 = s.bail2;
    float dr = 1.0f;
    float r2 = fmaf(pz, pz, w2c);
    if (!(r2 > bail2)) {
        const float z2 = pz * pz;
        const float r6 = r2 * r2 * r2;
        dr = fmaf(8.0f, r6 * fast_sqrt(r2), 1.0f);           // 8 r^7 * 1 + 1
        float A, Zh;
        p8_elevation(pz, z2, wc, w2c, A, Zh);
        float zx = fmaf(A, c8, px), zy = fmaf(2.0f * A, s8h, py), zz = fmaf(2.0f, Zh, pz);
        for (uint32_t left = s.max_iters - 1u; left; --left) {
            const float zz2 = zz * zz;
            const float w2 = fmaf(zx, zx, zy * zy);
            r2 = w2 + zz2;
            if (r2 > bail2) break;
            const float q6 = r2 * r2 * r2;
            dr = fmaf(8.0f * (q6 * fast_sqrt(r2)), dr, 1.0f);
            rotate_p8_fast(zx, zy, zz, zz2, w2, px, py, pz, zx, zy, zz);
        }
    }
    return (0.25f * 0.69314718056f) * fast_lg2(r2) * fast_sqrt(r2) * fast_rcp(dr);
}

// Sphere::min_distance_from (sphere.rs:33-35); cgmath magnitude = sqrt((x*x+y*y)+z*z)
__device__ __forceinline__ float sphere_de(const ShapeDev& s, float px, float py, float pz) {
    using M = MathExact;
    const float dx = M::sub(s.cx, px), dy = M::sub(s.cy, py), dz = M::sub(s.cz, pz);
    return M::sub(M::sqrt(M::add(M::add(M::mul(dx, dx), M::mul(dy, dy)), M::mul(dz, dz))), s.radius);
}

// Shape dispatch.  kVariant: 0 = Mandelbulb P=8, 1 = Mandelbulb generic P, 2 = Sphere.
enum : int { kVarP8 = 0, kVarGeneric = 1, kVarSphere = 2 };

template <bool kFast, int kVariant, bool kCheckAxis = true>
__device__ __forceinline__ float shape_de(const ShapeDev& s, float px, float py, float pz) {
    if (kVariant == kVarSphere) return sphere_de(s, px, py, pz);
    if (kFast) return mandelbulb_de_fast<kVariant == kVarP8, kCheckAxis>(s, px, py, pz);
    return mandelbulb_de_exact<kVariant == kVarP8>(s, px, py, pz);
}

}  // namespace ctc
