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
    float    kappa;      // FAST: sign-trust band |r^2 - 1| <= kappa * 2^lemax (see fast_suspect_*)
    int32_t  le_trust;   // FAST: smallest lemax with kappa * 2^lemax >= 4 (host, amp_trust_threshold): an escaped
                         // sample's orbit is trusted below it -- one integer compare instead of a multiply
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
                                                     uint32_t* iters_out = nullptr, float* r_out = nullptr) {
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
    if (r_out) *r_out = r;
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
// FAST: same recurrence, FMA-contracted and re-associated.
//
// The power-8 step is written ONCE over a lane-vector type V.  V = float evaluates one sample per
// thread; V = float2 evaluates two samples per thread in sm_100's packed FP32 instructions
// (FFMA2 / FMUL2 / FADD2: one issue slot, two IEEE-rounded results).  K1 and E3 are bound by issue
// slots in scalar form (ncu: issue 94 %, FMA pipe 67 %); in packed form the FMA pipe is the bound.
// Every operation is an explicit round-to-nearest intrinsic, so both instantiations produce the same
// bits for the same sample (the scalar form is what the tail paths, K3 and the probes run).
// ---------------------------------------------------------------------------
__device__ __forceinline__ float fast_rcp(float a) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
__device__ __forceinline__ float fast_sqrt(float a) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
__device__ __forceinline__ float fast_rsqrt(float a) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
__device__ __forceinline__ float fast_lg2(float a) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }

// unroll factor of the packed pair loop (A/B on the benched volume, scripts/gpu_ab.py)
#ifndef CTC_PAIR_UNROLL
#define CTC_PAIR_UNROLL 2
#endif
#define CTC_PRAGMA_(x) _Pragma(#x)
#define CTC_PRAGMA_UNROLL(n) CTC_PRAGMA_(unroll n)

template <class V> struct Lanes;

template <> struct Lanes<float> {
    __device__ static __forceinline__ float bc(float a) { return a; }
    __device__ static __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    __device__ static __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    __device__ static __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    __device__ static __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
    __device__ static __forceinline__ float fms(float a, float b, float c) { return __fmaf_rn(a, b, -c); }   // a*b - c
    __device__ static __forceinline__ float rsqrt(float a) { return fast_rsqrt(a); }
    __device__ static __forceinline__ float sqrt(float a) { return fast_sqrt(a); }
    __device__ static __forceinline__ float vmin(float a, float b) { return fminf(a, b); }
    __device__ static __forceinline__ float vmax(float a, float b) { return fmaxf(a, b); }
};

template <> struct Lanes<float2> {
    __device__ static __forceinline__ float2 bc(float a) { return make_float2(a, a); }
    __device__ static __forceinline__ float2 mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
    __device__ static __forceinline__ float2 add(float2 a, float2 b) { return __fadd2_rn(a, b); }
    __device__ static __forceinline__ float2 sub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
    __device__ static __forceinline__ float2 fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
    __device__ static __forceinline__ float2 fms(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, make_float2(-c.x, -c.y)); }
    __device__ static __forceinline__ float2 rsqrt(float2 a) { return make_float2(fast_rsqrt(a.x), fast_rsqrt(a.y)); }
    __device__ static __forceinline__ float2 sqrt(float2 a) { return make_float2(fast_sqrt(a.x), fast_sqrt(a.y)); }
    __device__ static __forceinline__ float2 vmin(float2 a, float2 b) { return make_float2(fminf(a.x, b.x), fminf(a.y, b.y)); }
    __device__ static __forceinline__ float2 vmax(float2 a, float2 b) { return make_float2(fmaxf(a.x, b.x), fmaxf(a.y, b.y)); }
};

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
// underflows near the z axis and injects inf/NaN into the reference's own arithmetic.  The squarings
// carry NEGATED half-angle sines (m4n = -2 sin, q4n = -sin 4x) so that no operand needs a separate
// negation or doubling: 15 + 10 + 4 FMA-pipe operations per step, + 4 for the magnitude and 5 for dr
// = 38 against 75 algorithmic flops.
template <class V>
__device__ __forceinline__ void p8_azimuth(V x, V y, V iw, V& c8, V& s8hn) {
    using L = Lanes<V>;
    const V u = L::mul(x, iw), v = L::mul(y, iw);
    const V m = L::mul(u, v);                               // sin 2phi / 2
    const V v2 = L::mul(v, v);
    const V c2 = L::fms(u, u, v2);                          // cos 2phi
    const V cc = L::mul(c2, c2);
    const V m4n = L::mul(m, L::bc(-4.0f));                  // -2 sin 2phi
    const V c4 = L::fma(m4n, m, cc);                        // cos 4phi = c2^2 - s2^2
    const V q4n = L::mul(m4n, c2);                          // -sin 4phi
    const V v4 = L::mul(v2, v2);
    const V v8 = L::mul(v4, v4);
    const V t = L::mul(q4n, q4n);
    const V c8a = L::fms(c4, c4, t);                        // cos 8phi
    c8 = L::fma(v8, L::bc(-2.0f), c8a);                     // cos 8phi - 2 v^8
    s8hn = L::mul(c4, q4n);                                 // -sin 8phi / 2
}

template <class V>
__device__ __forceinline__ void p8_elevation(V z, V z2, V w, V w2, V& A, V& Zhn) {
    using L = Lanes<V>;
    const V m = L::mul(z, w);
    const V c2 = L::sub(z2, w2);
    const V cc = L::mul(c2, c2);
    const V m4n = L::mul(m, L::bc(-4.0f));
    const V c4 = L::fma(m4n, m, cc);
    const V q4n = L::mul(m4n, c2);
    const V t = L::mul(q4n, q4n);
    A = L::fms(c4, c4, t);                                  // Re (z + i w)^8
    Zhn = L::mul(c4, q4n);                                  // -Im (z + i w)^8 / 2
}

// dr = 8 r^7 dr + 1 from r^2; g = 8 r^7 (the step's radial stretch) is kept for the error bookkeeping
template <class V>
__device__ __forceinline__ V p8_dr(V r2, V dr, V& g) {
    using L = Lanes<V>;
    const V r4 = L::mul(r2, r2);
    g = L::mul(L::mul(L::mul(r4, r2), L::sqrt(r2)), L::bc(8.0f));
    return L::fma(g, dr, L::bc(1.0f));
}

// z <- z^8 + p (the rotate and the "+ p" of mandelbulb.rs:74 in one FMA each).  A = Re (z + i w)^8
// (the new distance from the z axis before "+ p") and iw = 1 / w go to the error bookkeeping.
template <class V>
__device__ __forceinline__ void p8_step(V& zx, V& zy, V& zz, V z2, V w2, V px, V py, V pz, V& A, V& iw) {
    using L = Lanes<V>;
    iw = L::rsqrt(w2);
    V c8, s8hn, Zhn;
    p8_azimuth<V>(zx, zy, iw, c8, s8hn);
    p8_elevation<V>(zz, z2, L::mul(w2, iw), w2, A, Zhn);
    zx = L::fma(A, c8, px);
    zy = L::fma(L::mul(A, L::bc(-2.0f)), s8hn, py);
    zz = L::fma(L::bc(-2.0f), Zhn, pz);
}

// 0.5 * ln(r) * r / dr with ln(r) = 0.5 * ln2 * lg2(r2)
__device__ __forceinline__ float de_fast_epilogue(float r2, float dr) {
    return (0.25f * 0.69314718056f) * fast_lg2(r2) * fast_sqrt(r2) * fast_rcp(dr);
}

// the same for both halves of a pair (same association, so the same bits as the scalar form)
__device__ __forceinline__ float2 de_fast_epilogue2(float2 r2, float2 dr) {
    const float2 lg = make_float2(fast_lg2(r2.x), fast_lg2(r2.y));
    const float2 sq = make_float2(fast_sqrt(r2.x), fast_sqrt(r2.y));
    const float2 rc = make_float2(fast_rcp(dr.x), fast_rcp(dr.y));
    const float c = 0.25f * 0.69314718056f;
    return __fmul2_rn(__fmul2_rn(__fmul2_rn(make_float2(c, c), lg), sq), rc);
}

// ---------------------------------------------------------------------------
// Can the SIGN of a fast evaluation be trusted?  (DESIGN.md "sign-exact fast mode")
//
// The mesher's topology is a function of f32::is_sign_positive() of every sample, and the sign of
// 0.5 ln(r) r / dr is decided by ONE comparison: a sample that leaves through `r > bailout` is
// positive; one that runs all max_iters iterations is negative iff its last radius is < 1.  (A
// borderline bailout decision cannot flip the sign: with r ~ bailout the next radius is ~ bailout^8.)
// So the question is how far the fast path's last r^2 can be from the reference's.  To first order the
// orbit's accumulated rounding error obeys  e' = L e + eta  with L the largest stretch of the step:
//   * 8 r^7 radially and along the polar angle (what the reference's own dr carries), but
//   * 8 |A| / w along the azimuth (A = Re (z + i w)^8, the new distance from the z axis): the triplex
//     power is not conformal, near the poles it stretches azimuthal perturbations up to 8 r / w more.
//     Without this term four of the 1.1 G samples of the benched volume sit outside any useful band.
// The bound is kept in the LOG domain on raw float bits (2^23 per octave, piecewise-linear lg):
//     le' = max(le + max(lg 8 r^7, lg 8 |A| / w), 0),     Amp.lemax = max_k le_k
// -- seven ALU-pipe instructions per sample and iteration, none on the FMA pipe that bounds the kernels.
// Unlike a product of per-step factors it CONTRACTS where the map contracts, so a long orbit that
// settles on an attracting cycle stays trustworthy while a chaotic one does not.
// A sample is SUSPECT, and re-evaluated with the exact-order IEEE arithmetic, when
//   * it did not escape and |r^2 - 1| <= kappa 2^lemax           (sign of ln r not certain), or
//   * it escaped but kappa 2^lemax is O(1)                        (the orbit itself is not certain), or
//   * some iterate came within 2^-13 of the z axis               (the reference's w^8 underflows there, its
//                                                                 arithmetic turns inf/NaN, and the fast path's
//                                                                 1/w is inf on the axis itself), or
//   * anything is NaN (all tests are written so that NaN is suspect).
// kappa is calibrated on the GPU (ctc_fast_sign_probe, profiles/sign_probe_r2.md).
// kBand = false keeps only the axis/NaN rules (E3 and the generic powers: values, not signs).
// ---------------------------------------------------------------------------
constexpr int kFloatBias = 0x3f800000;
constexpr int kAxisIwBits = 0x46000000;       // bits of 8192.0f: 1/w >= 2^13

struct Amp { int le, lemax, iwmax; };

__device__ __forceinline__ Amp amp_init() { return Amp{0, 0, 0}; }

// one step: g = 8 r^7 (float), A, iw as above
template <bool kBand>
__device__ __forceinline__ void amp_step(Amp& a, float g, float A, float iw) {
    const unsigned iwb = __float_as_uint(iw);
    a.iwmax = max(a.iwmax, (int)iwb);
    if (kBand) {
        // lg(8 |A| / w) + bias = bits|A| + bits(1/w) + 3 octaves - bias        (u32: cannot overflow for finite A)
        const unsigned xa = (__float_as_uint(A) & 0x7fffffffu) + iwb - (unsigned)(kFloatBias - (3 << 23));
        const unsigned x = max(xa, __float_as_uint(g));
        a.le = max((int)((unsigned)a.le + x - (unsigned)kFloatBias), 0);
        a.lemax = max(a.lemax, a.le);
    }
}

// 2^lemax as a float factor (same piecewise-linear exponential; clamped far below overflow)
__device__ __forceinline__ float amp_factor(const Amp& a) { return __int_as_float(kFloatBias + min(a.lemax, 0x20000000)); }

// (kappa * amp_factor(lemax) is monotone in lemax, so "kappa 2^lemax >= 4" is lemax >= le_trust)
template <bool kBand>
__device__ __forceinline__ bool fast_suspect_escaped(const ShapeDev& s, const Amp& a) {
    return a.iwmax >= kAxisIwBits || (kBand && a.lemax >= s.le_trust);
}
template <bool kBand>
__device__ __forceinline__ bool fast_suspect_inside(const ShapeDev& s, float r2, const Amp& a) {
    if (!kBand) return a.iwmax >= kAxisIwBits || !(r2 == r2);
    return a.iwmax >= kAxisIwBits || !(fabsf(r2 - 1.0f) > s.kappa * amp_factor(a));
}

struct FastInfo { float r2, dr, amp; uint32_t escaped, axis; };   // probe output (ctc_fast_sign_probe)

// FAST power-8 DE of one sample.
template <bool kBand>
__device__ __forceinline__ float mandelbulb_de_fast_p8(const ShapeDev& s, float px, float py, float pz, bool& suspect,
                                                       FastInfo* info = nullptr) {
    using L = Lanes<float>;
    float zx = px, zy = py, zz = pz, dr = 1.0f, r2;
    Amp amp = amp_init();
    uint32_t left = s.max_iters;                 // >= 1 (checked on the host, mandelbulb.rs:20)
    for (;;) {
        const float z2 = L::mul(zz, zz);
        const float w2 = L::fma(zx, zx, L::mul(zy, zy));
        r2 = L::add(w2, z2);
        if (r2 > s.bail2) {                      // r > bailout, on squares
            suspect = fast_suspect_escaped<kBand>(s, amp);
            if (info) *info = FastInfo{r2, dr, amp_factor(amp), 1u, amp.iwmax >= kAxisIwBits ? 1u : 0u};
            return de_fast_epilogue(r2, dr);
        }
        float g, A, iw;
        dr = p8_dr<float>(r2, dr, g);
        if (--left == 0u) break;                 // the reference's last rotate is dead work
        p8_step<float>(zx, zy, zz, z2, w2, px, py, pz, A, iw);
        amp_step<kBand>(amp, g, A, iw);
    }
    suspect = fast_suspect_inside<kBand>(s, r2, amp);
    if (info) *info = FastInfo{r2, dr, amp_factor(amp), 0u, amp.iwmax >= kAxisIwBits ? 1u : 0u};
    return de_fast_epilogue(r2, dr);
}

// Two samples per thread: the iteration loop shared by the column form (K1) and the point form (E3).
// On entry (zx, zy, zz) hold the current iterate of both halves, `left` >= 1 radius tests remain.
// The common exits are packed: both halves leaving at the same radius test (neighbouring samples of a
// lattice column almost always do) or both running all max_iters iterations take ONE epilogue in
// FMUL2 form.  Otherwise a half that escapes alone gets its distance at once and is parked on NaN (its
// `+ p.z` operand), so it neither escapes again nor disturbs the other half; (px, py) stay broadcast
// scalars when the halves share a lattice column.  (A variant that snapshots the escaping half and
// computes both distances in packed form after the loop measured 10 % slower: the extra live registers
// cost more moves than the duplicated epilogue costs instructions.)
#define CTC_PAIR_ESCAPE(H, BIT, AMP)                                                             \
    if (r2.H > bail2) {                                                                          \
        d.H = de_fast_epilogue(r2.H, dr.H);                                                      \
        if (fast_suspect_escaped<kBand>(s, AMP)) suspect |= BIT;                                 \
        done |= BIT;                                                                             \
        pz.H = __int_as_float(0x7fffffff);                                                       \
    }
// `return`s from the enclosing function when the pair is finished
#define CTC_PAIR_RADIUS_TEST()                                                                   \
    {                                                                                            \
        const bool ex_ = r2.x > bail2, ey_ = r2.y > bail2;     /* a parked half is NaN: false */ \
        if (ex_ | ey_) {                                                                         \
            if (ex_ & ey_) {                                                                     \
                d = de_fast_epilogue2(r2, dr);                                                   \
                suspect |= (fast_suspect_escaped<kBand>(s, ampa) ? 1u : 0u) |                    \
                           (fast_suspect_escaped<kBand>(s, ampb) ? 2u : 0u);                     \
                return;                                                                          \
            }                                                                                    \
            CTC_PAIR_ESCAPE(x, 1u, ampa)                                                         \
            CTC_PAIR_ESCAPE(y, 2u, ampb)                                                         \
            if (done == 3u) return;                                                              \
        }                                                                                        \
    }
// the halves that ran all max_iters iterations
#define CTC_PAIR_TAIL()                                                                          \
    if (done == 0u) {                                                                            \
        d = de_fast_epilogue2(r2, dr);                                                           \
        suspect |= (fast_suspect_inside<kBand>(s, r2.x, ampa) ? 1u : 0u) |                       \
                   (fast_suspect_inside<kBand>(s, r2.y, ampb) ? 2u : 0u);                        \
    } else if (!(done & 1u)) {                                                                   \
        d.x = de_fast_epilogue(r2.x, dr.x);                                                      \
        if (fast_suspect_inside<kBand>(s, r2.x, ampa)) suspect |= 1u;                            \
    } else {                                                                                     \
        d.y = de_fast_epilogue(r2.y, dr.y);                                                      \
        if (fast_suspect_inside<kBand>(s, r2.y, ampb)) suspect |= 2u;                            \
    }

template <bool kBand>
__device__ __forceinline__ void p8_pair_loop(const ShapeDev& s, float2 px, float2 py, float2 pz, float2 zx, float2 zy, float2 zz,
                                             float2 dr, Amp ampa, Amp ampb, uint32_t left, float2& d,
                                             uint32_t done, uint32_t& suspect) {
    using L = Lanes<float2>;
    const float bail2 = s.bail2;
    float2 r2;
    // (A/B on the benched volume: K1 -- the band-tracking instantiation -- is fastest unrolled 3x, E3 2x)
    constexpr int kUnroll = kBand ? CTC_PAIR_UNROLL + 1 : CTC_PAIR_UNROLL;
#pragma unroll kUnroll
    for (;;) {
        const float2 z2 = L::mul(zz, zz);
        const float2 w2 = L::fma(zx, zx, L::mul(zy, zy));
        r2 = L::add(w2, z2);
        CTC_PAIR_RADIUS_TEST()
        float2 g, A, iw;
        dr = p8_dr<float2>(r2, dr, g);
        if (--left == 0u) break;
        p8_step<float2>(zx, zy, zz, z2, w2, px, py, pz, A, iw);
        amp_step<kBand>(ampa, g.x, A.x, iw.x);
        amp_step<kBand>(ampb, g.y, A.y, iw.y);
    }
    CTC_PAIR_TAIL()
}

// FAST power-8 DE of two arbitrary points.  Bit k of `suspect`: the result of half k needs the exact path.
template <bool kBand>
__device__ __forceinline__ float2 mandelbulb_de_fast_p8_pair(const ShapeDev& s, float2 px, float2 py, float2 pz, uint32_t& suspect) {
    float2 d = make_float2(0.0f, 0.0f);
    suspect = 0u;
    p8_pair_loop<kBand>(s, px, py, pz, px, py, pz, make_float2(1.0f, 1.0f), amp_init(), amp_init(), s.max_iters, d, 0u, suspect);
    return d;
}

// FAST power-8 DE along a lattice COLUMN: K1 walks z with (px, py) fixed, so the first iteration's
// azimuth factors (which depend on x and y only) are computed once per column and only the elevation
// half of the first step is per sample.  Same operations on the same operands as the point form.
struct ColumnFastP8 { float w2, w, iw, c8, s8hn; };

__device__ __forceinline__ ColumnFastP8 column_fast_p8(float px, float py) {
    using L = Lanes<float>;
    ColumnFastP8 c;
    c.w2 = L::fma(px, px, L::mul(py, py));
    c.iw = L::rsqrt(c.w2);
    p8_azimuth<float>(px, py, c.iw, c.c8, c.s8hn);
    c.w = L::mul(c.w2, c.iw);
    return c;
}

template <bool kBand>
__device__ __forceinline__ void p8_column_pair_impl(const ShapeDev& s, float px_, float py_, float2 pz, const ColumnFastP8& c,
                                                    float2& d, uint32_t& suspect) {
    using L = Lanes<float2>;
    const float bail2 = s.bail2;
    const float2 px = L::bc(px_), py = L::bc(py_);
    float2 dr = L::bc(1.0f);
    uint32_t done = 0u, left = s.max_iters;
    Amp ampa = amp_init(), ampb = amp_init();
    const float2 z2 = L::mul(pz, pz);
    const float2 w2 = L::bc(c.w2);
    float2 r2 = L::add(w2, z2);
    CTC_PAIR_RADIUS_TEST()
    float2 g;
    dr = p8_dr<float2>(r2, dr, g);
    if (--left == 0u) {
        CTC_PAIR_TAIL()
        return;
    }
    float2 A, Zhn;
    p8_elevation<float2>(pz, z2, L::bc(c.w), w2, A, Zhn);
    const float2 zx = L::fma(A, L::bc(c.c8), px);
    const float2 zy = L::fma(L::mul(A, L::bc(-2.0f)), L::bc(c.s8hn), py);
    const float2 zz = L::fma(L::bc(-2.0f), Zhn, pz);
    amp_step<kBand>(ampa, g.x, A.x, c.iw);
    amp_step<kBand>(ampb, g.y, A.y, c.iw);
    p8_pair_loop<kBand>(s, px, py, pz, zx, zy, zz, dr, ampa, ampb, left, d, done, suspect);
}

template <bool kBand>
__device__ __forceinline__ float2 mandelbulb_de_fast_p8_column_pair(const ShapeDev& s, float px_, float py_, float2 pz,
                                                                    const ColumnFastP8& c, uint32_t& suspect) {
    float2 d = make_float2(0.0f, 0.0f);
    suspect = 0u;
    p8_column_pair_impl<kBand>(s, px_, py_, pz, c, d, suspect);
    return d;
}
// ---------------------------------------------------------------------------
// FAST generic powers that are powers of two (config 4's P = 2, 4, 16): the same trig-free complex powers as
// mandelbulb_de_fast_generic below, but with the K = log2 P squarings unrolled and two samples per thread in
// packed FP32, like the power-8 path:
//   (z + i w)^P = r^P (cos P.theta + i sin P.theta),   ((x + i y)/w)^P = cos P.phi + i sin P.phi
//   X = r^P sin P.theta cos P.phi + px,  Y = r^P sin P.theta sin P.phi + py,  Z = r^P cos P.theta + pz
// (mandelbulb.rs:128-146; no "- y8" quirk here: the generic step is the true map).  No sign band (exact mode
// itself is a tolerance path for these powers): only the z-axis / NaN rule marks a half suspect.
// ---------------------------------------------------------------------------
template <int K, class V>
__device__ __forceinline__ void pow2_step(V& zx, V& zy, V& zz, V z2, V w2, V px, V py, V pz, V& iw) {
    using L = Lanes<V>;
    iw = L::rsqrt(w2);
    const V w = L::mul(w2, iw);
    V er = L::sub(z2, w2), ei = L::mul(L::mul(zz, w), L::bc(2.0f));                 // (z + i w)^2
    const V u = L::mul(zx, iw), v = L::mul(zy, iw);
    V ar = L::fms(u, u, L::mul(v, v)), ai = L::mul(L::mul(u, v), L::bc(2.0f));      // ((x + i y) / w)^2
#pragma unroll
    for (int sq = 1; sq < K; ++sq) {
        const V t = L::fms(er, er, L::mul(ei, ei)); ei = L::mul(L::mul(er, ei), L::bc(2.0f)); er = t;
        const V a = L::fms(ar, ar, L::mul(ai, ai)); ai = L::mul(L::mul(ar, ai), L::bc(2.0f)); ar = a;
    }
    zx = L::fma(ei, ar, px); zy = L::fma(ei, ai, py); zz = L::add(er, pz);
}

// dr = P r^(P-1) dr + 1 with P = 2^K, from r^2
template <int K, class V>
__device__ __forceinline__ V pow2_dr(V r2, V dr) {
    using L = Lanes<V>;
    V acc = L::sqrt(r2), cur = r2;                       // r^(2^j - 1) after j rounds
#pragma unroll
    for (int j = 1; j < K; ++j) { acc = L::mul(acc, cur); if (j + 1 < K) cur = L::mul(cur, cur); }
    return L::fma(L::mul(acc, L::bc((float)(1 << K))), dr, L::bc(1.0f));
}

template <int K>
__device__ __forceinline__ void pow2_pair_impl(const ShapeDev& s, float2 px, float2 py, float2 pz, float2& d, uint32_t& suspect) {
    using L = Lanes<float2>;
    constexpr bool kBand = false;
    const float bail2 = s.bail2;
    float2 zx = px, zy = py, zz = pz, dr = L::bc(1.0f), r2;
    Amp ampa = amp_init(), ampb = amp_init();
    uint32_t done = 0u, left = s.max_iters;
    for (;;) {
        const float2 z2 = L::mul(zz, zz);
        const float2 w2 = L::fma(zx, zx, L::mul(zy, zy));
        r2 = L::add(w2, z2);
        CTC_PAIR_RADIUS_TEST()
        dr = pow2_dr<K, float2>(r2, dr);
        if (--left == 0u) break;
        float2 iw;
        pow2_step<K, float2>(zx, zy, zz, z2, w2, px, py, pz, iw);
        ampa.iwmax = max(ampa.iwmax, __float_as_int(iw.x));
        ampb.iwmax = max(ampb.iwmax, __float_as_int(iw.y));
    }
    CTC_PAIR_TAIL()
}

// FAST DE of two points for P = 2^K.  Bit k of `suspect`: half k needs the exact path.
template <int K>
__device__ __forceinline__ float2 mandelbulb_de_fast_pow2_pair(const ShapeDev& s, float2 px, float2 py, float2 pz, uint32_t& suspect) {
    float2 d = make_float2(0.0f, 0.0f);
    suspect = 0u;
    pow2_pair_impl<K>(s, px, py, pz, d, suspect);
    return d;
}
#undef CTC_PAIR_ESCAPE
#undef CTC_PAIR_RADIUS_TEST
#undef CTC_PAIR_TAIL

// log2 P when the packed power-of-two path serves P (2, 4, 16; 8 is the polynomial variant), else 0
__device__ __forceinline__ int pow2_path(uint32_t P) { return P == 2u ? 1 : P == 4u ? 2 : P == 16u ? 4 : 0; }

// FAST generic-power DE of one sample (config 4's P = 2, 4, 16, ...): trig-free complex binary powers
//   (z + i w)^P = r^P (cos P.theta + i sin P.theta),   ((x + i y)/w)^P = cos P.phi + i sin P.phi
// Exact mode itself is a tolerance path for these powers (CUDA's libm against glibc's), so no sign band
// is kept: only the z-axis / NaN rule sends a sample to the exact evaluation.
__device__ __forceinline__ float mandelbulb_de_fast_generic(const ShapeDev& s, float px, float py, float pz, bool& suspect) {
    const uint32_t P = s.power;
    const float bail2 = s.bail2;
    float zx = px, zy = py, zz = pz;
    float dr = 1.0f, r2;
    int iwmax = 0;
    uint32_t left = s.max_iters;
    for (;;) {
        const float z2 = zz * zz;
        const float w2 = fmaf(zx, zx, zy * zy);
        r2 = w2 + z2;
        if (r2 > bail2) {
            suspect = iwmax >= kAxisIwBits;
            return de_fast_epilogue(r2, dr);
        }
        const float r = fast_sqrt(r2);
        float rp1 = 1.0f;                       // r^(P-1)
        { float cur = r; uint32_t n = P - 1u; while (n) { if (n & 1u) rp1 *= cur; n >>= 1; if (n) cur *= cur; } }
        dr = fmaf((float)P * rp1, dr, 1.0f);
        if (--left == 0u) break;
        const float iw = fast_rsqrt(w2);
        iwmax = max(iwmax, __float_as_int(iw));
        const float w = w2 * iw;
        float ct, st, cp, sp;
        cpow(zz, w, P, ct, st);                 // r^P cos(P theta), r^P sin(P theta)
        cpow(zx * iw, zy * iw, P, cp, sp);      // cos(P phi), sin(P phi)
        zx = fmaf(st, cp, px); zy = fmaf(st, sp, py); zz = ct + pz;
    }
    suspect = iwmax >= kAxisIwBits || !(r2 == r2);
    return de_fast_epilogue(r2, dr);
}

// Sphere::min_distance_from (sphere.rs:33-35); cgmath magnitude = sqrt((x*x+y*y)+z*z)
__device__ __forceinline__ float sphere_de(const ShapeDev& s, float px, float py, float pz) {
    using M = MathExact;
    const float dx = M::sub(s.cx, px), dy = M::sub(s.cy, py), dz = M::sub(s.cz, pz);
    return M::sub(M::sqrt(M::add(M::add(M::mul(dx, dx), M::mul(dy, dy)), M::mul(dz, dz))), s.radius);
}

// Shape dispatch.  kVariant: 0 = Mandelbulb P=8, 1 = Mandelbulb generic P, 2 = Sphere.
enum : int { kVarP8 = 0, kVarGeneric = 1, kVarSphere = 2 };

// The exact evaluation as an out-of-line call: the fallback of suspect fast results (rare).
template <bool kP8>
__device__ __noinline__ float mandelbulb_de_exact_cold(const ShapeDev& s, float px, float py, float pz) {
    return mandelbulb_de_exact<kP8>(s, px, py, pz);
}

// One sample.  In fast mode a result whose sign cannot be trusted (kBand = true) or that touched the
// z axis / went NaN (always) is replaced by the exact evaluation, so the SIGN FIELD of fast mode is
// the exact mode's (for P = 8: the reference's) and fast mode differs in value only.
template <bool kFast, int kVariant, bool kBand = true>
__device__ __forceinline__ float shape_de(const ShapeDev& s, float px, float py, float pz) {
    if (kVariant == kVarSphere) return sphere_de(s, px, py, pz);
    if (kFast) {
        // (generic powers: exact mode itself is a tolerance path there -- CUDA's libm against glibc's -- so the
        // sign band would buy nothing; only the z-axis / NaN rule sends a sample to the exact evaluation)
        bool suspect;
        float d = (kVariant == kVarP8) ? mandelbulb_de_fast_p8<kBand>(s, px, py, pz, suspect)
                                       : mandelbulb_de_fast_generic(s, px, py, pz, suspect);
        if (suspect) d = mandelbulb_de_exact_cold<kVariant == kVarP8>(s, px, py, pz);
        return d;
    }
    return mandelbulb_de_exact<kVariant == kVarP8>(s, px, py, pz);
}

}  // namespace ctc
