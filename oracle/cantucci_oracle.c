/*
 * cantucci_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY; see cantucci_oracle.h).
 *
 * A plain-C restatement of cantucci's hot path, written from the Rust sources
 * under /root/reference.  Every f32 expression keeps the reference's
 * evaluation order; build with -ffp-contract=off (rustc never contracts to
 * FMA) and without -ffast-math.  libm supplies logf/acosf/atan2f/sinf/cosf --
 * the same glibc entry points Rust's f32::ln/acos/atan2/sin/cos resolve to on
 * x86-64 Linux.
 *
 * PARITY UNPINNED: the reference has no golden vectors or tests for this path
 * and cannot be built here; see the header comment in cantucci_oracle.h.
 */
#define _GNU_SOURCE
#include "cantucci_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

/* ------------------------------------------------------------------------ */
/* small helpers                                                              */
/* ------------------------------------------------------------------------ */

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float    u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* f32::is_sign_positive: the sign BIT, so -0.0 and x86's default NaN
 * (0xFFC00000) are "negative". */
static inline int sign_positive(float f) { return (f2u(f) >> 31) == 0; }

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* f32::powi with a compile-time-constant exponent: rustc emits llvm.powi.f32,
 * which LLVM's SelectionDAGBuilder (ExpandPowI) unrolls into the binary
 * square-and-multiply chain below (same association as compiler-rt's
 * __powisf2).  Lives in the toolchain, not in /root/reference;
 * call sites: src/shape/mandelbulb.rs:73,121,136. */
static inline float powi_f32(float x, int n) {
    unsigned v = n < 0 ? (unsigned)(-n) : (unsigned)n;
    if (v == 0) return 1.0f;
    float res = 0.0f, cur = x;
    int have = 0;
    while (v) {
        if (v & 1u) { res = have ? res * cur : cur; have = 1; }
        v >>= 1;
        if (v) cur = cur * cur;
    }
    return n < 0 ? 1.0f / res : res;
}

/* ------------------------------------------------------------------------ */
/* Vec3 (src/shape/mandelbulb.rs:375-471)                                     */
/* ------------------------------------------------------------------------ */

/* Vec3::magnitude, sse4.1 flavour (mandelbulb.rs:411-418): _mm_dp_ps with
 * mask 0x71 sums (x*x + y*y) + (z*z + 0.0), then _mm_sqrt_ss. */
static inline float vec3_magnitude(const float v[3]) {
    float s = (v[0] * v[0] + v[1] * v[1]) + (v[2] * v[2] + 0.0f);
    return sqrtf(s);
}

/* Vec3::is_on_z_axis (mandelbulb.rs:420-426): x and y are +-0.0. */
static inline int vec3_is_on_z_axis(const float v[3]) {
    return ((f2u(v[0]) & 0x7FFFFFFFu) == 0) && ((f2u(v[1]) & 0x7FFFFFFFu) == 0);
}

/* ------------------------------------------------------------------------ */
/* rotate (src/shape/mandelbulb.rs:96-200)                                    */
/* ------------------------------------------------------------------------ */

/* rotate_on_z_axis::<P> (mandelbulb.rs:114-126) */
static void rotate_on_z_axis(uint32_t P, const float p[3], float out[3]) {
    float old_radius = vec3_magnitude(p);
    float theta = acosf(p[2] / old_radius);
    float new_radius = powi_f32(old_radius, (int)P);
    theta = theta * (float)P;
    out[0] = 0.0f;
    out[1] = 0.0f;
    out[2] = new_radius * cosf(theta);
}

/* rotate_inner_px_generic::<P> (mandelbulb.rs:128-146) */
static void rotate_inner_px_generic(uint32_t P, const float p[3], float out[3]) {
    float old_radius = vec3_magnitude(p);
    float theta = acosf(p[2] / old_radius);
    float phi = atan2f(p[1], p[0]);
    float new_radius = powi_f32(old_radius, (int)P);
    theta = theta * (float)P;
    phi = phi * (float)P;
    /* new_radius * Vec3::new(..) == Vec3 * f32 == _mm_mul_ps(v, splat(new_radius)) */
    float vx = sinf(theta) * cosf(phi);
    float vy = sinf(phi) * sinf(theta);
    float vz = cosf(theta);
    out[0] = vx * new_radius;
    out[1] = vy * new_radius;
    out[2] = vz * new_radius;
}

/* rotate_inner_p8_scalar (mandelbulb.rs:148-200): left-to-right, no FMA. */
static void rotate_inner_p8_scalar(const float p[3], float out[3]) {
    float x = p[0], y = p[1], z = p[2];

    float x2 = x * x;
    float x4 = x2 * x2;
    float x6 = x4 * x2;
    float x8 = x4 * x4;

    float y2 = y * y;
    float y4 = y2 * y2;
    float y6 = y4 * y2;
    float y8 = y4 * y4;

    float z2 = z * z;
    float z4 = z2 * z2;
    float z6 = z4 * z2;
    float z8 = z4 * z4;

    float rxy2 = x2 + y2;
    float rxy4 = rxy2 * rxy2;
    float rxy6 = rxy2 * rxy4;
    float rxy8 = rxy4 * rxy4;

    float a = 1.0f + (((z8 - (28.0f * z6) * rxy2) + (70.0f * z4) * rxy4) - (28.0f * z2) * rxy6) / rxy8;

    out[0] = a * ((((x8 - (28.0f * x6) * y2) + (70.0f * x4) * y4) - (28.0f * x2) * y6) - y8);
    out[1] = (((8.0f * a) * x) * y) * (((x6 - (7.0f * x4) * y2) + (7.0f * x2) * y4) - y6);
    out[2] = (((8.0f * z) * sqrtf(rxy2)) * (z2 - rxy2)) * ((z4 - (6.0f * z2) * rxy2) + rxy4);
}

/* rotate::<P> (mandelbulb.rs:96-112) */
static inline void rotate(uint32_t P, const float p[3], float out[3]) {
    if (vec3_is_on_z_axis(p)) { rotate_on_z_axis(P, p, out); return; }
    if (P == 8) rotate_inner_p8_scalar(p, out);
    else        rotate_inner_px_generic(P, p, out);
}

void orc_rotate_p8_scalar(const float in[3], float out[3]) { rotate_inner_p8_scalar(in, out); }
void orc_rotate_generic(uint32_t power, const float in[3], float out[3]) { rotate_inner_px_generic(power, in, out); }
void orc_rotate(uint32_t power, const float in[3], float out[3]) { rotate(power, in, out); }

/* ------------------------------------------------------------------------ */
/* Shape::min_distance_from                                                   */
/* ------------------------------------------------------------------------ */

/* Mandelbulb::<P>::min_distance_from (mandelbulb.rs:59-79) */
static float mandelbulb_de(const orc_shape *s, const float p[3], orc_de_info *info) {
    const uint32_t P = s->power;
    float z[3] = { p[0], p[1], p[2] };
    float dr = 1.0f;
    float r = 0.0f;
    uint32_t iters = 0, bailed = 0;
    float margin = INFINITY;

    for (uint64_t i = 0; i < s->max_iters; i++) {
        r = vec3_magnitude(z);
        if (info) {
            float m = fabsf(r - s->bailout) / s->bailout;
            if (m < margin) margin = m;
        }
        if (r > s->bailout) { bailed = 1; break; }

        dr = powi_f32(r, (int)P - 1) * (float)P * dr + 1.0f;
        float rz[3];
        rotate(P, z, rz);
        z[0] = rz[0] + p[0];
        z[1] = rz[1] + p[1];
        z[2] = rz[2] + p[2];
        iters++;
    }

    float ln_r = logf(r) * r;
    float out = 0.5f * ln_r / dr;
    if (info) { info->iters = iters; info->bailed = bailed; info->r = r; info->dr = dr; info->min_margin = margin; }
    return out;
}

/* Sphere::min_distance_from (src/shape/sphere.rs:33-35): (center - p).magnitude() - radius;
 * cgmath magnitude = sqrt((x*x + y*y) + z*z). */
static float sphere_de(const orc_shape *s, const float p[3]) {
    float dx = s->center[0] - p[0], dy = s->center[1] - p[1], dz = s->center[2] - p[2];
    return sqrtf((dx * dx + dy * dy) + dz * dz) - s->radius;
}

float orc_min_distance_from_info(const orc_shape *s, const float p[3], orc_de_info *info) {
    if (s->kind == ORC_SHAPE_SPHERE) {
        if (info) memset(info, 0, sizeof *info);
        return sphere_de(s, p);
    }
    return mandelbulb_de(s, p, info);
}

float orc_min_distance_from(const orc_shape *s, const float p[3]) {
    return orc_min_distance_from_info(s, p, NULL);
}

/* impl_batch_methods! (src/shape/util.rs:3-5) */
void orc_batch_min_distance_from(const orc_shape *s, const float *xyz, size_t n, float *out) {
    for (size_t i = 0; i < n; i++) out[i] = orc_min_distance_from(s, xyz + 3 * i);
}

/* ------------------------------------------------------------------------ */
/* glibc 2.39 logf, FMA build (sysdeps/ieee754/flt-32/e_logf.c, __logf_fma).  */
/* Third-party, not under /root/reference.  Restated so the CUDA exact mode  */
/* can be bit-compared; constants are __logf_data (logf_data.c).              */
/* ------------------------------------------------------------------------ */
static const double LOGF_TAB[16][2] = {
    { 0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2 },
    { 0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2 },
    { 0x1.49539f0f010b0p+0, -0x1.01eae7f513a67p-2 },
    { 0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3 },
    { 0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3 },
    { 0x1.25e227b0b8ea0p+0, -0x1.1aa2bc79c8100p-3 },
    { 0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4 },
    { 0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4 },
    { 0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5 },
    { 0x1.0000000000000p+0,  0x0.0p+0 },
    { 0x1.e608cfd9a47acp-1,  0x1.aa5aa5df25984p-5 },
    { 0x1.ca4b31f026aa0p-1,  0x1.c5e53aa362eb4p-4 },
    { 0x1.b2036576afce6p-1,  0x1.526e57720db08p-3 },
    { 0x1.9c2d163a1aa2dp-1,  0x1.bc2860d224770p-3 },
    { 0x1.886e6037841edp-1,  0x1.1058bc8a07ee1p-2 },
    { 0x1.767dcf5534862p-1,  0x1.4043057b6ee09p-2 },
};
static const double LOGF_LN2 = 0x1.62e42fefa39efp-1;
static const double LOGF_A[3] = { -0x1.00ea348b88334p-2, 0x1.5575b0be00b6ap-2, -0x1.ffffef20a4123p-2 };

float orc_logf_glibc_fma(float x) {
    uint32_t ix = f2u(x);
    if (ix == 0x3f800000u) return 0.0f;
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
        if (ix * 2 == 0) return -INFINITY;              /* log(+-0) = -inf */
        if (ix == 0x7f800000u) return x;                /* log(inf) = inf */
        if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u) {
            /* __math_invalidf: (x - x) / (x - x); x86 gives the default NaN
             * 0xFFC00000 for finite/inf x and propagates a NaN operand. */
            return (x != x) ? x : u2f(0xFFC00000u);
        }
        ix = f2u(x * 0x1p23f);
        ix -= 23u << 23;
    }
    uint32_t tmp = ix - 0x3f330000u;
    int i = (int)((tmp >> 19) & 15u);
    int k = (int32_t)tmp >> 23;
    uint32_t iz = ix - (tmp & 0xff800000u);
    double invc = LOGF_TAB[i][0], logc = LOGF_TAB[i][1];
    double z = (double)u2f(iz);
    double r  = fma(z, invc, -1.0);
    double y0 = fma((double)k, LOGF_LN2, logc);
    double r2 = r * r;
    double y  = fma(LOGF_A[1], r, LOGF_A[2]);
    y = fma(LOGF_A[0], r2, y);
    y = fma(y, r2, y0 + r);
    return (float)y;
}

/* ------------------------------------------------------------------------ */
/* octree span maths                                                          */
/* ------------------------------------------------------------------------ */

/* SpanExt::center (src/octree/mod.rs:21-23): start + (end - start) / 2.0 */
void orc_span_center(const orc_span *s, float out[3]) {
    for (int c = 0; c < 3; c++) out[c] = s->start[c] + (s->end[c] - s->start[c]) / 2.0f;
}

/* create_spans (src/octree/mod.rs:315-329): child i = (x,y,z) bits, z lowest. */
void orc_create_spans(const orc_span *parent, orc_span out[8]) {
    float center[3];
    orc_span_center(parent, center);
    for (int i = 0; i < 8; i++) {
        for (int c = 0; c < 3; c++) {
            int hi = (i >> (2 - c)) & 1;
            out[i].start[c] = hi ? center[c] : parent->start[c];
            out[i].end[c]   = hi ? parent->end[c] : center[c];
        }
    }
}

/* ------------------------------------------------------------------------ */
/* MeshBuffer::naive_surface_nets (src/mesh/buffer.rs:58-391)                 */
/* ------------------------------------------------------------------------ */

/* GridTable index (src/util/grid.rs:45-48): x*size^2 + y*size + z */
static inline size_t gidx(uint32_t size, uint32_t x, uint32_t y, uint32_t z) {
    return (size_t)x * size * size + (size_t)y * size + z;
}

/* buffer.rs:64-67 */
static void expand_span(const orc_span *in, uint32_t resolution, orc_span *out) {
    for (int c = 0; c < 3; c++) {
        float overflow = (in->end[c] - in->start[c]) / (float)resolution;
        out->start[c] = in->start[c] + (-overflow);
        out->end[c]   = in->end[c] + overflow;
    }
}

/* pass 1 (buffer.rs:77-83) over cube(R+1) (src/util/iter.rs:30-49) */
static void sample_grid(const orc_shape *s, const orc_span *span /* expanded */, uint32_t R,
                        float *dists, uint64_t *hist, uint64_t *n_bailed) {
    float across[3];
    for (int c = 0; c < 3; c++) across[c] = span->end[c] - span->start[c];
    const uint32_t n = R + 1;
    const float fr = (float)R;
    size_t o = 0;
    for (uint32_t x = 0; x < n; x++)
        for (uint32_t y = 0; y < n; y++)
            for (uint32_t z = 0; z < n; z++) {
                float v[3] = { (float)x / fr, (float)y / fr, (float)z / fr };
                float p[3];
                for (int c = 0; c < 3; c++) p[c] = span->start[c] + across[c] * v[c];
                if (hist) {
                    orc_de_info info;
                    dists[o++] = orc_min_distance_from_info(s, p, &info);
                    hist[info.iters]++;
                    *n_bailed += info.bailed;
                } else {
                    dists[o++] = orc_min_distance_from(s, p);
                }
            }
}

static int check_args(const orc_span *span, uint32_t resolution) {
    /* buffer.rs:35-39 asserts; GridTable::fill_with asserts size >= 2 (grid.rs:25) */
    for (int c = 0; c < 3; c++) if (!(span->start[c] < span->end[c])) return 1;
    if (resolution == 0 || (resolution & (resolution - 1)) != 0) return 1;
    if (resolution < 2) return 1;
    return 0;
}

int orc_sample_grid(const orc_shape *s, const orc_span *span, uint32_t resolution, float *out) {
    if (check_args(span, resolution)) return 1;
    orc_span ex;
    expand_span(span, resolution, &ex);
    sample_grid(s, &ex, resolution, out, NULL, NULL);
    return 0;
}

/* per-sample completed-iteration counts (divergence studies, flop accounting) */
int orc_sample_grid_iters(const orc_shape *s, const orc_span *span, uint32_t resolution,
                          float *out, uint8_t *iters_out) {
    if (check_args(span, resolution)) return 1;
    orc_span ex;
    expand_span(span, resolution, &ex);
    float across[3];
    for (int c = 0; c < 3; c++) across[c] = ex.end[c] - ex.start[c];
    const uint32_t n = resolution + 1;
    const float fr = (float)resolution;
    size_t o = 0;
    for (uint32_t x = 0; x < n; x++)
        for (uint32_t y = 0; y < n; y++)
            for (uint32_t z = 0; z < n; z++) {
                float v[3] = { (float)x / fr, (float)y / fr, (float)z / fr };
                float p[3];
                for (int c = 0; c < 3; c++) p[c] = ex.start[c] + across[c] * v[c];
                orc_de_info info;
                out[o] = orc_min_distance_from_info(s, p, &info);
                iters_out[o] = info.iters > 255 ? 255 : (uint8_t)info.iters;
                o++;
            }
    return 0;
}

int orc_sample_grid_info(const orc_shape *s, const orc_span *span, uint32_t resolution,
                         float *out, uint64_t *iter_hist, uint64_t *n_bailed) {
    if (check_args(span, resolution)) return 1;
    orc_span ex;
    expand_span(span, resolution, &ex);
    sample_grid(s, &ex, resolution, out, iter_hist, n_bailed);
    return 0;
}

typedef struct { orc_vertex *v; size_t n, cap; } vbuf;
typedef struct { uint32_t *v; size_t n, cap; } ibuf;

static void vpush(vbuf *b, const orc_vertex *x) {
    if (b->n == b->cap) { b->cap = b->cap ? b->cap * 2 : 1024; b->v = realloc(b->v, b->cap * sizeof *b->v); }
    b->v[b->n++] = *x;
}
static void ipush6(ibuf *b, const uint32_t x[6]) {
    if (b->n + 6 > b->cap) { b->cap = b->cap ? b->cap * 2 : 6144; b->v = realloc(b->v, b->cap * sizeof *b->v); }
    memcpy(b->v + b->n, x, 6 * sizeof(uint32_t));
    b->n += 6;
}

/* plane (may be NULL): the sign bit-plane of the sample grid, bit j of word j/32 =
 * !f32::is_sign_positive(dists[j]), j = x*n^2 + y*n + z; (n^3 + 31)/32 words.  Checker output for the
 * parity gate (sign mismatches of the CUDA path are counted against it). */
static int generate_for_box_impl(const orc_shape *s, const orc_span *span_in, uint32_t R, orc_mesh *out, uint32_t *plane);

int orc_generate_for_box(const orc_shape *s, const orc_span *span_in, uint32_t R, orc_mesh *out) {
    return generate_for_box_impl(s, span_in, R, out, NULL);
}

static int generate_for_box_impl(const orc_shape *s, const orc_span *span_in, uint32_t R, orc_mesh *out, uint32_t *plane) {
    memset(out, 0, sizeof *out);
    if (check_args(span_in, R)) return 1;

    orc_span span;
    expand_span(span_in, R, &span);

    const double before_first = now_s();

    /* ---- first step (buffer.rs:77-83) ---- */
    const uint32_t n = R + 1;
    float *dists = malloc((size_t)n * n * n * sizeof(float));
    sample_grid(s, &span, R, dists, NULL, NULL);
    if (plane) {
        const size_t n3 = (size_t)n * n * n;
        memset(plane, 0, ((n3 + 31) / 32) * sizeof(uint32_t));
        for (size_t j = 0; j < n3; j++)
            if (!sign_positive(dists[j])) plane[j >> 5] |= 1u << (j & 31);
    }

    const double before_second = now_s();

    /* ---- second step (buffer.rs:97-275) ---- */
    vbuf vertices = { 0 };
    const float fr = (float)R;
    float step[3];
    for (int c = 0; c < 3; c++) step[c] = (span.end[c] - span.start[c]) / fr;   /* :101 */
    /* corner id = 4*dx + 2*dy + dz (:102-111) */
    float corner_offsets[8][3];
    for (int i = 0; i < 8; i++) {
        corner_offsets[i][0] = (i & 4) ? step[0] : 0.0f;
        corner_offsets[i][1] = (i & 2) ? step[1] : 0.0f;
        corner_offsets[i][2] = (i & 1) ? step[2] : 0.0f;
    }
    static const uint8_t EDGES[12][2] = {          /* :155-176 */
        {0, 4}, {1, 5}, {2, 6}, {3, 7},
        {0, 2}, {1, 3}, {4, 6}, {5, 7},
        {0, 1}, {2, 3}, {4, 5}, {6, 7},
    };
    /* normal delta (:257): 0.7 * (end - start) / R  == (0.7 * v) / R */
    float delta[3];
    for (int c = 0; c < 3; c++) delta[c] = (0.7f * (span.end[c] - span.start[c])) / fr;

    uint32_t *points = malloc((size_t)R * R * R * sizeof(uint32_t));
    size_t po = 0;
    int panicked = 0;

    for (uint32_t x = 0; x < R && !panicked; x++)
    for (uint32_t y = 0; y < R && !panicked; y++)
    for (uint32_t z = 0; z < R && !panicked; z++) {
        const float distances[8] = {                 /* :116-125 */
            dists[gidx(n, x, y, z)],         dists[gidx(n, x, y, z + 1)],
            dists[gidx(n, x, y + 1, z)],     dists[gidx(n, x, y + 1, z + 1)],
            dists[gidx(n, x + 1, y, z)],     dists[gidx(n, x + 1, y, z + 1)],
            dists[gidx(n, x + 1, y + 1, z)], dists[gidx(n, x + 1, y + 1, z + 1)],
        };
        /* :130-141 */
        int first = sign_positive(distances[0]);
        int all_same = 1;
        for (int i = 1; i < 8; i++) if (sign_positive(distances[i]) != first) { all_same = 0; break; }
        if (all_same) { points[po++] = UINT32_MAX; continue; }   /* :143-147 */

        /* :150-151 */
        float p0[3] = {
            span.start[0] + (float)x * step[0],
            span.start[1] + (float)y * step[1],
            span.start[2] + (float)z * step[2],
        };

        /* :187-250 -- fold((0, zero), |(count,sum),p| (count+1, sum+p)) */
        int count = 0;
        float sum[3] = { 0.0f, 0.0f, 0.0f };
        for (int e = 0; e < 12; e++) {
            int from = EDGES[e][0], to = EDGES[e][1];
            if (sign_positive(distances[from]) == sign_positive(distances[to])) continue;   /* :194-196 */
            float d_from, d_to;
            if (distances[from] < 0.0f) { d_from = distances[from];  d_to = distances[to]; }   /* :209-213 */
            else                        { d_from = -distances[from]; d_to = -distances[to]; }
            float weight_from;
            if (d_to == d_from) weight_from = 0.5f;                /* :217-218 */
            else { float dl = d_to - d_from; weight_from = (d_from + dl) / dl; }   /* :238-239 */
            /* lerp (src/math.rs:14-21,45-48): assert 0<=t<=1; a*(1-t) + b*t */
            if (!(weight_from >= 0.0f && weight_from <= 1.0f)) { panicked = 1; break; }
            float one_minus = 1.0f - weight_from;
            for (int c = 0; c < 3; c++) {
                float a = p0[c] + corner_offsets[from][c];
                float b = p0[c] + corner_offsets[to][c];
                float pt = a * one_minus + b * weight_from;
                sum[c] = sum[c] + pt;
            }
            count++;
        }
        if (panicked) break;
        float p[3];
        for (int c = 0; c < 3; c++) p[c] = 0.0f + (sum[c] / (float)count);    /* :250 */

        float dist_p = orc_min_distance_from(s, p);                           /* :254 */

        /* :256-266  unit_x() * d = (1*d, 0*d, 0*d) */
        float nrm[3];
        for (int c = 0; c < 3; c++) {
            float dpos = delta[c], dneg = -delta[c];
            float pp[3], pm[3];
            for (int k = 0; k < 3; k++) {
                float u = (k == c) ? 1.0f : 0.0f;
                pp[k] = p[k] + u * dpos;
                pm[k] = p[k] + u * dneg;
            }
            nrm[c] = orc_min_distance_from(s, pp) - orc_min_distance_from(s, pm);
        }
        /* cgmath normalize: v * (1.0 / sqrt((x*x + y*y) + z*z)) */
        float mag = sqrtf((nrm[0] * nrm[0] + nrm[1] * nrm[1]) + nrm[2] * nrm[2]);
        float inv = 1.0f / mag;
        orc_vertex vert;
        for (int c = 0; c < 3; c++) { vert.position[c] = p[c]; vert.normal[c] = nrm[c] * inv; }
        vert.distance_from_surface = dist_p;
        vpush(&vertices, &vert);
        points[po++] = (uint32_t)vertices.n - 1;                              /* :268-274 */
    }

    const double before_third = now_s();

    /* ---- third step (buffer.rs:288-372) ---- */
    ibuf indices = { 0 };
    if (!panicked)
    for (uint32_t x = 0; x < R; x++)
    for (uint32_t y = 0; y < R; y++)
    for (uint32_t z = 0; z < R; z++) {
        float d = dists[gidx(n, x, y, z)];
        int base_sign = sign_positive(d);
        int neg = d < 0.0f;

        if (y > 0 && z > 0 && base_sign != sign_positive(dists[gidx(n, x + 1, y, z)])) {   /* :302-323 */
            uint32_t v0 = points[gidx(R, x, y - 1, z - 1)], v1 = points[gidx(R, x, y - 1, z)];
            uint32_t v2 = points[gidx(R, x, y, z - 1)],     v3 = points[gidx(R, x, y, z)];
            uint32_t a[6] = { v0, v2, v1, v1, v2, v3 }, b[6] = { v0, v1, v2, v1, v3, v2 };
            ipush6(&indices, neg ? a : b);
        }
        if (x > 0 && z > 0 && base_sign != sign_positive(dists[gidx(n, x, y + 1, z)])) {   /* :326-347 */
            uint32_t v0 = points[gidx(R, x - 1, y, z - 1)], v1 = points[gidx(R, x - 1, y, z)];
            uint32_t v2 = points[gidx(R, x, y, z - 1)],     v3 = points[gidx(R, x, y, z)];
            uint32_t a[6] = { v0, v1, v2, v1, v3, v2 }, b[6] = { v0, v2, v1, v1, v2, v3 };
            ipush6(&indices, neg ? a : b);
        }
        if (x > 0 && y > 0 && base_sign != sign_positive(dists[gidx(n, x, y, z + 1)])) {   /* :350-371 */
            uint32_t v0 = points[gidx(R, x - 1, y - 1, z)], v1 = points[gidx(R, x - 1, y, z)];
            uint32_t v2 = points[gidx(R, x, y - 1, z)],     v3 = points[gidx(R, x, y, z)];
            uint32_t a[6] = { v0, v2, v1, v1, v2, v3 }, b[6] = { v0, v1, v2, v1, v3, v2 };
            ipush6(&indices, neg ? a : b);
        }
    }

    const double after_third = now_s();

    free(dists);
    free(points);
    if (panicked) {
        free(vertices.v); free(indices.v);
        out->panicked = 1;
        return 0;
    }
    out->vertices = vertices.v;  out->n_vertices = vertices.n;
    out->indices = indices.v;    out->n_indices = indices.n;
    out->first_s = before_second - before_first;
    out->second_s = before_third - before_second;
    out->third_s = after_third - before_third;
    return 0;
}

void orc_mesh_free(orc_mesh *m) {
    free(m->vertices); free(m->indices);
    m->vertices = NULL; m->indices = NULL; m->n_vertices = m->n_indices = 0;
}

/* ------------------------------------------------------------------------ */
/* thread pool driver (src/mesh/mod.rs:61-62, 141-148)                        */
/* ------------------------------------------------------------------------ */

typedef struct {
    const orc_shape *shape; const orc_span *spans; size_t nspans; uint32_t R;
    orc_mesh *meshes; atomic_size_t next; int grids_only; double checksum; pthread_mutex_t mu;
    uint32_t *planes; size_t plane_words;
} pool_job;

static void *pool_worker(void *arg) {
    pool_job *j = arg;
    double local = 0.0;
    float *grid = NULL;
    if (j->grids_only) grid = malloc((size_t)(j->R + 1) * (j->R + 1) * (j->R + 1) * sizeof(float));
    for (;;) {
        size_t i = atomic_fetch_add(&j->next, 1);
        if (i >= j->nspans) break;
        if (j->grids_only) {
            orc_sample_grid(j->shape, &j->spans[i], j->R, grid);
            size_t n = (size_t)(j->R + 1) * (j->R + 1) * (j->R + 1);
            for (size_t k = 0; k < n; k += 97) if (grid[k] == grid[k]) local += grid[k];
        } else {
            generate_for_box_impl(j->shape, &j->spans[i], j->R, &j->meshes[i],
                                  j->planes ? j->planes + i * j->plane_words : NULL);
        }
    }
    free(grid);
    pthread_mutex_lock(&j->mu); j->checksum += local; pthread_mutex_unlock(&j->mu);
    return NULL;
}

static double run_pool(pool_job *j, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    pthread_t *th = malloc(sizeof(pthread_t) * (size_t)nthreads);
    atomic_init(&j->next, 0);
    pthread_mutex_init(&j->mu, NULL);
    double t0 = now_s();
    for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, pool_worker, j);
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    double t1 = now_s();
    pthread_mutex_destroy(&j->mu);
    free(th);
    return t1 - t0;
}

double orc_generate_for_boxes_mt(const orc_shape *s, const orc_span *spans, size_t nspans,
                                 uint32_t resolution, int nthreads, orc_mesh *meshes) {
    pool_job j = { .shape = s, .spans = spans, .nspans = nspans, .R = resolution, .meshes = meshes };
    return run_pool(&j, nthreads);
}

double orc_generate_for_boxes_signs_mt(const orc_shape *s, const orc_span *spans, size_t nspans,
                                       uint32_t resolution, int nthreads, orc_mesh *meshes, uint32_t *planes) {
    const size_t n = (size_t)resolution + 1;
    pool_job j = { .shape = s, .spans = spans, .nspans = nspans, .R = resolution, .meshes = meshes,
                   .planes = planes, .plane_words = (n * n * n + 31) / 32 };
    return run_pool(&j, nthreads);
}

double orc_sample_grids_mt(const orc_shape *s, const orc_span *spans, size_t nspans,
                           uint32_t resolution, int nthreads, double *checksum) {
    pool_job j = { .shape = s, .spans = spans, .nspans = nspans, .R = resolution, .grids_only = 1 };
    double t = run_pool(&j, nthreads);
    if (checksum) *checksum = j.checksum;
    return t;
}

int orc_hardware_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}
