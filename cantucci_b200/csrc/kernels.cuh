// kernels.cuh -- the hot-path kernels (sm_100a).
//
//   K1  sample_grids_kernel   pass 1 of naive_surface_nets (mesh/buffer.rs:77-83); also emits the
//                             sign bit-plane (1 bit per sample, reference layout)
//   K3  de_batch_kernel       Shape::batch_min_distance_from (shape/mod.rs:89)
//   E1  classify_count_kernel cell / edge sign classification from the bit-plane, 32 cells per thread
//                             (buffer.rs:116-147, 299-350): counts per chunk and span
//   E2a span_scan_kernel      order-preserving prefix over span totals, offset tables
//   E2b emit_lists_kernel     active-cell list, quad list, rank-query tables
//   E3  vertex_kernel         per-active-cell vertex (buffer.rs:150-274)
//   E4  quad_kernel           one thread per quad, reference emission order (buffer.rs:288-372)
//   N1  ray_march_kernel      get_focii's sphere tracing (mesh/mod.rs:229-241)
//   +   iteration_stats_kernel / fma_peak_kernel: measurement aids for bench.py
//
// Ordering contract: vertex ids are the rank of the cell among active cells in
// cube(R) order (x outer, z fastest; util/iter.rs:30-49), quads are emitted
// corner by corner, +x, +y, +z edge in that order.  All compaction is done by
// prefix sums (no atomics appends), so index buffers are identical to the
// reference's wherever the sign field is.
#pragma once

#include "de_device.cuh"

namespace ctc {

// Host-computed (IEEE, no FMA) per-span geometry after the skirt expansion
// (buffer.rs:64-67, 77, 101, 257).
struct SpanGeom {
    float s[3];       // expanded span.start
    float across[3];  // expanded end - start
    float step[3];    // across / R
    float delta[3];   // (0.7 * across) / R
};

// Device-resident call state: running totals chain launch groups without any
// host synchronisation.
struct MeshState {
    unsigned long long total_v;     // vertices required so far (all groups)
    unsigned long long total_q;     // quads required so far
    unsigned long long group_base_v;
    unsigned long long group_base_q;
    unsigned int group_v;           // vertices in the current group
    unsigned int group_q;
    unsigned int overflow;          // 1 = capacity exceeded somewhere
    unsigned int panic_span;        // min global span index whose lerp factor left [0,1]; 0xFFFFFFFF = none
    unsigned int wire_overflow;     // 1 = a vertex id did not fit the packed 16-bit quad record
    unsigned int pad_;
    unsigned long long suspects;    // fast mode: samples re-evaluated with the exact arithmetic (all groups)
    unsigned long long sign_fixups; // ... of which the sign differed from the fast evaluation's
};

constexpr int kThreads = 256;
// K1's CTA size (A/B on the benched volume, scripts/gpu_ab.py: 256 threads 4.49 ms, 128: 4.40, 64: 4.45 -- four warp
// walks per CTA retire and refill an SM's slots more evenly than eight)
#ifndef CTC_K1_THREADS
#define CTC_K1_THREADS 128
#endif
constexpr int kK1Threads = CTC_K1_THREADS;
#ifndef CTC_E3_THREADS
#define CTC_E3_THREADS 256
#endif
#ifndef CTC_E3_MINBLOCKS
#define CTC_E3_MINBLOCKS (4 * 256 / CTC_E3_THREADS)
#endif
constexpr int kE3Threads = CTC_E3_THREADS;
constexpr uint32_t kMaxChunkWords = 256;  // one word per thread in E1+E2

// ---------------------------------------------------------------------------
// sample index -> lattice coordinates
// ---------------------------------------------------------------------------
// One-thread-per-sample enumeration (used for R < 32 and by the iteration-statistics kernel; K1's
// main path for R >= 32 walks lattice columns instead, see sample_grids_kernel).  Every warp is full
// and covers a compact 2x4x4 brick of the R^3 core: iteration counts are spatially coherent, bricks
// lose ~1-4 % to divergence where 32-long z rows lose ~13 %.  The (R+1)^3 lattice = R^3 core + three
// R^2 faces + three R-long edges + 1 corner; every piece has power-of-two extent, so decoding is
// shifts and masks only.
__device__ __forceinline__ void decode_sample(uint32_t i, uint32_t R, uint32_t lg,
                                              uint32_t& x, uint32_t& y, uint32_t& z) {
    const uint32_t R2 = R << lg, R3 = R2 << lg;
    if (i < R3) {
        if (lg >= 2) {
            const uint32_t b = i >> 5;
            const uint32_t bz = b & ((R >> 2) - 1u);
            const uint32_t by = (b >> (lg - 2)) & ((R >> 2) - 1u);
            const uint32_t bx = b >> (2 * lg - 4);
            x = (bx << 1) | ((i >> 4) & 1u);
            y = (by << 2) | ((i >> 2) & 3u);
            z = (bz << 2) | (i & 3u);
        } else {
            x = i >> (2 * lg); y = (i >> lg) & (R - 1u); z = i & (R - 1u);
        }
        return;
    }
    i -= R3;
    if (i < 3u * R2) {
        const uint32_t f = i >> (2 * lg);          // which face
        const uint32_t u = (i >> lg) & (R - 1u), v = i & (R - 1u);
        if (f == 0)      { x = R; y = u; z = v; }
        else if (f == 1) { x = u; y = R; z = v; }
        else             { x = u; y = v; z = R; }
        return;
    }
    i -= 3u * R2;
    const uint32_t e = i >> lg, t = i & (R - 1u);
    if (e == 0)      { x = R; y = R; z = t; }
    else if (e == 1) { x = R; y = t; z = R; }
    else if (e == 2) { x = t; y = R; z = R; }
    else             { x = R; y = R; z = R; }
}

// ---------------------------------------------------------------------------
// K1: sample grids; gridDim.y = spans of the group.
//
// Besides the f32 grid the kernel emits the SIGN BIT-PLANE of the grid: bit j of
// `sign_bits` (per span, j = x*n^2 + y*n + z, the reference's GridTable index)
// is f32::is_sign_positive() == false.  Everything the mesher decides
// (buffer.rs:130-147, 194-196, 299-350) is a function of these bits, so the
// extraction passes never re-read the 4-byte samples except at active cells.
// The plane must be zeroed before the launch.
//
// Blocks [0, core_blocks), R >= 32: a warp owns a 2x4xL block of the R^3 core
// (or a 1x8xL / 8x1xL block of the x = R / y = R face; L = 64 for R >= 64, else 32)
// and walks it in L/8 steps of one 2x4x8 brick.  Every lane evaluates TWO samples
// of its lattice column per step (z and z + 4) in packed FP32 (FFMA2/FMUL2/FADD2),
// so iteration counts inside a warp stay coherent, the per-warp set-up (decode,
// geometry load, x/y position, the first iteration's azimuth factors) is paid once
// per 512 samples, and the ballots assemble the row words of the block's sign bits
// without shared memory.  Remaining blocks: the z = R face, edges and corner (and
// everything when R < 32), one thread per sample.
//
// Fast mode is SIGN-EXACT: a sample whose sign cannot be trusted (fast_suspect_*,
// de_device.cuh) is appended to the group's suspect list and re-evaluated with
// the exact arithmetic by fixup_suspects_kernel, which also repairs the plane.
// ---------------------------------------------------------------------------
struct SuspectList {
    uint2* entries;          // (span in group, sample index j | fast sign << 31); j < 1025^3 < 2^31
    unsigned int* count;     // appended so far (may exceed cap: the overflow was evaluated in place)
    uint32_t cap;
};

// fast / power 8: 64 registers, 4 CTAs per SM (A/B on the benched volume: 3 CTAs +8 %, 5 CTAs +1.5 %)
#ifndef CTC_K1_MINBLOCKS
#define CTC_K1_MINBLOCKS 4
#endif

template <bool kFast, int kVariant>
__global__ void __launch_bounds__(kK1Threads, kFast && kVariant == kVarP8 ? CTC_K1_MINBLOCKS * (256 / kK1Threads) : 1)
sample_grids_kernel(ShapeDev sh, const SpanGeom* __restrict__ geom, uint32_t R, uint32_t lg, float inv_r, float dvz8,
                    float* __restrict__ grids, size_t grid_stride,
                    uint32_t* __restrict__ sign_bits, uint32_t sign_stride /* words per span, 0 = no plane */,
                    uint32_t core_blocks, uint32_t lgw /* log2(32-sample words per warp walk): 0 or 1 */,
                    SuspectList sl) {
    const uint32_t n = R + 1u;
    const SpanGeom g = geom[blockIdx.y];
    float* __restrict__ grid = grids + (size_t)blockIdx.y * grid_stride;
    uint32_t* __restrict__ plane = sign_bits + (size_t)blockIdx.y * sign_stride;
    if (blockIdx.x < core_blocks) {
        // warp-blocks of L = 32 << lgw z-samples: [0, R^3/(8L)) core 2x4xL; then R^2/(8L) blocks 1x8xL of
        // the x = R face; then R^2/(8L) blocks 8x1xL of the y = R face.  Each is walked in L/8 brick steps.
        const uint32_t lane = threadIdx.x & 31u;
        const uint32_t wb = blockIdx.x * (kK1Threads / 32) + (threadIdx.x >> 5);
        const uint32_t lgL = 5u + lgw;
        const uint32_t n_core = 1u << (3 * lg - 3 - lgL), n_face = 1u << (2 * lg - 3 - lgL);
        if (wb >= n_core + 2u * n_face) return;
        uint32_t x, y, zb;
        if (wb < n_core) {
            zb = wb & ((R >> lgL) - 1u);
            x = ((wb >> (2 * lg - lgL - 2)) << 1) | (lane >> 4);
            y = (((wb >> (lg - lgL)) & ((R >> 2) - 1u)) << 2) | ((lane >> 2) & 3u);
        } else {
            const uint32_t f = wb - n_core, ff = f & (n_face - 1u);
            zb = ff & ((R >> lgL) - 1u);
            const uint32_t t = ((ff >> (lg - lgL)) << 3) | (lane >> 2);
            if (f < n_face) { x = R; y = t; } else { x = t; y = R; }
        }
        const uint32_t z = (zb << lgL) | (lane & 3u);   // this lane's samples of a step: z and z + 4
        // v = (x,y,z) as f32 / R  (exact: R is a power of two);  p = start + across * v  (buffer.rs:79-80)
        const float px = __fadd_rn(g.s[0], __fmul_rn(g.across[0], __fmul_rn((float)x, inv_r)));
        const float py = __fadd_rn(g.s[1], __fmul_rn(g.across[1], __fmul_rn((float)y, inv_r)));
        // exact increments: multiples of 1/R in [0,1].  The position arithmetic stays SCALAR: ptxas fuses a
        // packed mul.rn.f32x2 + add.rn.f32x2 pair into one FFMA2 (even with -fmad=false), which would change
        // the sample positions by an ulp (buffer.rs:79-80 is a multiply, then an add).
        float vza = __fmul_rn((float)z, inv_r), vzb = __fmul_rn((float)(z + 4u), inv_r);
        float* out = grid + ((size_t)x * n + y) * n + z;                         // util/grid.rs:45-48
        const float gs2 = g.s[2], ga2 = g.across[2];
        const uint32_t j_row = (x * n + y) * n + (zb << lgL);     // plane bit of this lane's row at the block's first z
        const bool writer = sign_stride != 0u && (lane & 3u) == 0u;
        // the on-axis special case of the column paths is tested once per warp, not once per sample
        const bool any_axis = (kFast || kVariant == kVarP8) && __any_sync(0xffffffffu, px == 0.0f && py == 0.0f);
        // 4 steps fill one 32-bit row word of the sign plane, which is then OR-ed in (two atomics).  Every lane
        // pushes the sign bits of its own samples into `acc` (one funnel shift each: no ballots in the walk);
        // after the 4 steps the 8 bits are spread to their places 8 j + 4 half + zl of the row word and the
        // four lanes of a row OR their parts together (two shuffles per 256 samples).
        // DE_PAIR sets `float2 d` from (px, py, pz.x) and (px, py, pz.y).
        const uint32_t zl = lane & 3u;
#define CTC_K1_STEPS(...)                                                                          \
        _Pragma("unroll 1")                                                                        \
        for (uint32_t h = 0; h < (1u << lgw); ++h) {                                               \
            uint32_t acc = 0;                                                                      \
            _Pragma("unroll 1")                                                                    \
            for (int j = 0; j < 4; ++j) {                                                          \
                const float2 pz = make_float2(__fadd_rn(gs2, __fmul_rn(ga2, vza)),                 \
                                              __fadd_rn(gs2, __fmul_rn(ga2, vzb)));                \
                float2 d;                                                                          \
                __VA_ARGS__                                                                        \
                out[0] = d.x;                                                                      \
                out[4] = d.y;                                                                      \
                out += 8;                                                                          \
                acc = __funnelshift_l(__float_as_uint(d.x), acc, 1);     /* push k = 2 j + half */ \
                acc = __funnelshift_l(__float_as_uint(d.y), acc, 1);                               \
                vza = __fadd_rn(vza, dvz8);                                                        \
                vzb = __fadd_rn(vzb, dvz8);                                                        \
            }                                                                                      \
            if (sign_stride != 0u) {                                                               \
                uint32_t t = __brev(acc) >> 24;                          /* push k at bit k */     \
                t = (t | (t << 12)) & 0x000F000Fu;                                                 \
                t = (t | (t << 6)) & 0x03030303u;                                                  \
                t = (t | (t << 3)) & 0x11111111u;                        /* push k at bit 4 k */   \
                t <<= zl;                                                                          \
                t |= __shfl_xor_sync(0xffffffffu, t, 1);                                           \
                const uint32_t word = t | __shfl_xor_sync(0xffffffffu, t, 2);                      \
                if (writer && word != 0u) {                                                        \
                    const uint32_t j0 = j_row + (h << 5);                                          \
                    const uint32_t sft = j0 & 31u;                                                 \
                    atomicOr(&plane[j0 >> 5], word << sft);                                        \
                    if (sft) atomicOr(&plane[(j0 >> 5) + 1u], word >> (32u - sft));                \
                }                                                                                  \
            }                                                                                      \
        }
        if (any_axis) {
            CTC_K1_STEPS(d.x = shape_de<kFast, kVariant>(sh, px, py, pz.x); d.y = shape_de<kFast, kVariant>(sh, px, py, pz.y);)
        } else if (kFast && kVariant == kVarP8) {
            // (px, py) is fixed along the walk: hoist the first iteration's azimuth factors
            const ColumnFastP8 col = column_fast_p8(px, py);
            const uint32_t lt = (1u << lane) - 1u;
            CTC_K1_STEPS(
                uint32_t susp;
                d = mandelbulb_de_fast_p8_column_pair<true>(sh, px, py, pz, col, susp);
                if (__any_sync(0xffffffffu, susp != 0u)) {   /* rare: queue the suspects for the exact re-evaluation */
                    const uint32_t ma = __ballot_sync(0xffffffffu, susp & 1u);
                    const uint32_t mb = __ballot_sync(0xffffffffu, susp & 2u);
                    uint32_t base = 0;
                    if (lane == 0u) base = atomicAdd(sl.count, (unsigned int)(__popc(ma) + __popc(mb)));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    const uint32_t jj = (uint32_t)(out - grid);
                    if (susp & 1u) {
                        const uint32_t slot = base + __popc(ma & lt);
                        if (slot < sl.cap) sl.entries[slot] = make_uint2(blockIdx.y, jj | (__float_as_uint(d.x) & 0x80000000u));
                        else d.x = mandelbulb_de_exact_cold<true>(sh, px, py, pz.x);
                    }
                    if (susp & 2u) {
                        const uint32_t slot = base + __popc(ma) + __popc(mb & lt);
                        if (slot < sl.cap) sl.entries[slot] = make_uint2(blockIdx.y, (jj + 4u) | (__float_as_uint(d.y) & 0x80000000u));
                        else d.y = mandelbulb_de_exact_cold<true>(sh, px, py, pz.y);
                    }
                })
        } else if (!kFast && kVariant == kVarP8) {
            // exact arithmetic: hoist the x/y-only sub-expressions of the first iteration (bit-identical CSE)
            const ColumnExactP8 col = column_exact_p8(px, py);
            CTC_K1_STEPS(d.x = mandelbulb_de_exact_p8_column(sh, px, py, pz.x, col); d.y = mandelbulb_de_exact_p8_column(sh, px, py, pz.y, col);)
        } else if (kFast && kVariant == kVarGeneric && pow2_path(sh.power) != 0) {
            // P = 2, 4, 16: packed pairs with the squarings unrolled (warp-uniform switch)
            const float2 px2 = make_float2(px, px), py2 = make_float2(py, py);
#define CTC_K1_POW2(K)                                                                                          \
            CTC_K1_STEPS(                                                                                       \
                uint32_t susp;                                                                                  \
                d = mandelbulb_de_fast_pow2_pair<K>(sh, px2, py2, pz, susp);                                    \
                if (susp & 1u) d.x = mandelbulb_de_exact_cold<false>(sh, px, py, pz.x);                         \
                if (susp & 2u) d.y = mandelbulb_de_exact_cold<false>(sh, px, py, pz.y);)
            const int k = pow2_path(sh.power);
            if (k == 1) { CTC_K1_POW2(1) } else if (k == 2) { CTC_K1_POW2(2) } else { CTC_K1_POW2(4) }
#undef CTC_K1_POW2
        } else {
            CTC_K1_STEPS(d.x = shape_de<kFast, kVariant>(sh, px, py, pz.x); d.y = shape_de<kFast, kVariant>(sh, px, py, pz.y);)
        }
#undef CTC_K1_STEPS
        return;
    }
    // one thread per remaining sample: everything when R < 32, else the z = R face, the three edges
    // x=y=R / x=z=R / y=z=R and the corner (R^2 + 3R + 1 samples)
    const uint32_t i = (blockIdx.x - core_blocks) * kK1Threads + threadIdx.x;
    uint32_t x, y, z;
    if (core_blocks == 0u) {
        if (i >= n * n * n) return;
        decode_sample(i, R, lg, x, y, z);
    } else {
        const uint32_t R2 = R << lg;
        if (i < R2) { x = i >> lg; y = i & (R - 1u); z = R; }
        else {
            const uint32_t e = (i - R2) >> lg, t = (i - R2) & (R - 1u);
            if (e == 0u)      { x = R; y = R; z = t; }
            else if (e == 1u) { x = R; y = t; z = R; }
            else if (e == 2u) { x = t; y = R; z = R; }
            else if (e == 3u && t == 0u) { x = R; y = R; z = R; }
            else return;
        }
    }
    const float px = __fadd_rn(g.s[0], __fmul_rn(g.across[0], __fmul_rn((float)x, inv_r)));
    const float py = __fadd_rn(g.s[1], __fmul_rn(g.across[1], __fmul_rn((float)y, inv_r)));
    const float pz = __fadd_rn(g.s[2], __fmul_rn(g.across[2], __fmul_rn((float)z, inv_r)));
    const float d = shape_de<kFast, kVariant>(sh, px, py, pz);      // fast mode: suspects re-evaluated in place
    const uint32_t j = (x * n + y) * n + z;
    grid[j] = d;
    if (sign_stride != 0u && (__float_as_uint(d) >> 31)) atomicOr(&plane[j >> 5], 1u << (j & 31u));
}

// Exact re-evaluation of the suspects K1 queued (fast mode): overwrites the sample and, where the
// sign changes, flips its bit of the plane.  Persistent grid-stride loop (the count lives on the device).
// The entry carries the fast evaluation's sign, so the kernel only WRITES the grid (a scattered read of
// the 4-byte sample, one DRAM sector and often a TLB miss per suspect, was 20x the cost of the arithmetic).
template <int kVariant>
__global__ void __launch_bounds__(kThreads)
fixup_suspects_kernel(ShapeDev sh, const SpanGeom* __restrict__ geom, uint32_t R, float inv_r,
                      float* __restrict__ grids, size_t grid_stride, uint32_t* __restrict__ sign_bits, uint32_t sign_stride,
                      SuspectList sl, MeshState* __restrict__ st /* may be NULL */) {
    const uint32_t n = R + 1u;
    const uint32_t total = *sl.count;
    const uint32_t cnt = min(total, sl.cap);
    uint32_t flips = 0;
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < cnt; i += gridDim.x * kThreads) {
        const uint2 e = sl.entries[i];
        const uint32_t j = e.y & 0x7fffffffu, was = e.y >> 31;
        const uint32_t z = j % n, xy = j / n, y = xy % n, x = xy / n;
        const SpanGeom g = geom[e.x];
        const float px = __fadd_rn(g.s[0], __fmul_rn(g.across[0], __fmul_rn((float)x, inv_r)));
        const float py = __fadd_rn(g.s[1], __fmul_rn(g.across[1], __fmul_rn((float)y, inv_r)));
        const float pz = __fadd_rn(g.s[2], __fmul_rn(g.across[2], __fmul_rn((float)z, inv_r)));
        const float d = mandelbulb_de_exact<kVariant == kVarP8>(sh, px, py, pz);
        grids[(size_t)e.x * grid_stride + j] = d;
        const uint32_t is = __float_as_uint(d) >> 31;
        if (was != is) {
            ++flips;
            if (sign_stride != 0u) atomicXor(&sign_bits[(size_t)e.x * sign_stride + (j >> 5)], 1u << (j & 31u));
        }
    }
    if (st) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) flips += __shfl_xor_sync(0xffffffffu, flips, o);
        if ((threadIdx.x & 31u) == 0u && flips) atomicAdd(&st->sign_fixups, (unsigned long long)flips);
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&st->suspects, (unsigned long long)total);
    }
}

// ---------------------------------------------------------------------------
// K3: point-list DE (12-byte packed Point3<f32>)
// ---------------------------------------------------------------------------
template <bool kFast, int kVariant>
__global__ void __launch_bounds__(kThreads)
de_batch_kernel(ShapeDev sh, const float* __restrict__ xyz, size_t n, float* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    const float px = xyz[3 * i], py = xyz[3 * i + 1], pz = xyz[3 * i + 2];
    out[i] = shape_de<kFast, kVariant>(sh, px, py, pz);
}

// ---------------------------------------------------------------------------
// block-wide exclusive scan of (v,q) pairs; returns block totals through tot_*
// ---------------------------------------------------------------------------
template <int kBlock>
__device__ __forceinline__ void block_exclusive_scan2(uint32_t v, uint32_t q, uint32_t& ev, uint32_t& eq,
                                                      uint32_t& tot_v, uint32_t& tot_q) {
    __shared__ uint32_t sv[kBlock / 32], sq[kBlock / 32];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t iv = v, iq = q;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t tv = __shfl_up_sync(0xffffffffu, iv, o);
        const uint32_t tq = __shfl_up_sync(0xffffffffu, iq, o);
        if (lane >= (uint32_t)o) { iv += tv; iq += tq; }
    }
    if (lane == 31u) { sv[warp] = iv; sq[warp] = iq; }
    __syncthreads();
    uint32_t bv = 0, bq = 0, tv = 0, tq = 0;
#pragma unroll
    for (int w = 0; w < kBlock / 32; ++w) {
        const uint32_t a = sv[w], b = sq[w];
        if ((uint32_t)w < warp) { bv += a; bq += b; }
        tv += a; tq += b;
    }
    ev = bv + iv - v; eq = bq + iq - q;
    tot_v = tv; tot_q = tq;
    __syncthreads();
}

// ---------------------------------------------------------------------------
// E1 / E2: classification from the sign bit-plane and order-preserving compaction.
//
// A word is 32 cells in cube(R) order; for R >= 32 that is 32 z-consecutive cells of one (x,y) row: its 8
// corner sign words are 33-bit windows of four plane rows (funnel shifts), and the active / edge masks are a
// handful of bitwise ops (buffer.rs:116-147, 299-350).  Words whose plane words are all 0 or all 1 (most of
// a volume) take a short cut.  A chunk is <= 256 consecutive words of one span, one CTA.
//
// Three short, wide kernels (no CTA ever waits for another):
//   E1  classify_count_kernel  (vertices, quads) per chunk, summed per span with one 64-bit atomic per chunk
//   E2a span_scan_kernel       prefix over the group's spans (hundreds), offset tables, totals
//   E2b emit_lists_kernel      chunks WITH active cells (~1/4 of a typical volume) recompute their masks and
//                              write the active-cell list, the quad list and the rank-query tables in place
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t plane_bit(const uint32_t* __restrict__ plane, uint32_t j) {
    return (plane[j >> 5] >> (j & 31u)) & 1u;
}

// masks of word `word` (cells 32 word .. 32 word + 31) of a span
__device__ __forceinline__ void word_masks(const uint32_t* __restrict__ plane, uint32_t R, uint32_t lg, uint32_t word,
                                           uint32_t& ma, uint32_t& mx, uint32_t& my, uint32_t& mz) {
    const uint32_t n = R + 1u;
    const uint32_t c0 = word << 5;
    ma = 0u; mx = 0u; my = 0u; mz = 0u;
    if (lg >= 5) {
        const uint32_t x = c0 >> (2 * lg), y = (c0 >> lg) & (R - 1u), z0 = c0 & (R - 1u);
        // the four plane rows of the word's corners: (x, y), (x, y+1), (x+1, y), (x+1, y+1); 33 bits from j each
        const uint32_t j0 = (x * n + y) * n + z0, j2 = j0 + n, j4 = j0 + n * n, j6 = j4 + n;
        const uint32_t a0 = plane[j0 >> 5], b0 = plane[(j0 >> 5) + 1u], a2 = plane[j2 >> 5], b2 = plane[(j2 >> 5) + 1u];
        const uint32_t a4 = plane[j4 >> 5], b4 = plane[(j4 >> 5) + 1u], a6 = plane[j6 >> 5], b6 = plane[(j6 >> 5) + 1u];
        const uint32_t any_w = a0 | b0 | a2 | b2 | a4 | b4 | a6 | b6, all_w = a0 & b0 & a2 & b2 & a4 & b4 & a6 & b6;
        if (any_w == 0u || all_w == 0xffffffffu) return;      // uniform plane words: no sign change in the word
        const uint32_t s0 = __funnelshift_r(a0, b0, j0 & 31u), h0 = (b0 >> (j0 & 31u)) & 1u;     // corner 0 (x, y, z)
        const uint32_t s2 = __funnelshift_r(a2, b2, j2 & 31u), h2 = (b2 >> (j2 & 31u)) & 1u;     // corner 2 (x, y+1, z)
        const uint32_t s4 = __funnelshift_r(a4, b4, j4 & 31u), h4 = (b4 >> (j4 & 31u)) & 1u;     // corner 4 (x+1, y, z)
        const uint32_t s6 = __funnelshift_r(a6, b6, j6 & 31u), h6 = (b6 >> (j6 & 31u)) & 1u;     // corner 6 (x+1, y+1, z)
        const uint32_t s1 = (s0 >> 1) | (h0 << 31), s3 = (s2 >> 1) | (h2 << 31);   // dz = 1 corners
        const uint32_t s5 = (s4 >> 1) | (h4 << 31), s7 = (s6 >> 1) | (h6 << 31);
        const uint32_t any = s0 | s1 | s2 | s3 | s4 | s5 | s6 | s7;
        const uint32_t all = s0 & s1 & s2 & s3 & s4 & s5 & s6 & s7;
        ma = any & ~all;
        const uint32_t zm = z0 == 0u ? ~1u : ~0u;                       // z > 0
        mx = (y > 0u) ? ((s0 ^ s4) & zm) : 0u;                          // y > 0 && z > 0 (:302)
        my = (x > 0u) ? ((s0 ^ s2) & zm) : 0u;                          // x > 0 && z > 0 (:326)
        mz = (x > 0u && y > 0u) ? (s0 ^ s1) : 0u;                       // x > 0 && y > 0 (:350)
    } else {
        const uint32_t R3 = R << (2 * lg);
        for (uint32_t b = 0; b < 32u; ++b) {
            const uint32_t c = c0 + b;
            if (c >= R3) break;
            const uint32_t x = c >> (2 * lg), y = (c >> lg) & (R - 1u), z = c & (R - 1u);
            const uint32_t j = (x * n + y) * n + z;
            const uint32_t s0 = plane_bit(plane, j), s1 = plane_bit(plane, j + 1u);
            const uint32_t s2 = plane_bit(plane, j + n), s3 = plane_bit(plane, j + n + 1u);
            const uint32_t s4 = plane_bit(plane, j + n * n), s5 = plane_bit(plane, j + n * n + 1u);
            const uint32_t s6 = plane_bit(plane, j + n * n + n), s7 = plane_bit(plane, j + n * n + n + 1u);
            const uint32_t sum = s0 + s1 + s2 + s3 + s4 + s5 + s6 + s7;
            ma |= (uint32_t)(sum != 0u && sum != 8u) << b;
            mx |= (uint32_t)(y > 0u && z > 0u && s0 != s4) << b;
            my |= (uint32_t)(x > 0u && z > 0u && s0 != s2) << b;
            mz |= (uint32_t)(x > 0u && y > 0u && s0 != s1) << b;
        }
    }
}

struct ExtractionLists {
    uint32_t* active;      // cell crosses the surface (buffer.rs:130-141), one bit per cell
    uint32_t* word_vpre;   // group-local vertex slot of the first active cell at/after each word
    uint32_t* cell_of;     // [cell_cap] span << 3 lg | cell of every active cell, cube(R) order
    uint32_t* quad_of;     // [quad_cap] (span << 3 lg | owner cell) | edge << 30, emission order
    uint32_t* span_first;  // [spans of the group] group-local slot of the span's first vertex
    uint8_t*  neg8;        // [cell_cap] written by E3: dists[lower corner of the cell] < 0.0 (winding, buffer.rs:310)
    uint32_t cell_cap, quad_cap;
};

// E1: counts.  gridDim = (chunks_per_span, spans); one thread per word.  Writes (active cells, quads) of every
// chunk and adds them to the span's total ({vertices:32 | quads:32} in one 64-bit atomic; zeroed per group).
__global__ void __launch_bounds__(kThreads)
classify_count_kernel(const uint32_t* __restrict__ sign_bits, uint32_t sign_stride, uint32_t R, uint32_t lg,
                      uint32_t chunk_words, uint2* __restrict__ chunk_cnt, unsigned long long* __restrict__ span_tot) {
    const uint32_t span = blockIdx.y, chunk = blockIdx.x, t = threadIdx.x;
    uint32_t ma = 0, mx = 0, my = 0, mz = 0;
    if (t < chunk_words) word_masks(sign_bits + (size_t)span * sign_stride, R, lg, chunk * chunk_words + t, ma, mx, my, mz);
    uint2* cnt = chunk_cnt + (size_t)span * gridDim.x + chunk;
    // a chunk without an active cell (most of them) has no quads either: every quad-owning corner is an active cell
    if (!__syncthreads_or(ma != 0u)) {
        if (t == 0) *cnt = make_uint2(0u, 0u);
        return;
    }
    __shared__ uint32_t s_v, s_q;
    if (t == 0) { s_v = 0u; s_q = 0u; }
    __syncthreads();
    const uint32_t v = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(ma));
    const uint32_t q = __reduce_add_sync(0xffffffffu, (uint32_t)(__popc(mx) + __popc(my) + __popc(mz)));
    if ((t & 31u) == 0u && v) { atomicAdd(&s_v, v); atomicAdd(&s_q, q); }
    __syncthreads();
    if (t == 0) {
        *cnt = make_uint2(s_v, s_q);
        atomicAdd(span_tot + span, ((unsigned long long)s_v << 32) | (unsigned long long)s_q);
    }
}

// E2a: exclusive scan over the group's SPAN totals (one CTA; a group has a few hundred spans), span offset
// tables, capacity check, running totals.
constexpr int kScanThreads = 1024;

__global__ void __launch_bounds__(kScanThreads)
span_scan_kernel(const unsigned long long* __restrict__ span_tot, uint2* __restrict__ span_pre, uint32_t* __restrict__ span_first,
                 uint32_t span0, uint32_t nspans_group,
                 unsigned long long* __restrict__ v_off, unsigned long long* __restrict__ i_off,
                 unsigned long long vcap, unsigned long long icap, MeshState* __restrict__ st,
                 volatile unsigned long long* __restrict__ progress /* mapped pinned host memory or NULL */) {
    const unsigned long long base_v = st->total_v, base_q = st->total_q;
    uint32_t carry_v = 0, carry_q = 0;
    for (uint32_t s0 = 0; s0 < nspans_group; s0 += kScanThreads) {
        const uint32_t s = s0 + threadIdx.x;
        const unsigned long long tot = s < nspans_group ? span_tot[s] : 0ull;
        const uint32_t v = (uint32_t)(tot >> 32), q = (uint32_t)tot;
        uint32_t ev, eq, tv, tq;
        block_exclusive_scan2<kScanThreads>(v, q, ev, eq, tv, tq);
        if (s < nspans_group) {
            const uint32_t pv = carry_v + ev, pq = carry_q + eq;
            span_pre[s] = make_uint2(pv, pq);
            span_first[s] = pv;
            v_off[span0 + s] = base_v + pv;
            i_off[span0 + s] = 6ull * (base_q + pq);
        }
        carry_v += tv; carry_q += tq;
    }
    if (threadIdx.x == 0) {
        const unsigned long long nv = base_v + carry_v, nq = base_q + carry_q;
        v_off[span0 + nspans_group] = nv;
        i_off[span0 + nspans_group] = 6ull * nq;
        st->group_base_v = base_v; st->group_base_q = base_q;
        st->group_v = carry_v; st->group_q = carry_q;
        st->total_v = nv; st->total_q = nq;
        if (nv > vcap || 6ull * nq > icap) st->overflow = 1u;
        // totals after this group, for the host's pipelined device->host copies
        if (progress) { progress[0] = nv; progress[1] = nq; }
    }
}

// E2b: the chunks with active cells write everything at its final place: `active` masks + `word_vpre`
// (GROUP-local vertex slot at each word) for the rank queries of E4, `cell_of` (compacted active cells, the
// work list of E3), `quad_of` (one entry per quad in the reference's emission order, the work list of E4).
// A chunk's prefix = its span's prefix + the counts of the span's preceding chunks (<= 31 of them at
// R = 64): no chain between CTAs.  The masks are recomputed from the plane (L2 hits) rather than stored by E1.
__global__ void __launch_bounds__(kThreads)
emit_lists_kernel(const uint32_t* __restrict__ sign_bits, uint32_t sign_stride, uint32_t R, uint32_t lg,
                  uint32_t words_per_span, uint32_t chunk_words, const uint2* __restrict__ chunk_cnt,
                  const uint2* __restrict__ span_pre, ExtractionLists ls) {
    const uint32_t span = blockIdx.y, chunk = blockIdx.x, t = threadIdx.x;
    const uint2* cnt = chunk_cnt + (size_t)span * gridDim.x;
    if (cnt[chunk].x == 0u) return;        // uniform: nothing of this chunk is ever read
    __shared__ uint32_t s_pre[2];
    if (t < 32u) {
        uint32_t pv = 0, pq = 0;
        for (uint32_t c = t; c < chunk; c += 32u) { const uint2 k = cnt[c]; pv += k.x; pq += k.y; }
        pv = __reduce_add_sync(0xffffffffu, pv); pq = __reduce_add_sync(0xffffffffu, pq);
        if (t == 0) { const uint2 sp = span_pre[span]; s_pre[0] = sp.x + pv; s_pre[1] = sp.y + pq; }
    }
    uint32_t ma = 0, mx = 0, my = 0, mz = 0;
    if (t < chunk_words) word_masks(sign_bits + (size_t)span * sign_stride, R, lg, chunk * chunk_words + t, ma, mx, my, mz);
    const uint32_t v = __popc(ma), q = __popc(mx) + __popc(my) + __popc(mz);
    uint32_t ev, eq, tv, tq;
    block_exclusive_scan2<kThreads>(v, q, ev, eq, tv, tq);       // (its barriers also publish s_pre)
    if (t >= chunk_words) return;
    const size_t o = (size_t)span * words_per_span + chunk * chunk_words + t;
    uint32_t gv = s_pre[0] + ev, slot = s_pre[1] + eq;
    ls.active[o] = ma;
    ls.word_vpre[o] = gv;
    const uint32_t cell0 = (span << (3 * lg)) + ((chunk * chunk_words + t) << 5);
    for (uint32_t m = ma; m; m &= m - 1u, ++gv)
        if (gv < ls.cell_cap) ls.cell_of[gv] = cell0 + (uint32_t)__ffs((int)m) - 1u;
    // the quads owned by the word's cells, corner by corner: +x, +y, +z edge (buffer.rs:299-371)
    for (uint32_t m = mx | my | mz; m; m &= m - 1u) {
        const uint32_t b = (uint32_t)__ffs((int)m) - 1u, cell = cell0 + b;
        if ((mx >> b) & 1u) { if (slot < ls.quad_cap) ls.quad_of[slot] = cell; ++slot; }
        if ((my >> b) & 1u) { if (slot < ls.quad_cap) ls.quad_of[slot] = cell | (1u << 30); ++slot; }
        if ((mz >> b) & 1u) { if (slot < ls.quad_cap) ls.quad_of[slot] = cell | (2u << 30); ++slot; }
    }
}

// ---------------------------------------------------------------------------
// E3: vertices.  Persistent grid-stride loop over the group's compacted active
// cells (the count lives on the device, no host sync).  One thread per vertex.
// ---------------------------------------------------------------------------
template <bool kFast, int kVariant>
__global__ void __launch_bounds__(kE3Threads, CTC_E3_MINBLOCKS)
vertex_kernel(ShapeDev sh, const SpanGeom* __restrict__ geom, const float* __restrict__ grids, size_t grid_stride,
              uint32_t R, uint32_t lg, const uint32_t* __restrict__ cell_of, uint32_t cell_cap, uint8_t* __restrict__ neg8,
              MeshState* st, uint32_t span0, float* __restrict__ out_v, unsigned long long vcap) {
    using M = MathExact;
    const uint32_t nv = min(st->group_v, cell_cap);
    const unsigned long long base_v = st->group_base_v;
    const uint32_t n = R + 1u;
    const uint32_t lg3 = 3 * lg;
    for (uint32_t v = blockIdx.x * kE3Threads + threadIdx.x; v < nv; v += gridDim.x * kE3Threads) {
        const unsigned long long slot = base_v + v;
        if (slot >= vcap) continue;
        const uint32_t cell = cell_of[v];
        const uint32_t span = cell >> lg3, c = cell & ((1u << lg3) - 1u);
        const uint32_t x = c >> (2 * lg), y = (c >> lg) & (R - 1u), z = c & (R - 1u);
        const SpanGeom g = geom[span];   // geom is pre-offset to the group
        const float* p = grids + (size_t)span * grid_stride + ((size_t)x * n + y) * n + z;
        const size_t sy = n, sx = (size_t)n * n;
        float dist[8];
        dist[0] = p[0]; dist[1] = p[1]; dist[2] = p[sy]; dist[3] = p[sy + 1];
        dist[4] = p[sx]; dist[5] = p[sx + 1]; dist[6] = p[sx + sy]; dist[7] = p[sx + sy + 1];
        // winding of the quads this cell owns uses `dists[(x,y,z)] < 0.0`, not the sign bit (buffer.rs:310:
        // -0.0 and NaN differ); E4 picks it up by vertex slot instead of re-reading the grid
        neg8[v] = dist[0] < 0.0f ? 1 : 0;

        // p0 = span.start + (x,y,z) * step  (buffer.rs:150-151)
        const float p0x = M::add(g.s[0], M::mul((float)x, g.step[0]));
        const float p0y = M::add(g.s[1], M::mul((float)y, g.step[1]));
        const float p0z = M::add(g.s[2], M::mul((float)z, g.step[2]));

        // 12 edges (buffer.rs:155-176): from/to corner ids packed 4 bits each
        // x edges (0,4)(1,5)(2,6)(3,7); y edges (0,2)(1,3)(4,6)(5,7); z edges (0,1)(2,3)(4,5)(6,7)
        const unsigned long long kFrom = 0x642054103210ull, kTo = 0x753176327654ull;
        float qx, qy, qz;
        bool bad = false;
        if (kFast) {
            // FAST: the centroid of the edge crossings in cell-local coordinates, branch-free.  The
            // reference's (mirrored) weight (d_from + D) / D, D = d_to - d_from, is d_to / (d_to - d_from)
            // whichever way the edge is oriented, and a crossing on an x edge sits at local
            // (w, y_from, z_from): one reciprocal and a handful of selects per edge instead of twelve
            // divergent regions with an IEEE division each (11 of 32 lanes active in round 1's profile).
            float lx = 0.0f, ly = 0.0f, lz = 0.0f, cnt = 0.0f;
#pragma unroll
            for (int e = 0; e < 12; ++e) {
                const int from = (int)((kFrom >> (4 * e)) & 15ull), to = (int)((kTo >> (4 * e)) & 15ull);
                const float df = dist[from], dt = dist[to];
                const bool cross = ((__float_as_uint(df) ^ __float_as_uint(dt)) >> 31) != 0u;   // :194-196
                const float den = dt - df;
                const float w = den == 0.0f ? 0.5f : dt * fast_rcp(den);                         // :217-239
                bad |= cross && !(w >= 0.0f && w <= 1.0f);                                       // math.rs:19
                const float c = cross ? 1.0f : 0.0f, cw = cross ? w : 0.0f;
                const int axis = (from ^ to);          // 4: x edge, 2: y edge, 1: z edge
                lx += axis == 4 ? cw : ((from & 4) ? c : 0.0f);
                ly += axis == 2 ? cw : ((from & 2) ? c : 0.0f);
                lz += axis == 1 ? cw : ((from & 1) ? c : 0.0f);
                cnt += c;
            }
            const float inv = fast_rcp(cnt);
            qx = fmaf(lx * inv, g.step[0], p0x);
            qy = fmaf(ly * inv, g.step[1], p0y);
            qz = fmaf(lz * inv, g.step[2], p0z);
        } else {
            // corner coordinates p0 + corner_offsets[i] (buffer.rs:102-111, 242): the offset is 0.0 or step per axis
            const float cx0 = M::add(p0x, 0.0f), cx1 = M::add(p0x, g.step[0]);
            const float cy0 = M::add(p0y, 0.0f), cy1 = M::add(p0y, g.step[1]);
            const float cz0 = M::add(p0z, 0.0f), cz1 = M::add(p0z, g.step[2]);
            int count = 0;
            float sx_ = 0.0f, sy_ = 0.0f, sz_ = 0.0f;
#pragma unroll
            for (int e = 0; e < 12; ++e) {
                const int from = (int)((kFrom >> (4 * e)) & 15ull), to = (int)((kTo >> (4 * e)) & 15ull);
                const float df = dist[from], dt = dist[to];
                if ((__float_as_uint(df) >> 31) == (__float_as_uint(dt) >> 31)) continue;   // :194-196
                float d_from, d_to;
                if (df < 0.0f) { d_from = df; d_to = dt; } else { d_from = -df; d_to = -dt; }   // :209-213
                float w;
                if (d_to == d_from) w = 0.5f;                                                  // :217-218
                else { const float dl = M::sub(d_to, d_from); w = M::div(M::add(d_from, dl), dl); }   // :238-239
                if (!(w >= 0.0f && w <= 1.0f)) bad = true;                                     // math.rs:19
                const float om = M::sub(1.0f, w);
                // lerp(p0 + off[from], p0 + off[to], w) = a*(1-w) + b*w   (math.rs:45-48)
                const float ax = (from & 4) ? cx1 : cx0, bx = (to & 4) ? cx1 : cx0;
                const float ay = (from & 2) ? cy1 : cy0, by = (to & 2) ? cy1 : cy0;
                const float az = (from & 1) ? cz1 : cz0, bz = (to & 1) ? cz1 : cz0;
                sx_ = M::add(sx_, M::add(M::mul(ax, om), M::mul(bx, w)));
                sy_ = M::add(sy_, M::add(M::mul(ay, om), M::mul(by, w)));
                sz_ = M::add(sz_, M::add(M::mul(az, om), M::mul(bz, w)));
                ++count;
            }
            // centroid (buffer.rs:247-250): origin + sum / count
            const float fc = (float)count;
            qx = M::add(0.0f, M::div(sx_, fc));
            qy = M::add(0.0f, M::div(sy_, fc));
            qz = M::add(0.0f, M::div(sz_, fc));
        }
        if (bad) atomicMin(&st->panic_span, span0 + span);

        // dist_p and the un-normalised central differences (buffer.rs:254-265).
        // unit_x() * d = (1*d, 0*d, 0*d): the zero products keep their sign.  The reference multiplies
        // the unit vector by +-delta.<axis>, so the two off-axis components are p + 0 * (+-delta.<axis>).
        float ex[7], ey[7], ez[7], de[7];
        ex[0] = qx; ey[0] = qy; ez[0] = qz;
#pragma unroll
        for (int k = 1; k < 7; ++k) {
            const int axis = (k - 1) >> 1;
            const float sg = (k & 1) ? 1.0f : -1.0f;       // odd k: +delta, even k: -delta
            const float da = sg * (axis == 0 ? g.delta[0] : axis == 1 ? g.delta[1] : g.delta[2]);
            ex[k] = M::add(qx, M::mul(axis == 0 ? 1.0f : 0.0f, da));
            ey[k] = M::add(qy, M::mul(axis == 1 ? 1.0f : 0.0f, da));
            ez[k] = M::add(qz, M::mul(axis == 2 ? 1.0f : 0.0f, da));
        }
        if (kFast && kVariant == kVarP8) {
            // the +-delta evaluations of an axis go through the packed FP32 path as one pair; values only
            // (normals, distance_from_surface), so only the z-axis / NaN rule sends a point to the exact path
            de[0] = shape_de<true, kVarP8, false>(sh, ex[0], ey[0], ez[0]);
#pragma unroll
            for (int k = 1; k < 7; k += 2) {
                uint32_t susp;
                const float2 d2 = mandelbulb_de_fast_p8_pair<false>(sh, make_float2(ex[k], ex[k + 1]), make_float2(ey[k], ey[k + 1]),
                                                                    make_float2(ez[k], ez[k + 1]), susp);
                de[k] = d2.x; de[k + 1] = d2.y;
                if (susp & 1u) de[k] = mandelbulb_de_exact_cold<true>(sh, ex[k], ey[k], ez[k]);
                if (susp & 2u) de[k + 1] = mandelbulb_de_exact_cold<true>(sh, ex[k + 1], ey[k + 1], ez[k + 1]);
            }
        } else if (kFast && kVariant == kVarGeneric && pow2_path(sh.power) != 0) {
            // P = 2, 4, 16: the +-delta evaluations as packed pairs, like the power-8 branch
            de[0] = shape_de<true, kVarGeneric, false>(sh, ex[0], ey[0], ez[0]);
            const int lp = pow2_path(sh.power);
#pragma unroll
            for (int k = 1; k < 7; k += 2) {
                uint32_t susp;
                const float2 ax = make_float2(ex[k], ex[k + 1]), ay = make_float2(ey[k], ey[k + 1]), az = make_float2(ez[k], ez[k + 1]);
                const float2 d2 = lp == 1 ? mandelbulb_de_fast_pow2_pair<1>(sh, ax, ay, az, susp)
                                : lp == 2 ? mandelbulb_de_fast_pow2_pair<2>(sh, ax, ay, az, susp)
                                          : mandelbulb_de_fast_pow2_pair<4>(sh, ax, ay, az, susp);
                de[k] = d2.x; de[k + 1] = d2.y;
                if (susp & 1u) de[k] = mandelbulb_de_exact_cold<false>(sh, ex[k], ey[k], ez[k]);
                if (susp & 2u) de[k + 1] = mandelbulb_de_exact_cold<false>(sh, ex[k + 1], ey[k + 1], ez[k + 1]);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 7; ++k) de[k] = shape_de<kFast, kVariant, false>(sh, ex[k], ey[k], ez[k]);
        }
        const float nx = M::sub(de[1], de[2]), ny = M::sub(de[3], de[4]), nz = M::sub(de[5], de[6]);
        // cgmath normalize: v * (1 / sqrt((x*x + y*y) + z*z))
        float inv;
        if (kFast) inv = fast_rsqrt(fmaf(nz, nz, fmaf(ny, ny, nx * nx)));
        // (a squared length in the flush-to-zero range makes the approximation inf: take the IEEE route then)
        if (!kFast || !(inv <= 1e18f))
            inv = M::div(1.0f, M::sqrt(M::add(M::add(M::mul(nx, nx), M::mul(ny, ny)), M::mul(nz, nz))));
        float* o = out_v + slot * 7ull;
        o[0] = qx; o[1] = qy; o[2] = qz;
        o[3] = M::mul(nx, inv); o[4] = M::mul(ny, inv); o[5] = M::mul(nz, inv);
        o[6] = de[0];
    }
}

// ---------------------------------------------------------------------------
// E4: quads.
// ---------------------------------------------------------------------------
// group-local vertex slot of active cell c of a span (rank query on the `active` masks)
__device__ __forceinline__ uint32_t vertex_slot(const uint32_t* __restrict__ active, const uint32_t* __restrict__ word_vpre,
                                                size_t span_w0, uint32_t c) {
    const size_t o = span_w0 + (c >> 5);
    return word_vpre[o] + __popc(active[o] & ((1u << (c & 31u)) - 1u));
}

// Packed wire record of one quad (8 bytes instead of 24): four 16-bit span-local vertex ids.  The four
// cells around an edge are always in increasing cube(R) order, so v0 < v1 < v2 < v3, and the winding
// flag rides on the ORDER of the first two: (v0, v1, v2, v3) = keep, (v1, v0, v2, v3) = flip.
// expand_quads_kernel turns a record back into the six u32 indices.
template <bool kPacked>
__device__ __forceinline__ void store_quad(uint32_t* __restrict__ out_idx, unsigned long long q, unsigned long long icap,
                                           bool flip, uint32_t v0, uint32_t v1, uint32_t v2, uint32_t v3,
                                           unsigned int* __restrict__ wire_overflow) {
    if (6ull * q + 6ull > icap) return;
    if (kPacked) {
        if ((v0 | v1 | v2 | v3) >> 16) { *wire_overflow = 1u; return; }
        reinterpret_cast<uint2*>(out_idx)[q] = make_uint2(flip ? (v1 | (v0 << 16)) : (v0 | (v1 << 16)), v2 | (v3 << 16));
        return;
    }
    uint2* d = reinterpret_cast<uint2*>(out_idx + 6ull * q);
    if (flip) { d[0] = make_uint2(v0, v2); d[1] = make_uint2(v1, v1); d[2] = make_uint2(v2, v3); }   // [v0,v2,v1, v1,v2,v3]
    else      { d[0] = make_uint2(v0, v1); d[1] = make_uint2(v2, v1); d[2] = make_uint2(v3, v2); }   // [v0,v1,v2, v1,v3,v2]
}

// One thread per QUAD (the list E1+E2 wrote, already in the reference's emission order): every lane has
// exactly one record to produce, consecutive lanes write consecutive 24-byte records.  The quad of the
// sign-changing edge from corner c along +x / +y / +z joins the vertices of the four cells around that edge
// (buffer.rs:302-371): c - a - b, c - a, c - b, c with (a, b) = (R, 1), (R^2, 1), (R^2, R).  Persistent
// grid-stride loop (the list length lives on the device).
template <bool kPacked>
__global__ void __launch_bounds__(kThreads)
quad_kernel(ExtractionLists ls, uint32_t R, uint32_t lg, uint32_t words_per_span,
            MeshState* st, uint32_t* __restrict__ out_idx, unsigned long long icap) {
    const uint32_t nq = min(st->group_q, ls.quad_cap);
    const unsigned long long base_q = st->group_base_q;
    const uint32_t R2 = R << lg, lg3 = 3 * lg;
    for (uint32_t q = blockIdx.x * kThreads + threadIdx.x; q < nq; q += gridDim.x * kThreads) {
        const uint32_t rec = ls.quad_of[q];
        const uint32_t edge = rec >> 30, cell = rec & 0x3fffffffu;
        const uint32_t span = cell >> lg3, c = cell & ((1u << lg3) - 1u);
        const size_t w0 = (size_t)span * words_per_span;
        const uint32_t first = ls.span_first[span];
        const uint32_t a = edge == 0u ? R : R2, b = edge == 2u ? R : 1u;
        const uint32_t s3 = vertex_slot(ls.active, ls.word_vpre, w0, c);
        const uint32_t v0 = vertex_slot(ls.active, ls.word_vpre, w0, c - a - b) - first;
        const uint32_t v1 = vertex_slot(ls.active, ls.word_vpre, w0, c - a) - first;
        const uint32_t v2 = vertex_slot(ls.active, ls.word_vpre, w0, c - b) - first;
        const bool neg = s3 < ls.cell_cap && ls.neg8[s3] != 0;
        // the +y edge's winding is flipped relative to x / z (buffer.rs:326-347)
        store_quad<kPacked>(out_idx, base_q + q, icap, edge == 1u ? !neg : neg, v0, v1, v2, s3 - first, &st->wire_overflow);
    }
}

// Widens packed quad records (see store_quad) into the reference's six u32 indices per quad.
__global__ void __launch_bounds__(kThreads)
expand_quads_kernel(const uint2* __restrict__ rec, size_t nquads, uint32_t* __restrict__ out_idx) {
    for (size_t q = (size_t)blockIdx.x * kThreads + threadIdx.x; q < nquads; q += (size_t)gridDim.x * kThreads) {
        const uint2 r = rec[q];
        const uint32_t a = r.x & 0xFFFFu, b = r.x >> 16, v2 = r.y & 0xFFFFu, v3 = r.y >> 16;
        const bool flip = a > b;
        const uint32_t v0 = flip ? b : a, v1 = flip ? a : b;
        uint2* d = reinterpret_cast<uint2*>(out_idx + 6 * q);
        if (flip) { d[0] = make_uint2(v0, v2); d[1] = make_uint2(v1, v1); d[2] = make_uint2(v2, v3); }
        else           { d[0] = make_uint2(v0, v1); d[1] = make_uint2(v2, v1); d[2] = make_uint2(v3, v2); }
    }
}

// ---------------------------------------------------------------------------
// N1: sphere tracing of the focus rays, ShapeMesh::get_focii
// (mesh/mod.rs:229-241).  One thread per ray:
//     for _ in 0..MAX_ITERS { d = DE(pos); pos += dir * d; if d < EPSILON { return Some(pos) } }
// hit[i] = 1 and out[i] = pos on success, hit[i] = 0 otherwise.  Arithmetic of the march itself
// (pos += dir * d) is exact-order; the DE follows the shape's math mode.
// ---------------------------------------------------------------------------
template <bool kFast, int kVariant>
__global__ void __launch_bounds__(kThreads)
ray_march_kernel(ShapeDev sh, const float* __restrict__ origin, const float* __restrict__ dir, size_t n,
                 uint32_t max_steps, float epsilon, float* __restrict__ out, uint32_t* __restrict__ hit) {
    const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    float px = origin[3 * i], py = origin[3 * i + 1], pz = origin[3 * i + 2];
    const float dx = dir[3 * i], dy = dir[3 * i + 1], dz = dir[3 * i + 2];
    uint32_t ok = 0;
    for (uint32_t s = 0; s < max_steps; ++s) {
        const float d = shape_de<kFast, kVariant>(sh, px, py, pz);
        px = __fadd_rn(px, __fmul_rn(dx, d));
        py = __fadd_rn(py, __fmul_rn(dy, d));
        pz = __fadd_rn(pz, __fmul_rn(dz, d));
        if (d < epsilon) { ok = 1; break; }
    }
    out[3 * i] = px; out[3 * i + 1] = py; out[3 * i + 2] = pz;
    hit[i] = ok;
}

// ---------------------------------------------------------------------------
// N4: the planned ray-marcher (README.md:12-16) as the same sphere-tracing loop, one thread per PIXEL.
// Pixel (i, j) looks from `eye` through  top_left + (i + 0.5) du + (j + 0.5) dv ; the ray is marched exactly like
// a focus ray of get_focii (mesh/mod.rs:229-241).  out[pixel] = (pos.x, pos.y, pos.z, t): the final position and
// the distance travelled, t < 0 when the ray did not come within `epsilon` of the surface in max_steps steps.
// The ray set-up is explicit IEEE arithmetic (mul, add, sqrt, div: no contraction), so a host can reproduce the
// rays bit for bit; the DE follows the shape's math mode.
// ---------------------------------------------------------------------------
struct CameraRays { float eye[3], top_left[3], du[3], dv[3]; };

template <bool kFast, int kVariant>
__global__ void __launch_bounds__(kThreads)
render_kernel(ShapeDev sh, CameraRays cam, uint32_t width, uint32_t height, uint32_t max_steps, float epsilon,
              float4* __restrict__ out) {
    // 16 x 16 pixel tiles per CTA: neighbouring rays take similar step counts
    const uint32_t tx = threadIdx.x & 15u, ty = threadIdx.x >> 4;
    const uint32_t i = blockIdx.x * 16u + tx, j = blockIdx.y * 16u + ty;
    if (i >= width || j >= height) return;
    const float fi = __fadd_rn((float)i, 0.5f), fj = __fadd_rn((float)j, 0.5f);
    float d3[3], p3[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float q = __fadd_rn(__fadd_rn(cam.top_left[c], __fmul_rn(fi, cam.du[c])), __fmul_rn(fj, cam.dv[c]));
        d3[c] = __fsub_rn(q, cam.eye[c]);
        p3[c] = cam.eye[c];
    }
    // cgmath normalize: v * (1 / sqrt((x*x + y*y) + z*z))
    const float inv = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(d3[0], d3[0]), __fmul_rn(d3[1], d3[1])), __fmul_rn(d3[2], d3[2]))));
    const float dx = __fmul_rn(d3[0], inv), dy = __fmul_rn(d3[1], inv), dz = __fmul_rn(d3[2], inv);
    float px = p3[0], py = p3[1], pz = p3[2], t = 0.0f;
    bool hit = false;
    for (uint32_t s = 0; s < max_steps; ++s) {
        const float d = shape_de<kFast, kVariant>(sh, px, py, pz);
        px = __fadd_rn(px, __fmul_rn(dx, d));
        py = __fadd_rn(py, __fmul_rn(dy, d));
        pz = __fadd_rn(pz, __fmul_rn(dz, d));
        t = __fadd_rn(t, d);
        if (d < epsilon) { hit = true; break; }
        if (!(t < 1e6f)) break;          // the ray has left for good (also ends NaN marches)
    }
    out[(size_t)j * width + i] = make_float4(px, py, pz, hit ? t : -1.0f);
}

// ---------------------------------------------------------------------------
// Measurement aids (not on the hot path).
// ---------------------------------------------------------------------------
// Completed iterations and bail-outs over the sample lattices of a span batch (EXACT arithmetic,
// i.e. the reference's own counts): out[0] += sum k, out[1] += #bailed, out[2] += #samples.
// Feeds the algorithmic flop count flops(sample) = 75 k + 6 [bailed] + 10.
template <int kVariant>
__global__ void __launch_bounds__(kThreads)
iteration_stats_kernel(ShapeDev sh, const SpanGeom* __restrict__ geom, uint32_t R, uint32_t lg, float inv_r,
                       unsigned long long* __restrict__ out) {
    const uint32_t n = R + 1u, n3 = n * n * n;
    const uint32_t i = blockIdx.x * kThreads + threadIdx.x;
    uint32_t k = 0, bailed = 0, cnt = 0;
    if (i < n3) {
        uint32_t x, y, z;
        decode_sample(i, R, lg, x, y, z);
        const SpanGeom g = geom[blockIdx.y];
        const float px = __fadd_rn(g.s[0], __fmul_rn(g.across[0], __fmul_rn((float)x, inv_r)));
        const float py = __fadd_rn(g.s[1], __fmul_rn(g.across[1], __fmul_rn((float)y, inv_r)));
        const float pz = __fadd_rn(g.s[2], __fmul_rn(g.across[2], __fmul_rn((float)z, inv_r)));
        if (kVariant != kVarSphere) {
            (void)mandelbulb_de_exact<kVariant == kVarP8>(sh, px, py, pz, &k);
            bailed = k < sh.max_iters;
        }
        cnt = 1;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        k += __shfl_xor_sync(0xffffffffu, k, o);
        bailed += __shfl_xor_sync(0xffffffffu, bailed, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if ((threadIdx.x & 31u) == 0u) {
        atomicAdd(&out[0], (unsigned long long)k);
        atomicAdd(&out[1], (unsigned long long)bailed);
        atomicAdd(&out[2], (unsigned long long)cnt);
    }
}

// The same statistics for a point list (packed xyz): the algorithmic flops of pass 2's 7 DE evaluations
// per vertex.
template <int kVariant>
__global__ void __launch_bounds__(kThreads)
iteration_stats_points_kernel(ShapeDev sh, const float* __restrict__ xyz, size_t n, unsigned long long* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
    uint32_t k = 0, bailed = 0, cnt = 0;
    if (i < n) {
        if (kVariant != kVarSphere) {
            (void)mandelbulb_de_exact<kVariant == kVarP8>(sh, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], &k);
            bailed = k < sh.max_iters;
        }
        cnt = 1;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        k += __shfl_xor_sync(0xffffffffu, k, o);
        bailed += __shfl_xor_sync(0xffffffffu, bailed, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if ((threadIdx.x & 31u) == 0u) {
        atomicAdd(&out[0], (unsigned long long)k);
        atomicAdd(&out[1], (unsigned long long)bailed);
        atomicAdd(&out[2], (unsigned long long)cnt);
    }
}

// Calibration of the sign-trust band (fast_suspect_*, de_device.cuh): every sample of the span batch is
// evaluated by the fast path (raw, no repair) AND by the exact path.  out[] (unsigned long long):
//   [0] samples                      [1] raw sign mismatches fast vs exact
//   [2] mismatches with an iterate on the z axis (1/w >= 2^13)
//   [3] samples not escaped in both  [4] samples whose escape status differs
//   [5] max |r2_fast - r2_exact| / dr over [3], as float bits    [6] the same over the amplification bound 2^lemax
//   [7] mismatches where the fast path escaped
//   [8 + 4q ..] for q = 0..23, kappa = 2^-(8+q), axis rule applied first:
//        suspects / uncovered mismatches with the band kappa * dr (for reference), then with the production
//        band kappa * 2^lemax (amp_step, de_device.cuh)
//   [kProbeWords ..] up to kProbeDump records of 12 floats: mismatches the production band misses at kappa = 2^-17
constexpr int kProbeKappas = 24;
constexpr int kProbeWords = 8 + 4 * kProbeKappas;
constexpr int kProbeDump = 64;

__global__ void __launch_bounds__(kThreads)
fast_sign_probe_kernel(ShapeDev sh, const SpanGeom* __restrict__ geom, uint32_t R, uint32_t lg, float inv_r,
                       unsigned long long* __restrict__ out, float* __restrict__ dump, unsigned int* __restrict__ dump_count) {
    __shared__ unsigned int acc[kProbeWords];
    for (int i = threadIdx.x; i < kProbeWords; i += kThreads) acc[i] = 0u;
    __syncthreads();
    const uint32_t n = R + 1u, n3 = n * n * n;
    const uint32_t i = blockIdx.x * kThreads + threadIdx.x;
    const bool live = i < n3;
    bool mism = false, axis = false, esc_e = false, both_in = false;
    FastInfo fi{1.0f, 1.0f, 1.0f, 1u, 0u};
    float e_max = 0.0f, e_pol = 0.0f, px = 0.f, py = 0.f, pz = 0.f, re = 0.f, df = 0.f, dx = 0.f;
    if (live) {
        uint32_t x, y, z;
        decode_sample(i, R, lg, x, y, z);
        const SpanGeom g = geom[blockIdx.y];
        px = __fadd_rn(g.s[0], __fmul_rn(g.across[0], __fmul_rn((float)x, inv_r)));
        py = __fadd_rn(g.s[1], __fmul_rn(g.across[1], __fmul_rn((float)y, inv_r)));
        pz = __fadd_rn(g.s[2], __fmul_rn(g.across[2], __fmul_rn((float)z, inv_r)));
        bool susp_unused;
        df = mandelbulb_de_fast_p8<true>(sh, px, py, pz, susp_unused, &fi);
        uint32_t it;
        dx = mandelbulb_de_exact<true>(sh, px, py, pz, &it, &re);
        esc_e = it < sh.max_iters;
        mism = (__float_as_uint(df) >> 31) != (__float_as_uint(dx) >> 31);
        axis = fi.axis != 0u;
        both_in = !fi.escaped && !esc_e;
        if (both_in && !axis) {
            const float err = fabsf(fi.r2 - re * re);
            if (err == err) { e_max = err / fi.dr; e_pol = err / fi.amp; }
        }
    }
    const uint32_t lane = threadIdx.x & 31u;
#define CTC_PROBE_COUNT(K, COND)                                                         \
    { const uint32_t b_ = __ballot_sync(0xffffffffu, live && (COND));                    \
      if (lane == 0u && b_) atomicAdd(&acc[K], (unsigned int)__popc(b_)); }
    CTC_PROBE_COUNT(0, true)
    CTC_PROBE_COUNT(1, mism)
    CTC_PROBE_COUNT(2, mism && axis)
    CTC_PROBE_COUNT(3, both_in)
    CTC_PROBE_COUNT(4, (fi.escaped != 0u) != esc_e)
    CTC_PROBE_COUNT(7, mism && fi.escaped)
    atomicMax(&acc[5], __float_as_uint(e_max));
    atomicMax(&acc[6], __float_as_uint(e_pol));
    float kappa = 1.0f / 256.0f;
    for (int q = 0; q < kProbeKappas; ++q, kappa *= 0.5f) {
        bool s_max, s_pol;
        if (fi.escaped) { s_max = axis || !(kappa * fi.dr < 4.0f); s_pol = axis || !(kappa * fi.amp < 4.0f); }
        else {
            s_max = axis || !(fabsf(fi.r2 - 1.0f) > kappa * fi.dr);
            s_pol = axis || !(fabsf(fi.r2 - 1.0f) > kappa * fi.amp);
        }
        CTC_PROBE_COUNT(8 + 4 * q, s_max)
        CTC_PROBE_COUNT(9 + 4 * q, mism && !s_max)
        CTC_PROBE_COUNT(10 + 4 * q, s_pol)
        CTC_PROBE_COUNT(11 + 4 * q, mism && !s_pol)
        if (q == 9 && live && mism && !s_pol) {
            const unsigned int slot = atomicAdd(dump_count, 1u);
            if (slot < (unsigned)kProbeDump) {
                float* o = dump + 12 * slot;
                o[0] = px; o[1] = py; o[2] = pz; o[3] = fi.r2; o[4] = re * re; o[5] = fi.dr; o[6] = fi.amp;
                o[7] = (float)fi.escaped; o[8] = esc_e ? 1.0f : 0.0f; o[9] = 0.0f; o[10] = df; o[11] = dx;
            }
        }
    }
#undef CTC_PROBE_COUNT
    __syncthreads();
    for (int k = threadIdx.x; k < kProbeWords; k += kThreads) {
        const unsigned long long v = acc[k];
        if (v == 0ull) continue;
        if (k == 5 || k == 6) atomicMax(&out[k], v); else atomicAdd(&out[k], v);
    }
}

// Dependent-free FFMA streams: the sustained FP32 FMA rate of this GPU at its running clock
// (the denominator SURVEY.md 8d asks to be reported beside the nominal SMs*128*2*clock).
__global__ void __launch_bounds__(kThreads)
fma_peak_kernel(float* __restrict__ out, uint32_t iters) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
    float a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float m = 0.999f, c = 1e-3f;
    for (uint32_t i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
            a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
        }
    }
    out[blockIdx.x * kThreads + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

// Resets the call state at the start of a mesh call.
__global__ void reset_state_kernel(MeshState* st) {
    st->total_v = 0; st->total_q = 0; st->group_base_v = 0; st->group_base_q = 0;
    st->group_v = 0; st->group_q = 0; st->overflow = 0; st->panic_span = 0xFFFFFFFFu;
    st->wire_overflow = 0; st->pad_ = 0; st->suspects = 0; st->sign_fixups = 0;
}

}  // namespace ctc
