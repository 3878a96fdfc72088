// kernels.cuh -- the hot-path kernels (sm_100a).
//
//   K1  sample_grids_kernel   pass 1 of naive_surface_nets (mesh/buffer.rs:77-83)
//   K3  de_batch_kernel       Shape::batch_min_distance_from (shape/mod.rs:89)
//   E1  classify_kernel       cell / edge sign classification (buffer.rs:116-147, 299-350)
//   E2a scan_chunks_kernel    order-preserving prefix over chunk counts
//   E2b apply_prefix_kernel   per-word prefixes + compacted active-cell list
//   E3  vertex_kernel         per-active-cell vertex (buffer.rs:150-274)
//   E4  quad_kernel           per-edge quads, reference emission order (buffer.rs:288-372)
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
};

constexpr int kThreads = 256;
constexpr uint32_t kMaxChunkWords = 256;  // one word per thread in E2b

// ---------------------------------------------------------------------------
// sample index -> lattice coordinates
// ---------------------------------------------------------------------------
// Samples of one span are enumerated so that (a) every warp is full and (b) a
// warp covers a compact 2x4x4 brick of the R^3 core (iteration counts are
// spatially coherent; bricks cut divergence loss from ~13% to ~4% versus
// 32-long z rows).  The (R+1)^3 lattice = R^3 core + three R^2 faces + three
// R-long edges + 1 corner; every piece has power-of-two extent, so decoding is
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
// K1: sample grids.  One thread per sample, gridDim.y = spans of the group.
// ---------------------------------------------------------------------------
template <bool kFast, int kVariant>
__global__ void __launch_bounds__(kThreads)
sample_grids_kernel(ShapeDev sh, const SpanGeom* __restrict__ geom, uint32_t R, uint32_t lg, float inv_r,
                    float* __restrict__ grids, size_t grid_stride) {
    const uint32_t n = R + 1u;
    const uint32_t n3 = n * n * n;
    const uint32_t i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= n3) return;
    uint32_t x, y, z;
    decode_sample(i, R, lg, x, y, z);
    const SpanGeom g = geom[blockIdx.y];
    // v = (x,y,z) as f32 / R  (exact: R is a power of two);  p = start + across * v  (buffer.rs:79-80)
    const float px = __fadd_rn(g.s[0], __fmul_rn(g.across[0], __fmul_rn((float)x, inv_r)));
    const float py = __fadd_rn(g.s[1], __fmul_rn(g.across[1], __fmul_rn((float)y, inv_r)));
    const float pz = __fadd_rn(g.s[2], __fmul_rn(g.across[2], __fmul_rn((float)z, inv_r)));
    const float d = shape_de<kFast, kVariant>(sh, px, py, pz);
    grids[(size_t)blockIdx.y * grid_stride + ((size_t)x * n + y) * n + z] = d;   // util/grid.rs:45-48
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
// E1: classification.  One CTA per chunk of <= 256 words (32 cells per word,
// cube(R) order).  gridDim = (chunks_per_span, spans).
// ---------------------------------------------------------------------------
struct Masks {
    uint32_t* active;  // cell crosses the surface (buffer.rs:130-141)
    uint32_t* ex;      // +x edge from the cell's lower corner emits a quad (:302)
    uint32_t* ey;      // +y edge (:326)
    uint32_t* ez;      // +z edge (:350)
    uint32_t* neg;     // dists[(x,y,z)] < 0.0 (winding, :310)
};

__global__ void __launch_bounds__(kThreads)
classify_kernel(const float* __restrict__ grids, size_t grid_stride, uint32_t R, uint32_t lg,
                uint32_t words_per_span, uint32_t chunk_words, Masks m, uint2* __restrict__ chunk_counts) {
    const uint32_t span = blockIdx.y, chunk = blockIdx.x;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t n = R + 1u, R3 = R << (2 * lg);
    const float* __restrict__ g = grids + (size_t)span * grid_stride;
    uint32_t vcnt = 0, qcnt = 0;
    for (uint32_t w = warp; w < chunk_words; w += kThreads / 32) {
        const uint32_t word = chunk * chunk_words + w;
        const uint32_t c = (word << 5) | lane;
        bool act = false, fx = false, fy = false, fz = false, ng = false;
        if (c < R3) {
            const uint32_t x = c >> (2 * lg), y = (c >> lg) & (R - 1u), z = c & (R - 1u);
            const float* p = g + ((size_t)x * n + y) * n + z;
            const size_t sy = n, sx = (size_t)n * n;
            // corner id = 4*dx + 2*dy + dz (buffer.rs:116-125)
            const float d0 = p[0], d1 = p[1], d2 = p[sy], d3 = p[sy + 1];
            const float d4 = p[sx], d5 = p[sx + 1], d6 = p[sx + sy], d7 = p[sx + sy + 1];
            const uint32_t s0 = __float_as_uint(d0) >> 31;
            const uint32_t s1 = __float_as_uint(d1) >> 31, s2 = __float_as_uint(d2) >> 31;
            const uint32_t s3 = __float_as_uint(d3) >> 31, s4 = __float_as_uint(d4) >> 31;
            const uint32_t s5 = __float_as_uint(d5) >> 31, s6 = __float_as_uint(d6) >> 31;
            const uint32_t s7 = __float_as_uint(d7) >> 31;
            const uint32_t sum = s0 + s1 + s2 + s3 + s4 + s5 + s6 + s7;
            act = (sum != 0u) && (sum != 8u);
            fx = (y > 0u) && (z > 0u) && (s0 != s4);
            fy = (x > 0u) && (z > 0u) && (s0 != s2);
            fz = (x > 0u) && (y > 0u) && (s0 != s1);
            ng = d0 < 0.0f;
        }
        const uint32_t ba = __ballot_sync(0xffffffffu, act);
        const uint32_t bx = __ballot_sync(0xffffffffu, fx);
        const uint32_t by = __ballot_sync(0xffffffffu, fy);
        const uint32_t bz = __ballot_sync(0xffffffffu, fz);
        const uint32_t bn = __ballot_sync(0xffffffffu, ng);
        if (lane == 0u) {
            const size_t o = (size_t)span * words_per_span + word;
            m.active[o] = ba; m.ex[o] = bx; m.ey[o] = by; m.ez[o] = bz; m.neg[o] = bn;
        }
        vcnt += __popc(ba);
        qcnt += __popc(bx) + __popc(by) + __popc(bz);
    }
    __shared__ uint32_t sv[kThreads / 32], sq[kThreads / 32];
    if (lane == 0u) { sv[warp] = vcnt; sq[warp] = qcnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tv = 0, tq = 0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) { tv += sv[w]; tq += sq[w]; }
        chunk_counts[(size_t)span * gridDim.x + chunk] = make_uint2(tv, tq);
    }
}

// ---------------------------------------------------------------------------
// E2a: exclusive scan over the group's chunk counts (single CTA), span offset
// tables, capacity check, running totals.
// ---------------------------------------------------------------------------
constexpr int kScanThreads = 1024;

__global__ void __launch_bounds__(kScanThreads)
scan_chunks_kernel(const uint2* __restrict__ chunk_counts, uint2* __restrict__ chunk_pre, uint32_t nchunks,
                   uint32_t chunks_per_span, uint32_t span0, uint32_t nspans_group,
                   unsigned long long* __restrict__ v_off, unsigned long long* __restrict__ i_off,
                   unsigned long long vcap, unsigned long long icap, MeshState* __restrict__ st) {
    const unsigned long long base_v = st->total_v, base_q = st->total_q;
    uint32_t carry_v = 0, carry_q = 0;
    for (uint32_t t0 = 0; t0 < nchunks; t0 += kScanThreads) {
        const uint32_t i = t0 + threadIdx.x;
        uint2 c = make_uint2(0u, 0u);
        if (i < nchunks) c = chunk_counts[i];
        uint32_t ev, eq, tv, tq;
        block_exclusive_scan2<kScanThreads>(c.x, c.y, ev, eq, tv, tq);
        if (i < nchunks) {
            const uint32_t pv = carry_v + ev, pq = carry_q + eq;
            chunk_pre[i] = make_uint2(pv, pq);
            if (i % chunks_per_span == 0u) {
                const uint32_t s = span0 + i / chunks_per_span;
                v_off[s] = base_v + pv;
                i_off[s] = 6ull * (base_q + pq);
            }
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
    }
}

// ---------------------------------------------------------------------------
// E2b: per-word prefixes (span-local vertex id base, group-local quad slot) and
// the compacted list of active cells.  Same grid as E1; one word per thread.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
apply_prefix_kernel(Masks m, const uint2* __restrict__ chunk_pre, uint32_t words_per_span, uint32_t chunk_words,
                    uint32_t lg3 /* log2(R^3) */, uint32_t* __restrict__ word_vpre, uint32_t* __restrict__ word_qpre,
                    uint32_t* __restrict__ cell_of, uint32_t cell_cap) {
    const uint32_t span = blockIdx.y, chunk = blockIdx.x;
    const uint32_t t = threadIdx.x;
    const size_t o = (size_t)span * words_per_span + chunk * chunk_words + t;
    uint32_t ma = 0, v = 0, q = 0;
    if (t < chunk_words) {
        ma = m.active[o];
        v = __popc(ma);
        q = __popc(m.ex[o]) + __popc(m.ey[o]) + __popc(m.ez[o]);
    }
    uint32_t ev, eq, tv, tq;
    block_exclusive_scan2<kThreads>(v, q, ev, eq, tv, tq);
    if (t < chunk_words) {
        const uint2 cp = chunk_pre[(size_t)span * gridDim.x + chunk];
        const uint32_t span_v0 = chunk_pre[(size_t)span * gridDim.x].x;   // group-local slot of the span's first vertex
        const uint32_t gv = cp.x + ev;                                    // group-local vertex slot
        word_vpre[o] = gv - span_v0;                                      // span-local vertex id
        word_qpre[o] = cp.y + eq;                                         // group-local quad slot
        const uint32_t cell0 = (span << lg3) + ((chunk * chunk_words + t) << 5);
        uint32_t k = 0;
        while (ma) {
            const uint32_t b = __ffs(ma) - 1;
            ma &= ma - 1u;
            if (gv + k < cell_cap) cell_of[gv + k] = cell0 + b;
            ++k;
        }
    }
}

// ---------------------------------------------------------------------------
// E3: vertices.  Persistent grid-stride loop over the group's compacted active
// cells (the count lives on the device, no host sync).  One thread per vertex.
// ---------------------------------------------------------------------------
struct VertexOut { float px, py, pz, nx, ny, nz, d; };

template <bool kFast, int kVariant>
__global__ void __launch_bounds__(kThreads)
vertex_kernel(ShapeDev sh, const SpanGeom* __restrict__ geom, const float* __restrict__ grids, size_t grid_stride,
              uint32_t R, uint32_t lg, const uint32_t* __restrict__ cell_of, uint32_t cell_cap, MeshState* st,
              uint32_t span0, float* __restrict__ out_v, unsigned long long vcap) {
    using M = MathExact;
    const uint32_t nv = min(st->group_v, cell_cap);
    const unsigned long long base_v = st->group_base_v;
    const uint32_t n = R + 1u;
    const uint32_t lg3 = 3 * lg;
    for (uint32_t v = blockIdx.x * kThreads + threadIdx.x; v < nv; v += gridDim.x * kThreads) {
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

        // p0 = span.start + (x,y,z) * step  (buffer.rs:150-151)
        const float p0x = M::add(g.s[0], M::mul((float)x, g.step[0]));
        const float p0y = M::add(g.s[1], M::mul((float)y, g.step[1]));
        const float p0z = M::add(g.s[2], M::mul((float)z, g.step[2]));

        // 12 edges (buffer.rs:155-176): from/to corner ids packed 4 bits each
        // x edges (0,4)(1,5)(2,6)(3,7); y edges (0,2)(1,3)(4,6)(5,7); z edges (0,1)(2,3)(4,5)(6,7)
        const unsigned long long kFrom = 0x642054103210ull, kTo = 0x753176327654ull;
        int count = 0;
        float sx_ = 0.0f, sy_ = 0.0f, sz_ = 0.0f;
        bool bad = false;
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
            const float ax = M::add(p0x, (from & 4) ? g.step[0] : 0.0f), bx = M::add(p0x, (to & 4) ? g.step[0] : 0.0f);
            const float ay = M::add(p0y, (from & 2) ? g.step[1] : 0.0f), by = M::add(p0y, (to & 2) ? g.step[1] : 0.0f);
            const float az = M::add(p0z, (from & 1) ? g.step[2] : 0.0f), bz = M::add(p0z, (to & 1) ? g.step[2] : 0.0f);
            sx_ = M::add(sx_, M::add(M::mul(ax, om), M::mul(bx, w)));
            sy_ = M::add(sy_, M::add(M::mul(ay, om), M::mul(by, w)));
            sz_ = M::add(sz_, M::add(M::mul(az, om), M::mul(bz, w)));
            ++count;
        }
        if (bad) atomicMin(&st->panic_span, span0 + span);
        // centroid (buffer.rs:247-250): origin + sum / count
        const float fc = (float)count;
        const float qx = M::add(0.0f, M::div(sx_, fc));
        const float qy = M::add(0.0f, M::div(sy_, fc));
        const float qz = M::add(0.0f, M::div(sz_, fc));

        // dist_p and the un-normalised central differences (buffer.rs:254-265).
        // unit_x() * d = (1*d, 0*d, 0*d): the zero products keep their sign.
        float de[7];
#pragma unroll 1
        for (int k = 0; k < 7; ++k) {
            const int axis = (k - 1) >> 1;                 // k=0: none
            const float sg = (k & 1) ? 1.0f : -1.0f;       // odd k: +delta, even k: -delta
            float ex = qx, ey = qy, ez = qz;
            if (k > 0) {
                // the reference multiplies the unit vector by +-delta.<axis>, so the two
                // off-axis components are p + 0 * (+-delta.<axis>)
                const float da = sg * (axis == 0 ? g.delta[0] : axis == 1 ? g.delta[1] : g.delta[2]);
                ex = M::add(qx, M::mul(axis == 0 ? 1.0f : 0.0f, da));
                ey = M::add(qy, M::mul(axis == 1 ? 1.0f : 0.0f, da));
                ez = M::add(qz, M::mul(axis == 2 ? 1.0f : 0.0f, da));
            }
            de[k] = shape_de<kFast, kVariant>(sh, ex, ey, ez);
        }
        const float nx = M::sub(de[1], de[2]), ny = M::sub(de[3], de[4]), nz = M::sub(de[5], de[6]);
        // cgmath normalize: v * (1 / sqrt((x*x + y*y) + z*z))
        const float mag = M::sqrt(M::add(M::add(M::mul(nx, nx), M::mul(ny, ny)), M::mul(nz, nz)));
        const float inv = M::div(1.0f, mag);
        float* o = out_v + slot * 7ull;
        o[0] = qx; o[1] = qy; o[2] = qz;
        o[3] = M::mul(nx, inv); o[4] = M::mul(ny, inv); o[5] = M::mul(nz, inv);
        o[6] = de[0];
    }
}

// ---------------------------------------------------------------------------
// E4: quads.  Same grid as E1 (thread per cell == per lower corner).
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t vertex_id(const uint32_t* __restrict__ active, const uint32_t* __restrict__ word_vpre,
                                              size_t span_w0, uint32_t c) {
    const size_t o = span_w0 + (c >> 5);
    return word_vpre[o] + __popc(active[o] & ((1u << (c & 31u)) - 1u));
}

__global__ void __launch_bounds__(kThreads)
quad_kernel(Masks m, const uint32_t* __restrict__ word_vpre, const uint32_t* __restrict__ word_qpre,
            uint32_t R, uint32_t lg, uint32_t words_per_span, uint32_t chunk_words,
            const MeshState* __restrict__ st, uint32_t* __restrict__ out_idx, unsigned long long icap) {
    const uint32_t span = blockIdx.y, chunk = blockIdx.x;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned long long base_q = st->group_base_q;
    const size_t w0 = (size_t)span * words_per_span;
    const uint32_t R2 = R << lg;
    for (uint32_t w = warp; w < chunk_words; w += kThreads / 32) {
        const uint32_t word = chunk * chunk_words + w;
        const size_t o = w0 + word;
        const uint32_t bx = m.ex[o], by = m.ey[o], bz = m.ez[o];
        if ((bx | by | bz) == 0u) continue;
        const uint32_t bit = 1u << lane, lt = bit - 1u;
        const uint32_t hx = (bx >> lane) & 1u, hy = (by >> lane) & 1u, hz = (bz >> lane) & 1u;
        if ((hx | hy | hz) == 0u) continue;
        const bool neg = (m.neg[o] >> lane) & 1u;
        const uint32_t c = (word << 5) | lane;
        // quads before this corner within the word, in corner order
        unsigned long long q = base_q + word_qpre[o] + __popc(bx & lt) + __popc(by & lt) + __popc(bz & lt);
        const uint32_t v3 = vertex_id(m.active, word_vpre, w0, c);
        if (hx) {   // buffer.rs:302-323
            const uint32_t v0 = vertex_id(m.active, word_vpre, w0, c - R - 1u);
            const uint32_t v1 = vertex_id(m.active, word_vpre, w0, c - R);
            const uint32_t v2 = vertex_id(m.active, word_vpre, w0, c - 1u);
            if (6ull * q + 6ull <= icap) {
                uint2* d = reinterpret_cast<uint2*>(out_idx + 6ull * q);
                if (neg) { d[0] = make_uint2(v0, v2); d[1] = make_uint2(v1, v1); d[2] = make_uint2(v2, v3); }
                else     { d[0] = make_uint2(v0, v1); d[1] = make_uint2(v2, v1); d[2] = make_uint2(v3, v2); }
            }
            ++q;
        }
        if (hy) {   // buffer.rs:326-347 (winding flipped relative to x/z)
            const uint32_t v0 = vertex_id(m.active, word_vpre, w0, c - R2 - 1u);
            const uint32_t v1 = vertex_id(m.active, word_vpre, w0, c - R2);
            const uint32_t v2 = vertex_id(m.active, word_vpre, w0, c - 1u);
            if (6ull * q + 6ull <= icap) {
                uint2* d = reinterpret_cast<uint2*>(out_idx + 6ull * q);
                if (neg) { d[0] = make_uint2(v0, v1); d[1] = make_uint2(v2, v1); d[2] = make_uint2(v3, v2); }
                else     { d[0] = make_uint2(v0, v2); d[1] = make_uint2(v1, v1); d[2] = make_uint2(v2, v3); }
            }
            ++q;
        }
        if (hz) {   // buffer.rs:350-371
            const uint32_t v0 = vertex_id(m.active, word_vpre, w0, c - R2 - R);
            const uint32_t v1 = vertex_id(m.active, word_vpre, w0, c - R2);
            const uint32_t v2 = vertex_id(m.active, word_vpre, w0, c - R);
            if (6ull * q + 6ull <= icap) {
                uint2* d = reinterpret_cast<uint2*>(out_idx + 6ull * q);
                if (neg) { d[0] = make_uint2(v0, v2); d[1] = make_uint2(v1, v1); d[2] = make_uint2(v2, v3); }
                else     { d[0] = make_uint2(v0, v1); d[1] = make_uint2(v2, v1); d[2] = make_uint2(v3, v2); }
            }
        }
    }
}

// Resets the call state at the start of a mesh call.
__global__ void reset_state_kernel(MeshState* st) {
    st->total_v = 0; st->total_q = 0; st->group_base_v = 0; st->group_base_q = 0;
    st->group_v = 0; st->group_q = 0; st->overflow = 0; st->panic_span = 0xFFFFFFFFu;
}

}  // namespace ctc
