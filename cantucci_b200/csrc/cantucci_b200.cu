// cantucci_b200.cu -- C ABI over the sm_100a kernels (see include/cantucci_b200.h).
//
// Host side of the drop-in boundary: argument checks that mirror the
// reference's asserts, span geometry, launch grouping, device workspace and the
// host<->device copies of the host-pointer entry points.  No CPU fallback: if
// CUDA is unavailable every entry point fails.
#include "../../include/cantucci_b200.h"
#include "kernels.cuh"

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#endif
#include <chrono>
#include <condition_variable>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace ctc;

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        // grow geometrically to avoid re-allocation churn
        size_t want = bytes + bytes / 4;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { (void)cudaGetLastError(); want = bytes; e = cudaMalloc(&p, want); }
        if (e == cudaSuccess) cap = want; else p = nullptr;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return static_cast<T*>(p); }
};

struct PinnedBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
        cudaError_t e = cudaMallocHost(&p, bytes);
        if (e == cudaSuccess) cap = bytes; else p = nullptr;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct EventPair { cudaEvent_t a, b; int pass; };

// Default of the sign-trust band (profiles/sign_probe_r2.md): over the 1.1 G samples of the benched volume the
// largest |r2_fast - r2_exact| / 2^lemax is 4.2e-6 = 2^-17.9 (5.5e-6 with (8, 5.0)); every sign mismatch sits
// below 2^-22.  2^-17 re-evaluates 0.46 % of the samples.
constexpr float kDefaultKappa = 1.0f / 131072.0f;   // 2^-17

// Widens packed 8-byte quad records (kernels.cuh, store_quad) into the reference's six u32 indices on HOST
// threads.  ctc_mesh_spans ships the records over PCIe (a third of the index bytes) and a small pool of
// workers widens each launch group's slice into the caller's index buffer as soon as its copy has landed,
// while the following groups are still computing / copying.  Every worker takes the same share of every task.
// kind 0: widen n quad records src -> 6 n u32 at dst;  kind 1: copy `bytes` bytes src -> dst in n 16-byte units
struct ExpandTask { cudaEvent_t ready; const void* src; void* dst; size_t n; int kind; size_t bytes; };

void expand_quads_scalar(const uint2* __restrict__ rec, size_t n, uint32_t* __restrict__ out) {
    for (size_t q = 0; q < n; ++q) {
        const uint2 r = rec[q];
        const uint32_t a = r.x & 0xFFFFu, b = r.x >> 16, v2 = r.y & 0xFFFFu, v3 = r.y >> 16;
        const bool flip = a > b;
        const uint32_t v0 = flip ? b : a, v1 = flip ? a : b;
        uint32_t* d = out + 6 * q;
        if (flip) { d[0] = v0; d[1] = v2; d[2] = v1; d[3] = v1; d[4] = v2; d[5] = v3; }     // [v0,v2,v1, v1,v2,v3]
        else      { d[0] = v0; d[1] = v1; d[2] = v2; d[3] = v1; d[4] = v3; d[5] = v2; }     // [v0,v1,v2, v1,v3,v2]
    }
}

#if defined(__x86_64__) && defined(__GNUC__)
// Four quads per step: zero-extend the four ids of each, order the first two, and let one cross-lane permute
// per quad place its six indices where they fall in the 96-byte group (the permute pattern is the winding);
// three blends assemble three aligned 32-byte lines, written with NON-TEMPORAL stores: the index buffer is
// write-only here, and a cached store would first read every line from a memory system the device->host
// copies of the vertices are saturating at the same time.
__attribute__((target("avx2"))) static inline __m256i quad_ids(const uint2* rec, __m256i* flip) {
    const __m128i w = _mm_cvtepu16_epi32(_mm_loadl_epi64(reinterpret_cast<const __m128i*>(rec)));   // a b v2 v3
    const __m128i sw = _mm_shuffle_epi32(w, 0xE1);                                                    // b a v2 v3
    const __m128i v = _mm_blend_epi32(_mm_min_epu32(w, sw), _mm_max_epu32(w, sw), 0x2);               // v0 v1 v2 v3
    *flip = _mm256_broadcastd_epi32(_mm_cmpgt_epi32(w, sw));                                          // a > b
    return _mm256_castsi128_si256(v);
}
__attribute__((target("avx2"))) void expand_quads_avx2(const uint2* __restrict__ rec, size_t n, uint32_t* __restrict__ out) {
    // keep = [v0,v1,v2, v1,v3,v2], flip = [v0,v2,v1, v1,v2,v3], rotated to the record's place in the group of four
    const __m256i k0 = _mm256_setr_epi32(0, 1, 2, 1, 3, 2, 0, 0), f0 = _mm256_setr_epi32(0, 2, 1, 1, 2, 3, 0, 0);   // q0: lanes 0..5
    const __m256i k1 = _mm256_setr_epi32(2, 1, 3, 2, 0, 0, 0, 1), f1 = _mm256_setr_epi32(1, 1, 2, 3, 0, 0, 0, 2);   // q1: [2..5] -> 0..3, [0..1] -> 6..7
    const __m256i k2 = _mm256_setr_epi32(3, 2, 0, 0, 0, 1, 2, 1), f2 = _mm256_setr_epi32(2, 3, 0, 0, 0, 2, 1, 1);   // q2: [4..5] -> 0..1, [0..3] -> 4..7
    const __m256i k3 = _mm256_setr_epi32(0, 0, 0, 1, 2, 1, 3, 2), f3 = _mm256_setr_epi32(0, 0, 0, 2, 1, 1, 2, 3);   // q3: [0..5] -> 2..7
    size_t q = 0;
    while (q < n && (reinterpret_cast<uintptr_t>(out + 6 * q) & 31u)) {
        if (q == 4) { expand_quads_scalar(rec + q, n - q, out + 6 * q); return; }     // destination never 32-byte aligned
        expand_quads_scalar(rec + q, 1, out + 6 * q); ++q;
    }
    for (; q + 4 <= n; q += 4) {
        __m256i fl0, fl1, fl2, fl3;
        const __m256i v0 = quad_ids(rec + q, &fl0), v1 = quad_ids(rec + q + 1, &fl1), v2 = quad_ids(rec + q + 2, &fl2), v3 = quad_ids(rec + q + 3, &fl3);
        const __m256i a = _mm256_permutevar8x32_epi32(v0, _mm256_blendv_epi8(k0, f0, fl0));
        const __m256i b = _mm256_permutevar8x32_epi32(v1, _mm256_blendv_epi8(k1, f1, fl1));
        const __m256i c = _mm256_permutevar8x32_epi32(v2, _mm256_blendv_epi8(k2, f2, fl2));
        const __m256i d = _mm256_permutevar8x32_epi32(v3, _mm256_blendv_epi8(k3, f3, fl3));
        __m256i* o = reinterpret_cast<__m256i*>(out + 6 * q);
        _mm256_stream_si256(o, _mm256_blend_epi32(a, b, 0xC0));
        _mm256_stream_si256(o + 1, _mm256_blend_epi32(b, c, 0xF0));
        _mm256_stream_si256(o + 2, _mm256_blend_epi32(c, d, 0xFC));
    }
    _mm_sfence();
    expand_quads_scalar(rec + q, n - q, out + 6 * q);
}
#endif

void expand_quads_host(const uint2* rec, size_t n, uint32_t* out) {
#if defined(__x86_64__) && defined(__GNUC__)
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    if (have_avx2) { expand_quads_avx2(rec, n, out); return; }
#endif
    expand_quads_scalar(rec, n, out);
}

// Worker pool: the caller pushes one task per launch group (records landed in the pinned wire buffer once
// `ready` has fired); workers claim 32 Ki-quad pieces of the tasks in order.
class HostExpander {
public:
    static constexpr size_t kPiece = 32768;
    ~HostExpander() { stop(); }
    bool start(int device, int share = 1) {      // share: contexts of this process that run such a pool side by side
        if (!th_.empty()) return true;
        const unsigned hw = std::thread::hardware_concurrency();
        // (the widening is bound by the host's memory system, not by cores -- 4 / 8 / 16 threads on a 16-thread host: 13.8 /
        // 12.5 / 12.2 ms per benched volume -- and the caller's own pool waits in this call: all but two hardware threads)
        int n = hw >= 8 ? (int)hw - 2 : (hw >= 2 ? (int)hw - 1 : 1);
        if (share > 1) { n /= share; if (n < 2) n = 2; }
        if (const char* e = getenv("CANTUCCI_B200_EXPAND_THREADS")) { const int v = atoi(e); if (v >= 1) n = v; }
        if (n > 32) n = 32;
        try {
            for (int k = 0; k < n; ++k) th_.emplace_back([this, device] { run(device); });
        } catch (...) { stop(); return false; }
        return true;
    }
    void begin() {
        std::lock_guard<std::mutex> lk(mu_);
        tasks_.clear(); closed_ = false; cur_ = 0; piece_ = 0; waited_ = 0; busy_ = 0;
    }
    void push(const ExpandTask& t) {
        std::lock_guard<std::mutex> lk(mu_);
        tasks_.push_back(t);
        cv_.notify_all();
    }
    void finish() {       // no more tasks for this call; returns when every piece has been widened
        std::unique_lock<std::mutex> lk(mu_);
        closed_ = true;
        cv_.notify_all();
        done_.wait(lk, [&] { return cur_ >= tasks_.size() && busy_ == 0; });
    }
    void stop() {
        { std::lock_guard<std::mutex> lk(mu_); quit_ = true; cv_.notify_all(); }
        for (std::thread& t : th_) if (t.joinable()) t.join();
        th_.clear();
        quit_ = false;
    }
    size_t threads() const { return th_.size(); }
private:
    void run(int device) {
        cudaSetDevice(device);
        std::unique_lock<std::mutex> lk(mu_);
        for (;;) {
            cv_.wait(lk, [&] { return quit_ || cur_ < tasks_.size(); });
            if (quit_) return;
            const size_t ti = cur_;
            const ExpandTask t = tasks_[ti];
            if (waited_ <= ti) {              // the task's copy has not been waited for yet: one worker does (blocking event)
                if (waiting_) { cv_.wait(lk, [&] { return quit_ || !waiting_; }); continue; }
                waiting_ = true;
                lk.unlock();
                cudaEventSynchronize(t.ready);
                lk.lock();
                waiting_ = false; waited_ = ti + 1;
                cv_.notify_all();
                continue;
            }
            const size_t lo = piece_ * kPiece;
            if (lo >= t.n) { if (cur_ == ti) { ++cur_; piece_ = 0; if (cur_ >= tasks_.size()) done_.notify_all(); } continue; }
            ++piece_; ++busy_;
            lk.unlock();
            const size_t hi = lo + kPiece < t.n ? lo + kPiece : t.n;
            if (t.kind == 0) {
                expand_quads_host(static_cast<const uint2*>(t.src) + lo, hi - lo, static_cast<uint32_t*>(t.dst) + 6 * lo);
            } else {
                memcpy(static_cast<char*>(t.dst) + 16 * lo, static_cast<const char*>(t.src) + 16 * lo, (16 * hi < t.bytes ? 16 * hi : t.bytes) - 16 * lo);
            }
            lk.lock();
            if (--busy_ == 0) done_.notify_all();
        }
    }
    std::vector<std::thread> th_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    std::vector<ExpandTask> tasks_;
    size_t cur_ = 0, piece_ = 0, waited_ = 0;
    int busy_ = 0;
    bool closed_ = false, quit_ = false, waiting_ = false;
};

}  // namespace

struct MeshRequest;

struct ctc_ctx {
    int device = 0;
    int num_sms = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::mutex mu;
    std::string err;
    // submission queue of the host-pointer mesh call: concurrent callers (the reference meshes one leaf per
    // thread-pool job, mesh/mod.rs:141-148) are coalesced into batched launches by whichever caller leads
    std::mutex q_mu;
    std::condition_variable q_cv;
    std::vector<MeshRequest*> q;
    bool q_leader = false;
    bool coalesce = true;       // ctc_ctx_set_coalescing
    uint64_t batches = 0, batched_requests = 0;
    std::vector<ctc_span> b_spans;                 // staging of a coalesced batch
    std::vector<uint64_t> b_voff, b_ioff;
    PinnedBuf b_v, b_i;
    bool async_inflight = false;   // an asynchronous (device-pointer) call may still be using the staging buffers
    uint64_t launches = 0;
    uint32_t group_spans = 0;   // 0 = auto
    bool timing = true;
    bool kernel_timing = false; // ctc_ctx_set_kernel_timing: an event pair around every kernel (measurement runs)
    double kernel_ms[CTC_NUM_KERNELS] = {0, 0, 0, 0, 0, 0, 0, 0};
    bool overlap = true;        // ctc_ctx_set_overlap
    bool wire_quads = false;    // ctc_ctx_set_index_wire: ctc_mesh_spans delivers packed 8-byte quad records
    // host destinations: indices cross PCIe as packed quad records and are widened by host threads (default on;
    // a span with >= 65536 vertices makes the call fall back to u32 indices on the wire)
    int host_wire = 1;          // ctc_ctx_set_host_index_wire: 0 off, 1 calls of >= 128 spans, 2 every call
    // ... of every hybrid_den launch groups hybrid_num go packed (widened by host threads), the others as six u32 per quad
    // straight into the caller's buffer: the packed wire loads the host's memory system, the u32 wire the PCIe link.
    // Measured on the benched volume (scripts/gpu_e2e_share.py): 4/4 10.5-10.8 ms, 3/4 11.2, 2/4 11.8, 1/4 13.7, 0/4 13.0
    // -- the link is the scarcer resource on the measured hosts, so everything travels packed by default.
    int hybrid_num = 1, hybrid_den = 1;        // CANTUCCI_B200_HOST_WIRE_SHARE = "num/den"
    std::vector<uint8_t> group_packed;         // per launch group of the last pipelined call
    uint64_t last_d2h_bytes = 0;               // mesh bytes the last host-pointer call copied device -> host
    bool last_wire_overflow = false;
    uint64_t host_wire_calls = 0, host_wire_fallbacks = 0;
    HostExpander expander;
    // peer destination, packed wire: after every launch group's records a progress word {done:1 | epoch:23 | quads:40}
    // is put into the destination GPU's memory, so that it can widen the slices that have landed while this GPU is still computing
    void* wire_progress = nullptr;          // ctc_ctx_set_wire_progress
    uint64_t wire_epoch = 0;
    PinnedBuf h_wire_prog;
    int expander_share = 1;     // ctc_multi: one pool per device, the host threads are shared out
    PinnedBuf h_wire, h_vstage;   // pinned landing buffers: packed quad records; vertices bound for PAGEABLE memory
    std::vector<cudaEvent_t> wire_events;
    // fast mode's sign-trust band (de_device.cuh, fast_suspect_*); calibrated by ctc_fast_sign_probe
    float kappa = kDefaultKappa;

    // workspace
    DevBuf suspects, suspect_count;               // fast mode: K1's suspect lists (double-buffered like the grids)
    DevBuf geom, grids, sign_bits, m_active, word_vpre, cell_of, quad_of, span_first, neg8, chunk_cnt, span_tot, span_pre, state;
    DevBuf out_v, out_idx, out_wire, off_v, off_i; // host-pointer entry points (out_wire: packed quad records of the hybrid index wire)
    DevBuf pts_in, pts_out;
    PinnedBuf h_geom, h_state, h_tables, h_pts;

    // extraction (passes 2-3) runs on its own stream so that group g's small, latency-bound kernels
    // overlap group g+1's DE kernel; grids and sign planes are double-buffered for that
    cudaStream_t ext_stream = nullptr;
    std::vector<cudaEvent_t> k1_done, ext_done;

    // pipelined device->host copies of the host-pointer mesh call
    cudaStream_t copy_stream = nullptr, copy_stream2 = nullptr;   // vertices / indices
    unsigned long long* progress_h = nullptr;     // mapped pinned: totals after each group
    unsigned long long* progress_d = nullptr;
    size_t progress_cap = 0;                      // groups
    std::vector<cudaEvent_t> group_events;
    size_t n_groups = 0;

    // events of the last mesh call
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    std::vector<EventPair> ev_pairs;
    bool mesh_pending = false;

    // interop buffers of this context (interop.inc): base address -> mapped size
    std::map<void*, size_t> interop;
};

void interop_release_all(ctc_ctx* c);     // interop.inc

namespace {

int fail(ctc_ctx* c, int code, const char* what) {
    if (c) c->err = std::string(what);       // (copies first: `what` may point into c->err's old buffer)
    return code;
}
int fail_cuda(ctc_ctx* c, cudaError_t e, const char* where) {
    if (c) { c->err = std::string(where) + ": " + cudaGetErrorString(e); }
    (void)cudaGetLastError();
    return CTC_ERR_CUDA;
}
#define CK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) return fail_cuda(ctx, _e, #call); } while (0)

// No C++ exception may cross the C ABI (std::vector / std::string can throw bad_alloc).
template <class F>
int guarded(ctc_ctx* ctx, F&& body) {
    try {
        return body();
    } catch (const std::exception& e) {
        return fail(ctx, CTC_ERR_CUDA, e.what());
    } catch (...) {
        return fail(ctx, CTC_ERR_CUDA, "unknown C++ exception");
    }
}

// Smallest lemax for which the device's test "kappa * amp_factor(lemax) < 4" fails (amp_factor, de_device.cuh:
// the float whose bits are bias + min(lemax, 2^29), monotone in lemax): escaped samples compare integers.
int32_t amp_trust_threshold(float kappa) {
    auto trusted = [kappa](int32_t le) {
        const int32_t bits = 0x3f800000 + (le < 0x20000000 ? le : 0x20000000);
        float f; memcpy(&f, &bits, 4);
        volatile float prod = kappa * f;
        return prod < 4.0f;
    };
    if (!trusted(0)) return 0;
    if (trusted(0x20000000)) return 0x7fffffff;
    int32_t lo = 0, hi = 0x20000000;            // trusted(lo), !trusted(hi)
    while (hi - lo > 1) { const int32_t mid = lo + (hi - lo) / 2; if (trusted(mid)) lo = mid; else hi = mid; }
    return hi;
}

// The reference's asserts on the shape (mandelbulb.rs:20) and what this build supports.
int check_shape(ctc_ctx* ctx, const ctc_shape* s, ShapeDev* out) {
    if (!s) return fail(ctx, CTC_ERR_INVALID_ARGUMENT, "shape is NULL");
    ShapeDev d{};
    d.kind = s->kind;
    if (s->kind == CTC_SHAPE_MANDELBULB) {
        if (s->max_iters < 1) return fail(ctx, CTC_ERR_INVALID_ARGUMENT, "assert!(max_iters >= 1) (mandelbulb.rs:20)");
        if (s->power < 1 || s->power > 255) return fail(ctx, CTC_ERR_INVALID_ARGUMENT, "power must be in 1..=255 (const generic P: u8; P = 0 has no r^(P-1))");
        d.power = s->power;
        d.max_iters = s->max_iters > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)s->max_iters;
        d.bailout = s->bailout;
        { volatile float b2 = s->bailout * s->bailout; d.bail2 = b2; }
        // The band is calibrated on orbits of up to 12 iterations (every BASELINE config with the default
        // (6, 2.5): zero sign mismatches over 1.1 G samples with an 8x margin).  Longer orbits are chaotic near
        // the surface -- a 1-ulp change of the input moves the reference's own result by O(1) there -- and the
        // same band leaves ~1e-7 (32 iterations) to ~7e-6 (128) of the samples mismatched; a wider band
        // (ctc_ctx_set_fast_band) trades speed for fewer of them: profiles/sign_probe_r2.md.
        const float kappa = ctx ? ctx->kappa : kDefaultKappa;
        d.kappa = kappa;
        d.le_trust = amp_trust_threshold(kappa);
    } else if (s->kind == CTC_SHAPE_SPHERE) {
        d.cx = s->center[0]; d.cy = s->center[1]; d.cz = s->center[2]; d.radius = s->radius;
    } else {
        return fail(ctx, CTC_ERR_INVALID_ARGUMENT, "unknown shape kind");
    }
    if (s->flags & ~(uint32_t)CTC_MATH_FAST) return fail(ctx, CTC_ERR_INVALID_ARGUMENT, "unknown shape flags");
    *out = d;
    return CTC_OK;
}

int shape_variant(const ctc_shape* s) {
    if (s->kind == CTC_SHAPE_SPHERE) return kVarSphere;
    return s->power == 8 ? kVarP8 : kVarGeneric;
}

// generate_for_box asserts (buffer.rs:35-39) + GridTable size >= 2 (grid.rs:25)
int check_spans(ctc_ctx* ctx, const ctc_span* spans, size_t nspans, uint32_t R, uint32_t* lg_out) {
    if (nspans && !spans) return fail(ctx, CTC_ERR_INVALID_ARGUMENT, "spans is NULL");
    if (R == 0 || (R & (R - 1)) != 0) return fail(ctx, CTC_ERR_INVALID_ARGUMENT, "assert!(resolution.is_power_of_two()) (buffer.rs:38-39)");
    if (R < 2) return fail(ctx, CTC_ERR_INVALID_ARGUMENT, "assert!(size >= 2) (grid.rs:25)");
    if (R > 1024) return fail(ctx, CTC_ERR_INVALID_ARGUMENT, "resolution > 1024 needs 64-bit sample indices (unsupported; shard into spans)");
    if (nspans >= 0xFFFFFFF0ull) return fail(ctx, CTC_ERR_INVALID_ARGUMENT, "too many spans");
    for (size_t i = 0; i < nspans; ++i)
        for (int c = 0; c < 3; ++c)
            if (!(spans[i].start[c] < spans[i].end[c]))
                return fail(ctx, CTC_ERR_INVALID_ARGUMENT, "assert!(span.start < span.end) (buffer.rs:35-37)");
    uint32_t lg = 0;
    while ((1u << lg) < R) ++lg;
    *lg_out = lg;
    return CTC_OK;
}

// Span geometry in IEEE f32, no contraction (this TU is built with
// -ffp-contract=off for the host compiler).  buffer.rs:64-67, 77, 101, 257.
void make_geom(const ctc_span& sp, uint32_t R, SpanGeom* g) {
    const float fr = (float)R;
    for (int c = 0; c < 3; ++c) {
        volatile float overflow = (sp.end[c] - sp.start[c]) / fr;
        volatile float s0 = sp.start[c] + (-overflow);
        volatile float e0 = sp.end[c] + overflow;
        volatile float across = e0 - s0;
        volatile float step = across / fr;
        volatile float t = 0.7f * across;
        volatile float delta = t / fr;
        g->s[c] = s0; g->across[c] = across; g->step[c] = step; g->delta[c] = delta;
    }
}

struct GroupPlan {
    uint32_t lg, n, words_per_span, chunk_words, chunks_per_span, group_spans;
    size_t n3;
};

GroupPlan plan_groups(const ctc_ctx* ctx, uint32_t R, uint32_t lg, size_t nspans) {
    GroupPlan p{};
    p.lg = lg; p.n = R + 1; p.n3 = (size_t)p.n * p.n * p.n;
    const size_t R3 = (size_t)R * R * R;
    p.words_per_span = (uint32_t)((R3 + 31) / 32);
    p.chunk_words = p.words_per_span < kMaxChunkWords ? p.words_per_span : kMaxChunkWords;
    p.chunks_per_span = p.words_per_span / p.chunk_words;
    size_t g = ctx->group_spans;
    if (g == 0) {
        // ~512 MiB of sample grids per launch group: large enough that the fixed cost of the small
        // kernels (scan, prefix) and of the launches vanishes, small enough to pipeline the
        // device->host copy of one group's mesh behind the next group's compute
        g = (512ull << 20) / (p.n3 * 4);
        if (g < 1) g = 1;
        // ... and at least four groups per call (when there is enough work) so the copy pipeline has stages
        const size_t quarter = (nspans + 3) / 4;
        if (quarter >= 64 && g > quarter) g = quarter;
    }
    // span << 3 lg | cell fits 30 bits (two more carry a quad's edge; the chunk scan's status word keeps 30 bits
    // of vertices and 32 of quads)
    const size_t max_by_cells = (size_t)1 << (30 - 3 * lg > 0 ? 30 - 3 * lg : 0);
    if (g > max_by_cells) g = max_by_cells;
    if (g > 32768) g = 32768;                                                      // gridDim.y
    if (g > nspans) g = nspans;
    if (g < 1) g = 1;
    p.group_spans = (uint32_t)g;
    return p;
}

cudaEvent_t take_event(ctc_ctx* ctx) {
    if (ctx->ev_used == ctx->ev_pool.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
        ctx->ev_pool.push_back(e);
    }
    return ctx->ev_pool[ctx->ev_used++];
}

// pass ids 0..2 = the reference's three passes (always timed); 3 + k = kernel k (only with kernel_timing)
struct PassTimer {
    ctc_ctx* ctx; int pass; cudaStream_t st; cudaEvent_t a = nullptr;
    PassTimer(ctc_ctx* c, int p, cudaStream_t s) : ctx(c), pass(p), st(s) {
        if (p < 3 ? ctx->timing : ctx->kernel_timing) { a = take_event(ctx); if (a) cudaEventRecord(a, st); }
    }
    ~PassTimer() {
        if (a) { cudaEvent_t b = take_event(ctx); if (b) { cudaEventRecord(b, st); ctx->ev_pairs.push_back({a, b, pass}); } }
    }
};

// Suspect-list capacity for a batch of `nspans` grids: 1/16 of the samples (measured: ~0.5 % are
// suspects on the 1024^3 volume); entries beyond it are re-evaluated in place by K1 itself.
size_t suspect_cap(size_t nspans, size_t n3) {
    size_t cap = nspans * n3 / 16 + 4096;
    if (cap > 0x7FFFFFF0ull) cap = 0x7FFFFFF0ull;
    return cap;
}

template <bool kFast, int kVariant>
void launch_sample(ctc_ctx* ctx, const ShapeDev& sh, const SpanGeom* geom, uint32_t R, uint32_t lg, float* grids,
                   size_t stride, uint32_t nspans, size_t n3, uint32_t* sign_bits, uint32_t sign_stride, SuspectList sl,
                   cudaStream_t stream) {
    // R >= 32: the R^3 core and the x = R / y = R faces go to the warp-walk path; a warp walks
    // L = 64 z-samples of its 8 columns when R >= 64 (else 32): per-warp set-up paid once per 512 samples
    const size_t R2 = (size_t)R * R, R3 = R2 * R;
    // (small batches -- the drop-in's steady state is 8 leaves per split -- walk 32: twice the warps, half the
    // dependent chain, the SMs are far from full anyway)
#ifdef CTC_K1_WALK32
    const uint32_t lgw = 0u;
#else
    const uint32_t lgw = (lg >= 6 && (size_t)nspans * R3 >= ((size_t)1 << 23)) ? 1u : 0u;
#endif
    const size_t warp_blocks = lg >= 5 ? (R3 + 2 * R2) / (256u << lgw) : 0;
    constexpr size_t kWarps = kK1Threads / 32;
    const uint32_t core_blocks = (uint32_t)((warp_blocks + kWarps - 1) / kWarps);
    const size_t rest = core_blocks ? R2 + 3 * (size_t)R + 1 : n3;
    dim3 grid(core_blocks + (unsigned)((rest + kK1Threads - 1) / kK1Threads), nspans);
    sample_grids_kernel<kFast, kVariant><<<grid, kK1Threads, 0, stream>>>(sh, geom, R, lg, 1.0f / (float)R, 8.0f / (float)R, grids, stride,
                                                                        sign_bits, sign_stride, core_blocks, lgw, sl);
    ctx->launches++;
}

// Fast mode, power 8: exact re-evaluation of the suspects the K1 launch queued (sign-exact fast mode).
void launch_fixup(ctc_ctx* ctx, const ShapeDev& sh, int variant, const SpanGeom* geom, uint32_t R, float* grids, size_t stride,
                  uint32_t* sign_bits, uint32_t sign_stride, SuspectList sl, MeshState* st, cudaStream_t stream) {
    const unsigned blocks = (unsigned)ctx->num_sms * 6u;
    if (variant == kVarP8)
        fixup_suspects_kernel<kVarP8><<<blocks, kThreads, 0, stream>>>(sh, geom, R, 1.0f / (float)R, grids, stride, sign_bits, sign_stride, sl, st);
    else
        fixup_suspects_kernel<kVarGeneric><<<blocks, kThreads, 0, stream>>>(sh, geom, R, 1.0f / (float)R, grids, stride, sign_bits, sign_stride, sl, st);
    ctx->launches++;
}

#define DISPATCH(fast, variant, CALL)                                                       \
    do {                                                                                    \
        if (fast) {                                                                         \
            if ((variant) == kVarP8) { CALL(true, kVarP8); }                                \
            else if ((variant) == kVarGeneric) { CALL(true, kVarGeneric); }                 \
            else { CALL(true, kVarSphere); }                                                \
        } else {                                                                            \
            if ((variant) == kVarP8) { CALL(false, kVarP8); }                               \
            else if ((variant) == kVarGeneric) { CALL(false, kVarGeneric); }                \
            else { CALL(false, kVarSphere); }                                               \
        }                                                                                   \
    } while (0)

int upload_geom(ctc_ctx* ctx, const ctc_span* spans, size_t nspans, uint32_t R) {
    CK(ctx->h_geom.ensure(nspans * sizeof(SpanGeom)));
    CK(ctx->geom.ensure(nspans * sizeof(SpanGeom)));
    SpanGeom* hg = static_cast<SpanGeom*>(ctx->h_geom.p);
    for (size_t i = 0; i < nspans; ++i) make_geom(spans[i], R, &hg[i]);
    CK(cudaMemcpyAsync(ctx->geom.p, hg, nspans * sizeof(SpanGeom), cudaMemcpyHostToDevice, ctx->stream));
    return CTC_OK;
}

int sample_grids_impl(ctc_ctx* ctx, const ctc_shape* shape, const ctc_span* spans, size_t nspans, uint32_t R,
                      float* d_grids) {
    ShapeDev sh; uint32_t lg;
    int rc = check_shape(ctx, shape, &sh); if (rc) return rc;
    rc = check_spans(ctx, spans, nspans, R, &lg); if (rc) return rc;
    if (nspans == 0) return CTC_OK;
    if (!d_grids) return fail(ctx, CTC_ERR_INVALID_ARGUMENT, "grids is NULL");
    CK(cudaSetDevice(ctx->device));
    // the pinned geometry staging buffer may still be in flight from a previous ASYNCHRONOUS call (the
    // host-pointer entry points return synchronised)
    if (ctx->async_inflight) { CK(cudaStreamSynchronize(ctx->stream)); ctx->async_inflight = false; }
    rc = upload_geom(ctx, spans, nspans, R); if (rc) return rc;
    const size_t n3 = (size_t)(R + 1) * (R + 1) * (R + 1);
    const bool fast = (shape->flags & CTC_MATH_FAST) != 0;
    const int variant = shape_variant(shape);
    const bool listed = fast && variant == kVarP8;     // K1 queues suspects for the exact re-evaluation
    // launch batches: gridDim.y limit, and the suspect list of a batch must fit its u32 counter
    size_t per = 32768;
    while (per > 1 && per * n3 > (size_t)0x7FFFFFF0ull) per /= 2;
    SuspectList sl{nullptr, nullptr, 0u};
    if (listed) {
        const size_t cap = suspect_cap(per < nspans ? per : nspans, n3);
        CK(ctx->suspects.ensure(cap * sizeof(uint2)));
        CK(ctx->suspect_count.ensure(2 * sizeof(unsigned int)));
        sl = SuspectList{ctx->suspects.as<uint2>(), ctx->suspect_count.as<unsigned int>(), (uint32_t)cap};
    }
    for (size_t s0 = 0; s0 < nspans; s0 += per) {
        const uint32_t cnt = (uint32_t)((nspans - s0) < per ? (nspans - s0) : per);
        if (listed) CK(cudaMemsetAsync(sl.count, 0, sizeof(unsigned int), ctx->stream));
#define CALL(F, V) launch_sample<F, V>(ctx, sh, ctx->geom.as<SpanGeom>() + s0, R, lg, d_grids + s0 * n3, n3, cnt, n3, nullptr, 0u, sl, ctx->stream)
        DISPATCH(fast, variant, CALL);
#undef CALL
        if (listed)
            launch_fixup(ctx, sh, variant, ctx->geom.as<SpanGeom>() + s0, R, d_grids + s0 * n3, n3, nullptr, 0u, sl, nullptr, ctx->stream);
    }
    CK(cudaGetLastError());
    return CTC_OK;
}

int de_batch_impl(ctc_ctx* ctx, const ctc_shape* shape, const float* d_xyz, size_t n, float* d_out) {
    ShapeDev sh;
    int rc = check_shape(ctx, shape, &sh); if (rc) return rc;
    if (n == 0) return CTC_OK;
    if (!d_xyz || !d_out) return fail(ctx, CTC_ERR_INVALID_ARGUMENT, "NULL point/output buffer");
    CK(cudaSetDevice(ctx->device));
    const bool fast = (shape->flags & CTC_MATH_FAST) != 0;
    const int variant = shape_variant(shape);
    const unsigned blocks = (unsigned)((n + kThreads - 1) / kThreads);
#define CALL(F, V) de_batch_kernel<F, V><<<blocks, kThreads, 0, ctx->stream>>>(sh, d_xyz, n, d_out)
    DISPATCH(fast, variant, CALL);
#undef CALL
    ctx->launches++;
    CK(cudaGetLastError());
    return CTC_OK;
}

template <bool kFast, int kVariant>
void launch_vertex(ctc_ctx* ctx, const ShapeDev& sh, const SpanGeom* geom, const float* grids, size_t stride, uint32_t R,
                   uint32_t lg, const uint32_t* cell_of, uint32_t cell_cap, uint8_t* neg8, MeshState* st, uint32_t span0, float* out_v,
                   unsigned long long vcap, unsigned blocks, cudaStream_t stream) {
    vertex_kernel<kFast, kVariant><<<blocks * (kThreads / kE3Threads), kE3Threads, 0, stream>>>(sh, geom, grids, stride, R, lg, cell_of, cell_cap, neg8,
                                                                         st, span0, out_v, vcap);
    ctx->launches++;
}

int ensure_progress(ctc_ctx* ctx, size_t groups) {
    if (groups <= ctx->progress_cap) return CTC_OK;
    if (ctx->progress_h) { cudaFreeHost(ctx->progress_h); ctx->progress_h = nullptr; ctx->progress_cap = 0; }
    const size_t cap = groups + 64;
    CK(cudaHostAlloc(reinterpret_cast<void**>(&ctx->progress_h), cap * 2 * sizeof(unsigned long long), cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&ctx->progress_d), ctx->progress_h, 0));
    ctx->progress_cap = cap;
    return CTC_OK;
}

int mesh_spans_impl(ctc_ctx* ctx, const ctc_shape* shape, const ctc_span* spans, size_t nspans, uint32_t R,
                    ctc_vertex* d_v, size_t vcap, uint32_t* d_idx, size_t icap, uint64_t* d_v_off, uint64_t* d_i_off,
                    bool pipeline = false, int packed_quads = 0 /* 1: every group, 2: hybrid_num of hybrid_den groups */,
                    const std::function<int(size_t)>* on_group = nullptr /* called with the number of groups enqueued so far */,
                    uint2* d_wire = nullptr /* packed records go here (at quad offsets) instead of into d_idx */) {
    ShapeDev sh; uint32_t lg;
    int rc = check_shape(ctx, shape, &sh); if (rc) return rc;
    rc = check_spans(ctx, spans, nspans, R, &lg); if (rc) return rc;
    if (!d_v_off || !d_i_off) return fail(ctx, CTC_ERR_INVALID_ARGUMENT, "offset tables are NULL");
    if ((vcap && !d_v) || (icap && !d_idx)) return fail(ctx, CTC_ERR_INVALID_ARGUMENT, "NULL output buffer with non-zero capacity");
    if ((reinterpret_cast<uintptr_t>(d_idx) & 7u) || (reinterpret_cast<uintptr_t>(d_v) & 3u))
        return fail(ctx, CTC_ERR_INVALID_ARGUMENT, "output buffers must be 8-byte (indices) / 4-byte (vertices) aligned");
    CK(cudaSetDevice(ctx->device));
    // staging buffers / events of a previous ASYNCHRONOUS call (host-pointer calls return synchronised)
    if (ctx->async_inflight) { CK(cudaStreamSynchronize(ctx->stream)); ctx->async_inflight = false; }
    ctx->ev_used = 0; ctx->ev_pairs.clear();
    ctx->mesh_pending = true;
    ctx->n_groups = 0;

    CK(ctx->state.ensure(sizeof(MeshState)));
    MeshState* st = ctx->state.as<MeshState>();
    reset_state_kernel<<<1, 1, 0, ctx->stream>>>(st);
    ctx->launches++;
    if (nspans == 0) {
        CK(cudaMemsetAsync(d_v_off, 0, sizeof(uint64_t), ctx->stream));
        CK(cudaMemsetAsync(d_i_off, 0, sizeof(uint64_t), ctx->stream));
        return CTC_OK;
    }
    rc = upload_geom(ctx, spans, nspans, R); if (rc) return rc;

    const GroupPlan gp = plan_groups(ctx, R, lg, nspans);
    const size_t G = gp.group_spans;
    const size_t words = G * gp.words_per_span, chunks = G * gp.chunks_per_span;
    const size_t group_cells = G * ((size_t)R * R * R);
    const uint32_t cell_cap = (uint32_t)(group_cells < vcap ? group_cells : (vcap < 0xFFFFFFFFull ? vcap : 0xFFFFFFFFull));
    // more than one group: overlap extraction(g) with DE(g+1)
    const bool two_streams = (nspans > G) && ctx->overlap;
    const size_t nbuf = two_streams ? 2 : 1;
    CK(ctx->grids.ensure(nbuf * G * gp.n3 * sizeof(float)));
    CK(ctx->m_active.ensure(words * 4));
    const uint32_t sign_stride = (uint32_t)((gp.n3 + 31) / 32 + 1);
    CK(ctx->sign_bits.ensure(nbuf * G * (size_t)sign_stride * 4));
    CK(ctx->word_vpre.ensure(words * 4));
    // quads: at most three per active cell, and never more than the caller's index buffer holds
    const size_t quad_room = 3 * (size_t)cell_cap, quad_out = icap / 6;
    const uint32_t quad_cap = (uint32_t)(quad_room < quad_out ? quad_room : (quad_out < 0xFFFFFFFFull ? quad_out : 0xFFFFFFFFull));
    CK(ctx->cell_of.ensure((size_t)(cell_cap ? cell_cap : 1) * 4));
    CK(ctx->neg8.ensure((size_t)(cell_cap ? cell_cap : 1)));
    CK(ctx->quad_of.ensure((size_t)(quad_cap ? quad_cap : 1) * 4));
    CK(ctx->span_first.ensure(G * 4));
    CK(ctx->chunk_cnt.ensure(chunks * sizeof(uint2))); CK(ctx->span_tot.ensure(G * 8)); CK(ctx->span_pre.ensure(G * sizeof(uint2)));
    const bool fast = (shape->flags & CTC_MATH_FAST) != 0;
    const int variant = shape_variant(shape);
    // fast mode, power 8: K1 queues the samples whose sign cannot be trusted; the extraction stream
    // re-evaluates them exactly before it classifies (lists double-buffered like the grids)
    const bool listed = fast && variant == kVarP8;
    const size_t list_cap = listed ? suspect_cap(G, gp.n3) : 0;
    if (listed) {
        CK(ctx->suspects.ensure(nbuf * list_cap * sizeof(uint2)));
        CK(ctx->suspect_count.ensure(2 * sizeof(unsigned int)));
    }
    const ExtractionLists ls{ctx->m_active.as<uint32_t>(), ctx->word_vpre.as<uint32_t>(), ctx->cell_of.as<uint32_t>(),
                             ctx->quad_of.as<uint32_t>(), ctx->span_first.as<uint32_t>(), ctx->neg8.as<uint8_t>(), cell_cap, quad_cap};

    const unsigned vblocks = (unsigned)ctx->num_sms * 8u;
    // Launch groups.  With the copy pipeline on, the first groups are small (64, 128, 256, ... spans)
    // so that the device->host / peer copies start almost immediately instead of after a full group,
    // and the last ones shrink again (..., 256, 128, 64) so that little is left to copy once the
    // compute has finished.
    std::vector<std::pair<size_t, uint32_t>> groups;
    {
        std::vector<size_t> tail;                       // sizes of the ramp-down groups, last first
        size_t tail_total = 0;
        if (pipeline && ctx->group_spans == 0)
            for (size_t t = 64; t < G && tail_total + t + 64 + 128 + 256 < nspans; t *= 2) { tail.push_back(t); tail_total += t; }
        size_t s0 = 0, ramp = 64;
        const size_t body_end = nspans - tail_total;
        while (s0 < body_end) {
            size_t cnt = G;
            if (pipeline && ctx->group_spans == 0 && ramp < G) { cnt = ramp; ramp *= 2; }
            if (cnt > body_end - s0) cnt = body_end - s0;
            groups.emplace_back(s0, (uint32_t)cnt);
            s0 += cnt;
        }
        for (size_t k = tail.size(); k-- > 0;) { groups.emplace_back(s0, (uint32_t)tail[k]); s0 += tail[k]; }
    }
    const size_t n_groups = groups.size();
    cudaStream_t sA = ctx->stream, sE = ctx->stream;
    if (two_streams) {
        if (!ctx->ext_stream) {
            // highest priority: the short extraction kernels take SM slots as the long DE kernel's CTAs retire
            int lo = 0, hi = 0;
            CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            CK(cudaStreamCreateWithPriority(&ctx->ext_stream, cudaStreamNonBlocking, hi));
        }
        sE = ctx->ext_stream;
        while (ctx->k1_done.size() < n_groups) {
            cudaEvent_t a, b;
            CK(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
            ctx->k1_done.push_back(a); ctx->ext_done.push_back(b);
        }
        // the extraction stream starts after everything already enqueued on the caller's stream
        // (state reset, geometry upload)
        CK(cudaEventRecord(ctx->k1_done[0], sA));
        CK(cudaStreamWaitEvent(sE, ctx->k1_done[0], 0));
    }
    if (pipeline) {
        ctx->group_packed.assign(n_groups, 0);
        rc = ensure_progress(ctx, n_groups); if (rc) return rc;
        while (ctx->group_events.size() < n_groups) {
            cudaEvent_t e;
            CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ctx->group_events.push_back(e);
        }
    }

    for (size_t gi = 0; gi < n_groups; ++gi) {
        const size_t s0 = groups[gi].first;
        float* grids = ctx->grids.as<float>() + (gi % nbuf) * G * gp.n3;
        uint32_t* sign_bits = ctx->sign_bits.as<uint32_t>() + (gi % nbuf) * G * (size_t)sign_stride;
        // this buffer pair was last read by the extraction of group gi-2
        if (two_streams && gi >= 2) CK(cudaStreamWaitEvent(sA, ctx->ext_done[gi - 2], 0));
        const uint32_t cnt = groups[gi].second;
        const SpanGeom* geom = ctx->geom.as<SpanGeom>() + s0;
        SuspectList sl{nullptr, nullptr, 0u};
        if (listed)
            sl = SuspectList{ctx->suspects.as<uint2>() + (gi % nbuf) * list_cap,
                             ctx->suspect_count.as<unsigned int>() + (gi % nbuf), (uint32_t)list_cap};
        {   // pass 1
            PassTimer t(ctx, 0, sA);
            CK(cudaMemsetAsync(sign_bits, 0, (size_t)cnt * sign_stride * 4, sA));
            if (listed) CK(cudaMemsetAsync(sl.count, 0, sizeof(unsigned int), sA));
            PassTimer k(ctx, 3 + CTC_K_SAMPLE_GRIDS, sA);
#define CALL(F, V) launch_sample<F, V>(ctx, sh, geom, R, lg, grids, gp.n3, cnt, gp.n3, sign_bits, sign_stride, sl, sA)
            DISPATCH(fast, variant, CALL);
#undef CALL
        }
        if (two_streams) {
            CK(cudaEventRecord(ctx->k1_done[gi], sA));
            CK(cudaStreamWaitEvent(sE, ctx->k1_done[gi], 0));
        }
        {   // pass 2: (fast mode: sign repair,) classify + compact, vertices
            PassTimer t(ctx, 1, sE);
            if (listed) {
                PassTimer k(ctx, 3 + CTC_K_FIXUP, sE);
                launch_fixup(ctx, sh, variant, geom, R, grids, gp.n3, sign_bits, sign_stride, sl, st, sE);
            }
            dim3 cgrid(gp.chunks_per_span, cnt);
            {
                PassTimer k(ctx, 3 + CTC_K_CLASSIFY, sE);
                CK(cudaMemsetAsync(ctx->span_tot.p, 0, (size_t)cnt * 8, sE));
                classify_count_kernel<<<cgrid, kThreads, 0, sE>>>(sign_bits, sign_stride, R, lg, gp.chunk_words,
                                                                  ctx->chunk_cnt.as<uint2>(), ctx->span_tot.as<unsigned long long>());
            }
            {
                PassTimer k(ctx, 3 + CTC_K_SCAN, sE);
                span_scan_kernel<<<1, kScanThreads, 0, sE>>>(
                    ctx->span_tot.as<unsigned long long>(), ctx->span_pre.as<uint2>(), ls.span_first, (uint32_t)s0, cnt,
                    reinterpret_cast<unsigned long long*>(d_v_off), reinterpret_cast<unsigned long long*>(d_i_off),
                    (unsigned long long)vcap, (unsigned long long)icap, st, pipeline ? ctx->progress_d + 2 * gi : nullptr);
            }
            {
                PassTimer k(ctx, 3 + CTC_K_PREFIX, sE);
                emit_lists_kernel<<<cgrid, kThreads, 0, sE>>>(sign_bits, sign_stride, R, lg, gp.words_per_span, gp.chunk_words,
                                                              ctx->chunk_cnt.as<uint2>(), ctx->span_pre.as<uint2>(), ls);
            }
            ctx->launches += 3;
            PassTimer k(ctx, 3 + CTC_K_VERTEX, sE);
#define CALL(F, V) launch_vertex<F, V>(ctx, sh, geom, grids, gp.n3, R, lg, ls.cell_of, cell_cap, ls.neg8, st, \
                                       (uint32_t)s0, reinterpret_cast<float*>(d_v), (unsigned long long)vcap, vblocks, sE)
            DISPATCH(fast, variant, CALL);
#undef CALL
        }
        {   // pass 3
            PassTimer t(ctx, 2, sE);
            const bool pk = pipeline && (packed_quads == 1 ||
                                         (packed_quads == 2 && (gi * (size_t)ctx->hybrid_num) / ctx->hybrid_den != ((gi + 1) * (size_t)ctx->hybrid_num) / ctx->hybrid_den));
            if (pipeline) ctx->group_packed[gi] = pk ? 1 : 0;
            if (pk)
                quad_kernel<true><<<vblocks, kThreads, 0, sE>>>(ls, R, lg, gp.words_per_span, st,
                                                                d_wire ? reinterpret_cast<uint32_t*>(d_wire) : d_idx, (unsigned long long)icap);
            else
                quad_kernel<false><<<vblocks, kThreads, 0, sE>>>(ls, R, lg, gp.words_per_span, st, d_idx, (unsigned long long)icap);
            ctx->launches++;
        }
        if (pipeline) CK(cudaEventRecord(ctx->group_events[gi], sE));
        if (two_streams) CK(cudaEventRecord(ctx->ext_done[gi], sE));
        // (the host-pointer call starts copying the groups that have already finished while the rest is being enqueued)
        if (pipeline && on_group && gi + 1 < n_groups) { rc = (*on_group)(gi + 1); if (rc) return rc; }
    }
    // the caller's stream owns completion: it joins the extraction stream
    if (two_streams) CK(cudaStreamWaitEvent(sA, ctx->ext_done[n_groups - 1], 0));
    ctx->n_groups = pipeline ? n_groups : 0;
    CK(cudaGetLastError());
    return CTC_OK;
}

int mesh_result_impl(ctc_ctx* ctx, uint64_t* n_vertices, uint64_t* n_indices, ctc_timings* timings,
                     bool state_already_copied = false) {
    if (!ctx->mesh_pending) return fail(ctx, CTC_ERR_INVALID_ARGUMENT, "no mesh call pending");
    CK(cudaSetDevice(ctx->device));
    if (!state_already_copied) {
        CK(ctx->h_state.ensure(sizeof(MeshState)));
        CK(cudaMemcpyAsync(ctx->h_state.p, ctx->state.p, sizeof(MeshState), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->async_inflight = false;
    const MeshState* st = static_cast<const MeshState*>(ctx->h_state.p);
    if (n_vertices) *n_vertices = st->total_v;
    if (n_indices) *n_indices = 6ull * st->total_q;
    if (timings || ctx->kernel_timing) {
        double ms[3 + CTC_NUM_KERNELS] = {0};
        for (const EventPair& p : ctx->ev_pairs) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, p.a, p.b) == cudaSuccess) ms[p.pass] += t;
        }
        if (timings) {
            timings->first_ms = ms[0]; timings->second_ms = ms[1]; timings->third_ms = ms[2];
            timings->vertices = st->total_v; timings->faces = st->total_q;
        }
        for (int k = 0; k < CTC_NUM_KERNELS; ++k) ctx->kernel_ms[k] = ms[3 + k];
        ctx->kernel_ms[CTC_K_QUADS] = ms[2];      // pass 3 is one kernel
    }
    // A truncated output is reported ahead of the lerp assert ("mesh still delivered" must not hide a
    // mesh that is NOT all there); the assert's span stays readable in the message either way.
    char lerp[160] = "";
    if (st->panic_span != 0xFFFFFFFFu)
        snprintf(lerp, sizeof lerp, "lerp factor outside [0,1] in span %u: the reference panics at math.rs:19", st->panic_span);
    ctx->last_wire_overflow = st->wire_overflow != 0;
    if (st->wire_overflow)
        return fail(ctx, CTC_ERR_OVERFLOW, "a span has >= 65536 vertices: packed quad records cannot carry it, use the u32 index wire");
    if (st->overflow)
        return fail(ctx, CTC_ERR_OVERFLOW, (std::string("output capacity too small; required totals reported") +
                                            (lerp[0] ? std::string("; also: ") + lerp : std::string())).c_str());
    if (lerp[0]) return fail(ctx, CTC_ERR_LERP_ASSERT, lerp);
    return CTC_OK;
}

}  // namespace

// ---------------------------------------------------------------------------
// extern "C"
// ---------------------------------------------------------------------------
extern "C" {

int ctc_version(void) { return CTC_VERSION; }

int ctc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    return n;
}

int ctc_ctx_create(int device, ctc_ctx** out) {
    if (!out) return CTC_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { (void)cudaGetLastError(); return CTC_ERR_NO_DEVICE; }
    if (device < 0 || device >= n) return CTC_ERR_INVALID_ARGUMENT;
    if (cudaSetDevice(device) != cudaSuccess) { (void)cudaGetLastError(); return CTC_ERR_CUDA; }
    ctc_ctx* c = new ctc_ctx();
    c->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->num_sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        (void)cudaGetLastError(); delete c; return CTC_ERR_CUDA;
    }
    c->stream = c->own_stream;
    if (const char* e = getenv("CANTUCCI_B200_HOST_WIRE_SHARE")) {        // "num/den" (measurement runs)
        int num = 0, den = 0;
        if (sscanf(e, "%d/%d", &num, &den) == 2 && den > 0 && den <= 1024 && num >= 0 && num <= den) { c->hybrid_num = num; c->hybrid_den = den; }
    }
    *out = c;
    return CTC_OK;
}

void ctc_ctx_destroy(ctc_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (DevBuf* b : {&c->geom, &c->grids, &c->m_active, &c->sign_bits, &c->word_vpre, &c->cell_of, &c->quad_of,
                      &c->span_first, &c->neg8, &c->chunk_cnt, &c->span_tot, &c->span_pre, &c->state, &c->out_v, &c->out_idx, &c->out_wire,
                      &c->off_v, &c->off_i, &c->pts_in, &c->pts_out, &c->suspects, &c->suspect_count})
        b->release();
    interop_release_all(c);
    c->expander.stop();
    c->h_geom.release(); c->h_pts.release(); c->h_state.release(); c->h_tables.release(); c->b_v.release(); c->b_i.release(); c->h_wire.release(); c->h_vstage.release(); c->h_wire_prog.release();
    for (cudaEvent_t e : c->wire_events) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    for (cudaEvent_t e : c->group_events) cudaEventDestroy(e);
    for (cudaEvent_t e : c->k1_done) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ext_done) cudaEventDestroy(e);
    if (c->ext_stream) cudaStreamDestroy(c->ext_stream);
    if (c->progress_h) cudaFreeHost(c->progress_h);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->copy_stream2) cudaStreamDestroy(c->copy_stream2);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

int ctc_ctx_set_stream(ctc_ctx* ctx, void* cuda_stream) {
    if (!ctx) return CTC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    return CTC_OK;
}

int ctc_ctx_set_group_spans(ctc_ctx* ctx, uint32_t spans_per_group) {
    if (!ctx) return CTC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->group_spans = spans_per_group;
    return CTC_OK;
}

int ctc_ctx_set_kernel_timing(ctc_ctx* ctx, int enable) {
    if (!ctx) return CTC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->kernel_timing = enable != 0;
    return CTC_OK;
}

int ctc_mesh_kernel_times(ctc_ctx* ctx, double* ms, size_t n) {
    if (!ctx || !ms) return CTC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    for (size_t k = 0; k < n; ++k) ms[k] = k < (size_t)CTC_NUM_KERNELS ? ctx->kernel_ms[k] : 0.0;
    return CTC_OK;
}

int ctc_ctx_set_overlap(ctc_ctx* ctx, int enable) {
    if (!ctx) return CTC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->overlap = enable != 0;
    return CTC_OK;
}

int ctc_ctx_set_index_wire(ctc_ctx* ctx, int packed_quads) {
    if (!ctx) return CTC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->wire_quads = packed_quads != 0;
    return CTC_OK;
}

int ctc_ctx_set_wire_progress(ctc_ctx* ctx, void* d_progress_word) {
    if (!ctx) return CTC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (reinterpret_cast<uintptr_t>(d_progress_word) & 7u) return fail(ctx, CTC_ERR_INVALID_ARGUMENT, "progress word must be 8-byte aligned");
    ctx->wire_progress = d_progress_word;
    ctx->wire_epoch = 0;
    return CTC_OK;
}

int ctc_ctx_set_host_index_wire(ctc_ctx* ctx, int packed_quads) {
    if (!ctx) return CTC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->host_wire = packed_quads < 0 ? 0 : (packed_quads > 2 ? 2 : packed_quads);
    return CTC_OK;
}

int ctc_ctx_set_host_wire_share(ctc_ctx* ctx, uint32_t num, uint32_t den) {
    if (!ctx || den == 0 || den > 1024 || num > den) return CTC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->hybrid_num = (int)num; ctx->hybrid_den = (int)den;
    return CTC_OK;
}

int ctc_mesh_d2h_bytes(ctc_ctx* ctx, uint64_t* bytes) {
    if (!ctx || !bytes) return CTC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    *bytes = ctx->last_d2h_bytes;
    return CTC_OK;
}

int ctc_ctx_host_index_wire_stats(ctc_ctx* ctx, uint64_t* calls, uint64_t* fallbacks, uint32_t* threads) {
    if (!ctx) return CTC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (calls) *calls = ctx->host_wire_calls;
    if (fallbacks) *fallbacks = ctx->host_wire_fallbacks;
    if (threads) *threads = (uint32_t)ctx->expander.threads();
    return CTC_OK;
}

int ctc_expand_quads_host(const void* records, size_t nquads, uint32_t* idx) {
    if (nquads == 0) return CTC_OK;
    if (!records || !idx || (reinterpret_cast<uintptr_t>(records) & 7u) || (reinterpret_cast<uintptr_t>(idx) & 3u)) return CTC_ERR_INVALID_ARGUMENT;
    expand_quads_host(static_cast<const uint2*>(records), nquads, idx);
    return CTC_OK;
}

int ctc_expand_quads(ctc_ctx* ctx, const void* d_records, size_t nquads, uint32_t* d_idx) {
    if (!ctx) return CTC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (nquads == 0) return CTC_OK;
    if (!d_records || !d_idx || (reinterpret_cast<uintptr_t>(d_idx) & 7u) || (reinterpret_cast<uintptr_t>(d_records) & 7u))
        return fail(ctx, CTC_ERR_INVALID_ARGUMENT, "NULL or misaligned quad buffers");
    CK(cudaSetDevice(ctx->device));
    const unsigned blocks = (unsigned)ctx->num_sms * 8u;
    expand_quads_kernel<<<blocks, kThreads, 0, ctx->stream>>>(static_cast<const uint2*>(d_records), nquads, d_idx);
    ctx->launches++;
    CK(cudaGetLastError());
    return CTC_OK;
}

int ctc_ctx_synchronize(ctc_ctx* ctx) {
    if (!ctx) return CTC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->async_inflight = false;
    return CTC_OK;
}

const char* ctc_last_error(const ctc_ctx* ctx) { return ctx ? ctx->err.c_str() : "no context"; }

size_t ctc_last_error_copy(ctc_ctx* ctx, char* buf, size_t len) {
    if (!ctx) return 0;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (buf && len) {
        const size_t n = ctx->err.size() < len - 1 ? ctx->err.size() : len - 1;
        memcpy(buf, ctx->err.data(), n);
        buf[n] = 0;
    }
    return ctx->err.size();
}

uint64_t ctc_kernel_launches(const ctc_ctx* ctx) { return ctx ? ctx->launches : 0; }

int ctc_de_batch_device(ctc_ctx* ctx, const ctc_shape* shape, const float* d_xyz, size_t n, float* d_out) {
    if (!ctx) return CTC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    return de_batch_impl(ctx, shape, d_xyz, n, d_out);
}

int ctc_de_batch(ctc_ctx* ctx, const ctc_shape* shape, const float* xyz, size_t n, float* out) {
    if (!ctx) return CTC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ShapeDev sh;
    int rc = check_shape(ctx, shape, &sh); if (rc) return rc;
    if (n == 0) return CTC_OK;
    if (!xyz || !out) return fail(ctx, CTC_ERR_INVALID_ARGUMENT, "NULL point/output buffer");
    CK(cudaSetDevice(ctx->device));
    CK(ctx->pts_in.ensure(n * 12)); CK(ctx->pts_out.ensure(n * 4));
    CK(cudaMemcpyAsync(ctx->pts_in.p, xyz, n * 12, cudaMemcpyHostToDevice, ctx->stream));
    rc = de_batch_impl(ctx, shape, ctx->pts_in.as<float>(), n, ctx->pts_out.as<float>()); if (rc) return rc;
    CK(cudaMemcpyAsync(out, ctx->pts_out.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return CTC_OK;
}

int ctc_sample_grids_device(ctc_ctx* ctx, const ctc_shape* shape, const ctc_span* spans, size_t nspans,
                            uint32_t resolution, float* d_grids) {
    if (!ctx) return CTC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    const int rc = sample_grids_impl(ctx, shape, spans, nspans, resolution, d_grids);
    ctx->async_inflight = true;
    return rc;
}

int ctc_sample_grids(ctc_ctx* ctx, const ctc_shape* shape, const ctc_span* spans, size_t nspans, uint32_t resolution,
                     float* grids) {
    if (!ctx) return CTC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ShapeDev sh; uint32_t lg;
    int rc = check_shape(ctx, shape, &sh); if (rc) return rc;
    rc = check_spans(ctx, spans, nspans, resolution, &lg); if (rc) return rc;
    if (nspans == 0) return CTC_OK;
    if (!grids) return fail(ctx, CTC_ERR_INVALID_ARGUMENT, "grids is NULL");
    CK(cudaSetDevice(ctx->device));
    const size_t n3 = (size_t)(resolution + 1) * (resolution + 1) * (resolution + 1);
    // stage through the context's grid buffer in slices of <= 1 GiB
    size_t per = (1ull << 30) / (n3 * 4);
    if (per < 1) per = 1;
    if (per > nspans) per = nspans;
    CK(ctx->grids.ensure(per * n3 * 4));
    for (size_t s0 = 0; s0 < nspans; s0 += per) {
        const size_t cnt = (nspans - s0) < per ? (nspans - s0) : per;
        rc = sample_grids_impl(ctx, shape, spans + s0, cnt, resolution, ctx->grids.as<float>()); if (rc) return rc;
        CK(cudaMemcpyAsync(grids + s0 * n3, ctx->grids.p, cnt * n3 * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return CTC_OK;
}

int ctc_mesh_spans_device(ctc_ctx* ctx, const ctc_shape* shape, const ctc_span* spans, size_t nspans, uint32_t resolution,
                          ctc_vertex* d_v, size_t vcap, uint32_t* d_idx, size_t icap, uint64_t* d_v_off,
                          uint64_t* d_i_off) {
    if (!ctx) return CTC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->async_inflight = true;
    return guarded(ctx, [&] { return mesh_spans_impl(ctx, shape, spans, nspans, resolution, d_v, vcap, d_idx, icap, d_v_off, d_i_off); });
}

int ctc_mesh_result(ctc_ctx* ctx, uint64_t* n_vertices, uint64_t* n_indices, ctc_timings* timings) {
    if (!ctx) return CTC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    return guarded(ctx, [&] { return mesh_result_impl(ctx, n_vertices, n_indices, timings); });
}

static int mesh_spans_host_once(ctc_ctx* ctx, const ctc_shape* shape, const ctc_span* spans, size_t nspans, uint32_t resolution,
                                ctc_vertex* v, size_t vcap, uint32_t* idx, size_t icap, uint64_t* v_off, uint64_t* i_off,
                                ctc_timings* timings, bool host_wire) {
    // CANTUCCI_B200_TRACE=1: one stderr line per call with the host-side milestones of the copy pipeline (ms)
    static const bool trace = getenv("CANTUCCI_B200_TRACE") != nullptr;
    const auto t_entry = std::chrono::steady_clock::now();
    auto since = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_entry).count(); };
    double t_enq = 0, t_first = 0, t_groups = 0, t_widen = 0, t_copies = 0;
    CK(cudaSetDevice(ctx->device));
    CK(ctx->out_v.ensure((vcap ? vcap : 1) * sizeof(ctc_vertex)));
    CK(ctx->out_idx.ensure((icap ? icap : 1) * sizeof(uint32_t)));
    CK(ctx->off_v.ensure((nspans + 1) * 8)); CK(ctx->off_i.ensure((nspans + 1) * 8));
    if (!ctx->copy_stream) CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    if (!ctx->copy_stream2) CK(cudaStreamCreateWithFlags(&ctx->copy_stream2, cudaStreamNonBlocking));
    // Buffers in PAGEABLE host memory (what the reference's caller owns: a Vec) are filled through pinned landing
    // buffers at full PCIe rate, and host threads copy them on, instead of the driver's single staged pageable copy
    // (calls of a few spans keep the direct copy: less than a megabyte is not worth a hand-over between threads).
    auto pageable = [](const void* p) {
        cudaPointerAttributes a{};
        const bool un = cudaPointerGetAttributes(&a, p) == cudaSuccess && a.type == cudaMemoryTypeUnregistered;
        (void)cudaGetLastError();
        return un;
    };
    bool stage_v = nspans > 8 && v && vcap && pageable(v);
    bool stage_i = nspans > 8 && !host_wire && !ctx->wire_quads && idx && icap && pageable(idx);
    if ((host_wire || stage_v || stage_i) && !ctx->expander.start(ctx->device, ctx->expander_share)) host_wire = stage_v = stage_i = false;
    if (host_wire) { CK(ctx->h_wire.ensure((icap / 6 + 1) * sizeof(uint2))); CK(ctx->out_wire.ensure((icap / 6 + 1) * sizeof(uint2))); }
    if (stage_i && ctx->h_wire.ensure(icap * sizeof(uint32_t)) != cudaSuccess) { (void)cudaGetLastError(); stage_i = false; }
    if (stage_v && ctx->h_vstage.ensure(vcap * sizeof(ctc_vertex)) != cudaSuccess) { (void)cudaGetLastError(); stage_v = false; }
    // (hybrid only towards page-locked index buffers: a u32 group copied straight into PAGEABLE memory would be a staged copy)
    const int packed = ctx->wire_quads ? 1 : (host_wire ? ((ctx->hybrid_num >= ctx->hybrid_den || pageable(idx)) ? 1 : 2) : 0);
    ctx->last_d2h_bytes = 0;
    const bool workers = host_wire || stage_v || stage_i;
    if (workers) {
        ctx->expander.begin();
        if (host_wire) ctx->host_wire_calls++;
    }
    unsigned long long* wprog = nullptr;
    unsigned long long wtag = 0;
    const bool want_prog = ctx->wire_progress && ctx->wire_quads && !workers;
    if (want_prog) {     // (one word per launch group + the closing one; sized up front: the words are sources of async copies)
        wtag = (++ctx->wire_epoch & 0x7FFFFFull) << 40;
        CK(ctx->h_wire_prog.ensure((nspans + 2) * 8));
        wprog = static_cast<unsigned long long*>(ctx->h_wire_prog.p);
    }
    // Each launch group's slice of the mesh is copied on a second stream as soon as that group has finished, while
    // the following groups are still computing -- or still being ENQUEUED: `issue` is also called (non-blocking) from
    // the enqueue loop, so the first, small groups are on their way before the last launch has been made.
    size_t done_v = 0, done_i = 0, next_g = 0;
    int copy_rc = CTC_OK;
    auto issue = [&](size_t upto, bool blocking) -> int {
        if (workers)
            while (ctx->wire_events.size() < 2 * upto) {
                cudaEvent_t e;
                CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                ctx->wire_events.push_back(e);
            }
        for (; next_g < upto && copy_rc == CTC_OK; ++next_g) {
            const size_t g = next_g;
            cudaError_t e;
            if (blocking) e = cudaEventSynchronize(ctx->group_events[g]);
            else { e = cudaEventQuery(ctx->group_events[g]); if (e == cudaErrorNotReady) { (void)cudaGetLastError(); return CTC_OK; } }
            if (trace && g == 0) t_first = since();
            const unsigned long long tv = ctx->progress_h[2 * g], ti = 6ull * ctx->progress_h[2 * g + 1];
            const size_t cv = tv < vcap ? (size_t)tv : vcap, ci = ti < icap ? (size_t)ti : icap;
            if (e == cudaSuccess && cv > done_v) {
                ctx->last_d2h_bytes += (cv - done_v) * sizeof(ctc_vertex);
                if (stage_v) {
                    ctc_vertex* hv = static_cast<ctc_vertex*>(ctx->h_vstage.p);
                    const size_t bytes = (cv - done_v) * sizeof(ctc_vertex);
                    e = cudaMemcpyAsync(hv + done_v, ctx->out_v.as<ctc_vertex>() + done_v, bytes, cudaMemcpyDeviceToHost, ctx->copy_stream);
                    cudaEvent_t ev = ctx->wire_events[2 * g + 1];
                    if (e == cudaSuccess) e = cudaEventRecord(ev, ctx->copy_stream);
                    if (e == cudaSuccess) ctx->expander.push(ExpandTask{ev, hv + done_v, v + done_v, (bytes + 15) / 16, 1, bytes});
                } else {
                    e = cudaMemcpyAsync(v + done_v, ctx->out_v.as<ctc_vertex>() + done_v, (cv - done_v) * sizeof(ctc_vertex),
                                        cudaMemcpyDefault, ctx->copy_stream);
                }
                done_v = cv;
            }
            if (e == cudaSuccess && ci > done_i) {
                const size_t q0 = done_i / 6, q1 = ci / 6;
                if (host_wire && ctx->group_packed[g]) {   // packed records to the pinned wire buffer; host threads widen them into idx
                    uint2* hw = static_cast<uint2*>(ctx->h_wire.p);
                    ctx->last_d2h_bytes += (q1 - q0) * sizeof(uint2);
                    e = cudaMemcpyAsync(hw + q0, ctx->out_wire.as<uint2>() + q0, (q1 - q0) * sizeof(uint2), cudaMemcpyDeviceToHost, ctx->copy_stream2);
                    if (e == cudaSuccess) e = cudaEventRecord(ctx->wire_events[2 * g], ctx->copy_stream2);
                    if (e == cudaSuccess) ctx->expander.push(ExpandTask{ctx->wire_events[2 * g], hw + q0, idx + 6 * q0, q1 - q0, 0, 0});
                } else if (stage_i) {
                    uint32_t* hi = static_cast<uint32_t*>(ctx->h_wire.p);
                    const size_t bytes = (ci - done_i) * sizeof(uint32_t);
                    ctx->last_d2h_bytes += bytes;
                    e = cudaMemcpyAsync(hi + done_i, ctx->out_idx.as<uint32_t>() + done_i, bytes, cudaMemcpyDeviceToHost, ctx->copy_stream2);
                    if (e == cudaSuccess) e = cudaEventRecord(ctx->wire_events[2 * g], ctx->copy_stream2);
                    if (e == cudaSuccess) ctx->expander.push(ExpandTask{ctx->wire_events[2 * g], hi + done_i, idx + done_i, (bytes + 15) / 16, 1, bytes});
                } else if (ctx->wire_quads) {      // packed records: 8 bytes per quad (= per 6 indices), densely at quad offsets
                    e = cudaMemcpyAsync(reinterpret_cast<uint2*>(idx) + q0, ctx->out_idx.as<uint2>() + q0, (q1 - q0) * sizeof(uint2),
                                        cudaMemcpyDefault, ctx->copy_stream2);
                    if (wprog && e == cudaSuccess) {       // ... and the word that says how far the records have come
                        wprog[g] = wtag | (unsigned long long)q1;
                        e = cudaMemcpyAsync(ctx->wire_progress, wprog + g, 8, cudaMemcpyDefault, ctx->copy_stream2);
                    }
                } else {
                    ctx->last_d2h_bytes += (ci - done_i) * sizeof(uint32_t);
                    e = cudaMemcpyAsync(idx + done_i, ctx->out_idx.as<uint32_t>() + done_i, (ci - done_i) * sizeof(uint32_t),
                                        cudaMemcpyDefault, ctx->copy_stream2);
                }
                done_i = ci;
            }
            if (e != cudaSuccess) copy_rc = fail_cuda(ctx, e, "pipelined device->host copy");
        }
        return CTC_OK;
    };
    const std::function<int(size_t)> during_enqueue = [&](size_t enqueued) { return issue(enqueued, false); };
    int rc = mesh_spans_impl(ctx, shape, spans, nspans, resolution, ctx->out_v.as<ctc_vertex>(), vcap,
                             ctx->out_idx.as<uint32_t>(), icap, ctx->off_v.as<uint64_t>(), ctx->off_i.as<uint64_t>(),
                             /*pipeline=*/true, packed, &during_enqueue, host_wire ? ctx->out_wire.as<uint2>() : nullptr);
    if (rc) {
        if (workers) ctx->expander.finish();
        cudaStreamSynchronize(ctx->copy_stream); cudaStreamSynchronize(ctx->copy_stream2);
        return rc;
    }
    if (trace) t_enq = since();
    // Everything is enqueued.  The offset tables follow the kernels on the compute stream (straight
    // to the destination when it is device/peer memory, through a pinned staging buffer when it is
    // host memory, so the call never blocks on a pageable copy).
    cudaPointerAttributes pa{};
    const bool tables_on_device = cudaPointerGetAttributes(&pa, v_off) == cudaSuccess &&
                                  (pa.type == cudaMemoryTypeDevice || pa.type == cudaMemoryTypeManaged);
    (void)cudaGetLastError();
    const size_t tbytes = (nspans + 1) * 8;
    if (tables_on_device) {
        CK(cudaMemcpyAsync(v_off, ctx->off_v.p, tbytes, cudaMemcpyDefault, ctx->stream));
        CK(cudaMemcpyAsync(i_off, ctx->off_i.p, tbytes, cudaMemcpyDefault, ctx->stream));
    } else {
        CK(ctx->h_tables.ensure(2 * tbytes));
        CK(cudaMemcpyAsync(ctx->h_tables.p, ctx->off_v.p, tbytes, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(static_cast<char*>(ctx->h_tables.p) + tbytes, ctx->off_i.p, tbytes, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(ctx->h_state.ensure(sizeof(MeshState)));
    CK(cudaMemcpyAsync(ctx->h_state.p, ctx->state.p, sizeof(MeshState), cudaMemcpyDeviceToHost, ctx->stream));
    {
        const int irc = issue(ctx->n_groups, true);        // the rest, blocking on each group's event
        if (irc != CTC_OK && copy_rc == CTC_OK) copy_rc = irc;
    }
    if (wprog && copy_rc == CTC_OK) {        // the last word: everything of this call is there
        wprog[ctx->n_groups] = (1ull << 63) | wtag | (unsigned long long)(done_i / 6);
        if (cudaMemcpyAsync(ctx->wire_progress, wprog + ctx->n_groups, 8, cudaMemcpyDefault, ctx->copy_stream2) != cudaSuccess)
            copy_rc = fail_cuda(ctx, cudaGetLastError(), "wire progress");
    }
    if (trace) t_groups = since();
    if (workers) ctx->expander.finish();         // (always: the workers must be idle before the landing buffers are reused)
    if (trace) t_widen = since();
    if (copy_rc != CTC_OK) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamSynchronize(ctx->copy_stream2); return copy_rc; }
    CK(cudaStreamSynchronize(ctx->copy_stream));
    CK(cudaStreamSynchronize(ctx->copy_stream2));
    if (trace) t_copies = since();
    uint64_t nv = 0, ni = 0;
    const int status = mesh_result_impl(ctx, &nv, &ni, timings, /*state_already_copied=*/true);   // syncs the compute stream
    if (trace)
        fprintf(stderr, "[ctc trace] %zu spans, %zu groups: enqueued %.2f, first group %.2f, last group %.2f, host threads done %.2f, "
                        "copies done %.2f, end %.2f ms (host wire %d)\n", nspans, ctx->n_groups, t_enq, t_first, t_groups, t_widen,
                t_copies, since(), (int)host_wire);
    if (status != CTC_ERR_CUDA && !tables_on_device) {
        memcpy(v_off, ctx->h_tables.p, tbytes);
        memcpy(i_off, static_cast<char*>(ctx->h_tables.p) + tbytes, tbytes);
    }
    return status;
}

static int mesh_spans_host(ctc_ctx* ctx, const ctc_shape* shape, const ctc_span* spans, size_t nspans, uint32_t resolution,
                           ctc_vertex* v, size_t vcap, uint32_t* idx, size_t icap, uint64_t* v_off, uint64_t* i_off,
                           ctc_timings* timings) {
    if (!v_off || !i_off) return fail(ctx, CTC_ERR_INVALID_ARGUMENT, "offset tables are NULL");
    if ((vcap && !v) || (icap && !idx)) return fail(ctx, CTC_ERR_INVALID_ARGUMENT, "NULL output buffer with non-zero capacity");
    // Index wire for HOST destinations: packed 8-byte quad records over PCIe, widened by host threads.
    // (calls of fewer than 128 spans are not PCIe-bound: they keep six u32 per quad on the wire)
    bool host_wire = ctx->host_wire && !ctx->wire_quads && idx != nullptr && icap >= 6 && (nspans >= 128 || ctx->host_wire == 2);
    if (host_wire) {
        cudaPointerAttributes pa{};
        if (cudaPointerGetAttributes(&pa, idx) == cudaSuccess && (pa.type == cudaMemoryTypeDevice || pa.type == cudaMemoryTypeManaged))
            host_wire = false;
        (void)cudaGetLastError();
    }
    int rc = mesh_spans_host_once(ctx, shape, spans, nspans, resolution, v, vcap, idx, icap, v_off, i_off, timings, host_wire);
    if (host_wire && rc == CTC_ERR_OVERFLOW && ctx->last_wire_overflow) {
        // a span with >= 65536 vertices does not fit the 16-bit records: once more with u32 indices on the wire
        ctx->host_wire_fallbacks++;
        rc = mesh_spans_host_once(ctx, shape, spans, nspans, resolution, v, vcap, idx, icap, v_off, i_off, timings, false);
    }
    return rc;
}

}  // extern "C" (reopened below)

// A queued host-pointer mesh call (see ctc_ctx::q).
struct MeshRequest {
    const ctc_shape* shape; const ctc_span* spans; size_t nspans; uint32_t resolution;
    ctc_vertex* v; size_t vcap; uint32_t* idx; size_t icap; uint64_t* v_off; uint64_t* i_off; ctc_timings* timings;
    int status = CTC_OK;
    std::string err;
    bool done = false;
};

namespace {

bool same_shape(const ctc_shape& a, const ctc_shape& b) {
    return a.kind == b.kind && a.power == b.power && a.max_iters == b.max_iters && a.flags == b.flags &&
           memcmp(&a.bailout, &b.bailout, sizeof(float)) == 0 && memcmp(a.center, b.center, sizeof a.center) == 0 &&
           memcmp(&a.radius, &b.radius, sizeof(float)) == 0;
}

int run_request(ctc_ctx* ctx, MeshRequest* r) {
    return guarded(ctx, [&] {
        return mesh_spans_host(ctx, r->shape, r->spans, r->nspans, r->resolution, r->v, r->vcap, r->idx, r->icap, r->v_off,
                               r->i_off, r->timings);
    });
}

// One batched launch for several queued requests of the same shape and resolution: the spans are
// concatenated, the batch is meshed into the context's pinned staging buffers, and every request gets its
// slice (offset tables rebased to its own buffers).  Anything but a clean batch -- an error status, or a
// request whose buffers are too small -- falls back to running the requests one by one, which reproduces
// the per-call error contract exactly.
void run_batch(ctc_ctx* ctx, std::vector<MeshRequest*>& batch) {
    std::lock_guard<std::mutex> lk(ctx->mu);
    auto one_by_one = [&] {
        for (MeshRequest* r : batch) { r->status = run_request(ctx, r); r->err = ctx->err; }
    };
    if (batch.size() == 1) { one_by_one(); return; }
    size_t nspans = 0, vcap = 0, icap = 0;
    for (MeshRequest* r : batch) { nspans += r->nspans; vcap += r->vcap; icap += r->icap; }
    const int rc = guarded(ctx, [&]() -> int {
        ctx->b_spans.clear();
        for (MeshRequest* r : batch) ctx->b_spans.insert(ctx->b_spans.end(), r->spans, r->spans + r->nspans);
        ctx->b_voff.assign(nspans + 1, 0); ctx->b_ioff.assign(nspans + 1, 0);
        CK(cudaSetDevice(ctx->device));
        CK(ctx->b_v.ensure((vcap ? vcap : 1) * sizeof(ctc_vertex)));
        CK(ctx->b_i.ensure((icap ? icap : 1) * sizeof(uint32_t)));
        ctc_timings t{};
        const int st = mesh_spans_host(ctx, batch[0]->shape, ctx->b_spans.data(), nspans, batch[0]->resolution,
                                       static_cast<ctc_vertex*>(ctx->b_v.p), vcap, static_cast<uint32_t*>(ctx->b_i.p), icap,
                                       ctx->b_voff.data(), ctx->b_ioff.data(), &t);
        if (st != CTC_OK) return st;
        size_t s0 = 0;
        for (MeshRequest* r : batch) {      // does every request's slice fit its own buffers?
            const uint64_t nv = ctx->b_voff[s0 + r->nspans] - ctx->b_voff[s0], ni = ctx->b_ioff[s0 + r->nspans] - ctx->b_ioff[s0];
            if (nv > r->vcap || ni > r->icap) return CTC_ERR_OVERFLOW;
            s0 += r->nspans;
        }
        s0 = 0;
        const double share = 1.0 / (double)(nspans ? nspans : 1);
        for (MeshRequest* r : batch) {
            const uint64_t v0 = ctx->b_voff[s0], i0 = ctx->b_ioff[s0];
            const uint64_t nv = ctx->b_voff[s0 + r->nspans] - v0, ni = ctx->b_ioff[s0 + r->nspans] - i0;
            if (nv) memcpy(r->v, static_cast<ctc_vertex*>(ctx->b_v.p) + v0, nv * sizeof(ctc_vertex));
            if (ni) memcpy(r->idx, static_cast<uint32_t*>(ctx->b_i.p) + i0, ni * sizeof(uint32_t));
            for (size_t k = 0; k <= r->nspans; ++k) { r->v_off[k] = ctx->b_voff[s0 + k] - v0; r->i_off[k] = ctx->b_ioff[s0 + k] - i0; }
            if (r->timings) {   // the batch's kernels served every request: device time is apportioned by span count
                const double f = share * (double)r->nspans;
                r->timings->first_ms = t.first_ms * f; r->timings->second_ms = t.second_ms * f; r->timings->third_ms = t.third_ms * f;
                r->timings->vertices = nv; r->timings->faces = ni / 6;
            }
            r->status = CTC_OK;
            s0 += r->nspans;
        }
        ctx->batches++; ctx->batched_requests += batch.size();
        return CTC_OK;
    });
    if (rc != CTC_OK) one_by_one();
}

}  // namespace

extern "C" {

int ctc_mesh_spans(ctc_ctx* ctx, const ctc_shape* shape, const ctc_span* spans, size_t nspans, uint32_t resolution,
                   ctc_vertex* v, size_t vcap, uint32_t* idx, size_t icap, uint64_t* v_off, uint64_t* i_off,
                   ctc_timings* timings) {
    if (!ctx) return CTC_ERR_INVALID_ARGUMENT;
    MeshRequest req{shape, spans, nspans, resolution, v, vcap, idx, icap, v_off, i_off, timings};
    // Destinations the device copies to directly (peer / device memory) and large calls gain nothing from
    // coalescing: they go straight through.
    bool direct = !ctx->coalesce || !shape || !v_off || !i_off || nspans == 0 || nspans > 64;
    if (!direct) {
        cudaPointerAttributes pa{};
        if (cudaPointerGetAttributes(&pa, v ? static_cast<const void*>(v) : static_cast<const void*>(v_off)) == cudaSuccess &&
            (pa.type == cudaMemoryTypeDevice || pa.type == cudaMemoryTypeManaged))
            direct = true;
        (void)cudaGetLastError();
    }
    if (direct) {
        std::lock_guard<std::mutex> lk(ctx->mu);
        return run_request(ctx, &req);
    }
    std::unique_lock<std::mutex> ql(ctx->q_mu);
    ctx->q.push_back(&req);
    if (ctx->q_leader) {                      // somebody is launching: wait for a leader to serve this request
        ctx->q_cv.wait(ql, [&] { return req.done; });
        if (req.status != CTC_OK) { std::lock_guard<std::mutex> lk(ctx->mu); ctx->err = req.err; }
        return req.status;
    }
    ctx->q_leader = true;                     // lead: serve the queue until it is empty (own request included)
    while (!ctx->q.empty()) {
        std::vector<MeshRequest*> batch;
        size_t total = 0;
        for (size_t k = 0; k < ctx->q.size();) {          // everything compatible with the queue's head, up to 4096 spans
            MeshRequest* r = ctx->q[k];
            if (batch.empty() || (r->resolution == batch[0]->resolution && same_shape(*r->shape, *batch[0]->shape) && total + r->nspans <= 4096)) {
                batch.push_back(r); total += r->nspans;
                ctx->q.erase(ctx->q.begin() + (long)k);
            } else ++k;
        }
        ql.unlock();
        run_batch(ctx, batch);
        ql.lock();
        for (MeshRequest* r : batch) r->done = true;
        ctx->q_cv.notify_all();
    }
    ctx->q_leader = false;
    if (req.status != CTC_OK) { std::lock_guard<std::mutex> lk(ctx->mu); ctx->err = req.err; }
    return req.status;
}

int ctc_ctx_set_coalescing(ctc_ctx* ctx, int enable) {
    if (!ctx) return CTC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->q_mu);
    ctx->coalesce = enable != 0;
    return CTC_OK;
}

int ctc_ctx_coalescing_stats(ctc_ctx* ctx, uint64_t* batches, uint64_t* requests) {
    if (!ctx) return CTC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (batches) *batches = ctx->batches;
    if (requests) *requests = ctx->batched_requests;
    return CTC_OK;
}

}  // extern "C"

namespace {

// DE at every span's centre (exact arithmetic) and the span's reach = half-diagonal of its skirt-expanded box
// (buffer.rs:64-67): what span culling and the surface-first order are computed from.  Through a pinned buffer, one
// kernel, one synchronisation.  d and reach point into ctx->h_pts.
int centre_distances(ctc_ctx* ctx, const ctc_shape* shape, const ctc_span* spans, size_t nspans, uint32_t resolution,
                     const float** d_out, const float** reach_out) {
    CK(cudaSetDevice(ctx->device));
    CK(ctx->h_pts.ensure(nspans * 20));
    float* centres = static_cast<float*>(ctx->h_pts.p);
    float* d = centres + 3 * nspans;
    float* reach = d + nspans;
    for (size_t i = 0; i < nspans; ++i) {
        double r2 = 0.0;
        for (int c = 0; c < 3; ++c) {
            const double ext = (double)spans[i].end[c] - (double)spans[i].start[c];
            const double half = 0.5 * ext + ext / (double)resolution;       // half extent + the one-cell skirt (buffer.rs:64-67)
            centres[3 * i + c] = (float)(0.5 * ((double)spans[i].start[c] + (double)spans[i].end[c]));
            r2 += half * half;
        }
        reach[i] = (float)(std::sqrt(r2) * 1.0001);           // (rounding of the centre and of the sample positions)
    }
    ctc_shape exact = *shape;
    exact.flags &= ~(uint32_t)CTC_MATH_FAST;
    CK(ctx->pts_in.ensure(nspans * 12)); CK(ctx->pts_out.ensure(nspans * 4));
    CK(cudaMemcpyAsync(ctx->pts_in.p, centres, nspans * 12, cudaMemcpyHostToDevice, ctx->stream));
    const int rc = de_batch_impl(ctx, &exact, ctx->pts_in.as<float>(), nspans, ctx->pts_out.as<float>()); if (rc) return rc;
    CK(cudaMemcpyAsync(d, ctx->pts_out.p, nspans * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *d_out = d; *reach_out = reach;
    return CTC_OK;
}

}  // namespace

extern "C" {

// DE-bound span culling.  The distance estimate at a span's centre bounds how close the surface can
// be (Shape::min_distance_from, shape/mod.rs:26-37: "a lower bound of the distance"); a span whose skirt-expanded
// box lies inside that ball has an all-positive sample grid and an empty mesh.  Shapes with an upper bound as well
// (Sphere, sphere.rs:37-39) also drop the spans that lie entirely inside.  keep[i] = 0: span i needs no meshing.
int ctc_cull_spans(ctc_ctx* ctx, const ctc_shape* shape, const ctc_span* spans, size_t nspans, uint32_t resolution,
                   float safety, uint8_t* keep) {
    if (!ctx) return CTC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    return guarded(ctx, [&]() -> int {
        ShapeDev sh; uint32_t lg;
        int rc = check_shape(ctx, shape, &sh); if (rc) return rc;
        rc = check_spans(ctx, spans, nspans, resolution, &lg); if (rc) return rc;
        if (nspans == 0) return CTC_OK;
        if (!keep) return fail(ctx, CTC_ERR_INVALID_ARGUMENT, "keep is NULL");
        if (!(safety >= 1.0f)) return fail(ctx, CTC_ERR_INVALID_ARGUMENT, "safety must be >= 1");
        const float *d, *reach;
        rc = centre_distances(ctx, shape, spans, nspans, resolution, &d, &reach); if (rc) return rc;
        const bool two_sided = shape->kind == CTC_SHAPE_SPHERE;         // max_distance_from is Some(..) (sphere.rs:37)
        for (size_t i = 0; i < nspans; ++i) {
            const bool outside = d[i] > safety * reach[i];
            const bool inside = two_sided && -d[i] > reach[i];
            keep[i] = (outside || inside) ? 0 : 1;                       // (NaN compares false: kept)
        }
        return CTC_OK;
    });
}

// Cost-aware span order ("surface spans are 10-50x more expensive than empty ones").  order[k] = index of
// the span to mesh k-th: ascending |DE(centre)| / reach, i.e. the spans most likely to hold surface first and the
// provably empty ones last (stable: equal keys keep the caller's order; NaN counts as 0).  A call made in this order
// produces its mesh bytes EARLY, so the copy pipeline behind the launch groups (device->host over PCIe, or the puts
// of the multi-GPU gather over NVLink) is busy from the first group on and the groups computed last -- empty space --
// leave nothing to copy after the kernels end.  The meshes themselves do not depend on the order.
int ctc_order_spans(ctc_ctx* ctx, const ctc_shape* shape, const ctc_span* spans, size_t nspans, uint32_t resolution,
                    uint32_t* order) {
    if (!ctx) return CTC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    return guarded(ctx, [&]() -> int {
        ShapeDev sh; uint32_t lg;
        int rc = check_shape(ctx, shape, &sh); if (rc) return rc;
        rc = check_spans(ctx, spans, nspans, resolution, &lg); if (rc) return rc;
        if (nspans == 0) return CTC_OK;
        if (!order) return fail(ctx, CTC_ERR_INVALID_ARGUMENT, "order is NULL");
        if (nspans > 0xFFFFFFFFull) return fail(ctx, CTC_ERR_INVALID_ARGUMENT, "more than 2^32 - 1 spans");
        const float *d, *reach;
        rc = centre_distances(ctx, shape, spans, nspans, resolution, &d, &reach); if (rc) return rc;
        // keys are non-negative floats: their bit patterns sort like unsigned integers (LSD radix sort, 3 x 11 bits)
        std::vector<uint32_t> key(nspans), tmp_key(nspans), tmp_idx(nspans);
        for (size_t i = 0; i < nspans; ++i) {
            float k = std::fabs(d[i]) / reach[i];
            if (!(k == k)) k = 0.0f;
            memcpy(&key[i], &k, 4);
            order[i] = (uint32_t)i;
        }
        uint32_t* ka = key.data(); uint32_t* kb = tmp_key.data();
        uint32_t* ia = order;      uint32_t* ib = tmp_idx.data();
        for (int pass = 0; pass < 3; ++pass) {
            const int shift = 11 * pass;
            size_t hist[2049] = {0};
            for (size_t i = 0; i < nspans; ++i) hist[((ka[i] >> shift) & 2047u) + 1]++;
            for (int b = 0; b < 2048; ++b) hist[b + 1] += hist[b];
            for (size_t i = 0; i < nspans; ++i) { const size_t at = hist[(ka[i] >> shift) & 2047u]++; kb[at] = ka[i]; ib[at] = ia[i]; }
            std::swap(ka, kb); std::swap(ia, ib);
        }
        if (ia != order) memcpy(order, ia, nspans * sizeof(uint32_t));    // (three passes: the result sits in the scratch array)
        return CTC_OK;
    });
}

int ctc_ray_march(ctc_ctx* ctx, const ctc_shape* shape, const float* origin, const float* dir, size_t n,
                  uint32_t max_steps, float epsilon, float* out_pos, uint32_t* out_hit) {
    if (!ctx) return CTC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ShapeDev sh;
    int rc = check_shape(ctx, shape, &sh); if (rc) return rc;
    if (n == 0) return CTC_OK;
    if (!origin || !dir || !out_pos || !out_hit) return fail(ctx, CTC_ERR_INVALID_ARGUMENT, "NULL ray buffer");
    CK(cudaSetDevice(ctx->device));
    CK(ctx->pts_in.ensure(n * 24)); CK(ctx->pts_out.ensure(n * 16));
    float* d_o = ctx->pts_in.as<float>(); float* d_d = d_o + 3 * n;
    float* d_p = ctx->pts_out.as<float>(); uint32_t* d_h = reinterpret_cast<uint32_t*>(d_p + 3 * n);
    CK(cudaMemcpyAsync(d_o, origin, n * 12, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_d, dir, n * 12, cudaMemcpyHostToDevice, ctx->stream));
    const bool fast = (shape->flags & CTC_MATH_FAST) != 0;
    const int variant = shape_variant(shape);
    const unsigned blocks = (unsigned)((n + kThreads - 1) / kThreads);
#define CALL(F, V) ray_march_kernel<F, V><<<blocks, kThreads, 0, ctx->stream>>>(sh, d_o, d_d, n, max_steps, epsilon, d_p, d_h)
    DISPATCH(fast, variant, CALL);
#undef CALL
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out_pos, d_p, n * 12, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(out_hit, d_h, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return CTC_OK;
}

// N4: sphere-traced image of the shape, one thread per pixel (render_kernel).
static int render_impl(ctc_ctx* ctx, const ctc_shape* shape, const ctc_camera_rays* cam, uint32_t width, uint32_t height,
                       uint32_t max_steps, float epsilon, float* out, bool out_on_device) {
    ShapeDev sh;
    int rc = check_shape(ctx, shape, &sh); if (rc) return rc;
    if (!cam || !out) return fail(ctx, CTC_ERR_INVALID_ARGUMENT, "NULL camera / output");
    if (width == 0 || height == 0) return CTC_OK;
    if ((size_t)width * height > ((size_t)1 << 28)) return fail(ctx, CTC_ERR_INVALID_ARGUMENT, "image too large");
    if (out_on_device && (reinterpret_cast<uintptr_t>(out) & 15u)) return fail(ctx, CTC_ERR_INVALID_ARGUMENT, "output must be 16-byte aligned");
    CK(cudaSetDevice(ctx->device));
    const size_t n = (size_t)width * height;
    float4* d_out = reinterpret_cast<float4*>(out);
    if (!out_on_device) { CK(ctx->pts_out.ensure(n * sizeof(float4))); d_out = ctx->pts_out.as<float4>(); }
    CameraRays cr;
    memcpy(cr.eye, cam->eye, 12); memcpy(cr.top_left, cam->top_left, 12); memcpy(cr.du, cam->du, 12); memcpy(cr.dv, cam->dv, 12);
    const bool fast = (shape->flags & CTC_MATH_FAST) != 0;
    const int variant = shape_variant(shape);
    const dim3 grid((width + 15) / 16, (height + 15) / 16);
#define CALL(F, V) render_kernel<F, V><<<grid, kThreads, 0, ctx->stream>>>(sh, cr, width, height, max_steps, epsilon, d_out)
    DISPATCH(fast, variant, CALL);
#undef CALL
    ctx->launches++;
    CK(cudaGetLastError());
    if (!out_on_device) {
        CK(cudaMemcpyAsync(out, d_out, n * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    } else {
        ctx->async_inflight = true;
    }
    return CTC_OK;
}

int ctc_render(ctc_ctx* ctx, const ctc_shape* shape, const ctc_camera_rays* cam, uint32_t width, uint32_t height,
               uint32_t max_steps, float epsilon, float* out) {
    if (!ctx) return CTC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    return guarded(ctx, [&] { return render_impl(ctx, shape, cam, width, height, max_steps, epsilon, out, false); });
}

int ctc_render_device(ctc_ctx* ctx, const ctc_shape* shape, const ctc_camera_rays* cam, uint32_t width, uint32_t height,
                      uint32_t max_steps, float epsilon, float* d_out) {
    if (!ctx) return CTC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    return guarded(ctx, [&] { return render_impl(ctx, shape, cam, width, height, max_steps, epsilon, d_out, true); });
}

int ctc_device_alloc(ctc_ctx* ctx, size_t bytes, void** d_ptr) {
    if (!ctx || !d_ptr) return CTC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    CK(cudaMalloc(d_ptr, bytes ? bytes : 1));
    return CTC_OK;
}

int ctc_device_free(ctc_ctx* ctx, void* d_ptr) {
    if (!ctx) return CTC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    CK(cudaFree(d_ptr));
    return CTC_OK;
}

int ctc_ipc_export(ctc_ctx* ctx, const void* d_ptr, unsigned char handle[64]) {
    if (!ctx || !d_ptr || !handle) return CTC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    CK(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, const_cast<void*>(d_ptr)));
    memcpy(handle, &h, 64);
    return CTC_OK;
}

int ctc_ipc_open(ctc_ctx* ctx, const unsigned char handle[64], void** d_ptr) {
    if (!ctx || !d_ptr || !handle) return CTC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CK(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return CTC_OK;
}

int ctc_ipc_close(ctc_ctx* ctx, void* d_ptr) {
    if (!ctx) return CTC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    CK(cudaIpcCloseMemHandle(d_ptr));
    return CTC_OK;
}

int ctc_host_register(ctc_ctx* ctx, void* ptr, size_t bytes) {
    if (!ctx || !ptr) return CTC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    CK(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    return CTC_OK;
}

int ctc_host_unregister(ctc_ctx* ctx, void* ptr) {
    if (!ctx || !ptr) return CTC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    CK(cudaHostUnregister(ptr));
    return CTC_OK;
}

int ctc_iteration_stats(ctc_ctx* ctx, const ctc_shape* shape, const ctc_span* spans, size_t nspans, uint32_t resolution,
                        uint64_t out[3]) {
    if (!ctx || !out) return CTC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ShapeDev sh; uint32_t lg;
    int rc = check_shape(ctx, shape, &sh); if (rc) return rc;
    rc = check_spans(ctx, spans, nspans, resolution, &lg); if (rc) return rc;
    out[0] = out[1] = out[2] = 0;
    if (nspans == 0) return CTC_OK;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    rc = upload_geom(ctx, spans, nspans, resolution); if (rc) return rc;
    CK(ctx->pts_out.ensure(3 * sizeof(unsigned long long)));
    CK(cudaMemsetAsync(ctx->pts_out.p, 0, 3 * sizeof(unsigned long long), ctx->stream));
    const size_t n3 = (size_t)(resolution + 1) * (resolution + 1) * (resolution + 1);
    const int variant = shape_variant(shape);
    for (size_t s0 = 0; s0 < nspans; s0 += 32768) {
        const uint32_t cnt = (uint32_t)((nspans - s0) < 32768 ? (nspans - s0) : 32768);
        dim3 grid((unsigned)((n3 + kThreads - 1) / kThreads), cnt);
        const SpanGeom* geom = ctx->geom.as<SpanGeom>() + s0;
        unsigned long long* o = ctx->pts_out.as<unsigned long long>();
        const float inv_r = 1.0f / (float)resolution;
        if (variant == kVarP8) iteration_stats_kernel<kVarP8><<<grid, kThreads, 0, ctx->stream>>>(sh, geom, resolution, lg, inv_r, o);
        else if (variant == kVarGeneric) iteration_stats_kernel<kVarGeneric><<<grid, kThreads, 0, ctx->stream>>>(sh, geom, resolution, lg, inv_r, o);
        else iteration_stats_kernel<kVarSphere><<<grid, kThreads, 0, ctx->stream>>>(sh, geom, resolution, lg, inv_r, o);
        ctx->launches++;
    }
    CK(cudaGetLastError());
    unsigned long long h[3];
    CK(cudaMemcpyAsync(h, ctx->pts_out.p, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    out[0] = h[0]; out[1] = h[1]; out[2] = h[2];
    return CTC_OK;
}

int ctc_iteration_stats_points(ctc_ctx* ctx, const ctc_shape* shape, const float* d_xyz, size_t n, uint64_t out[3]) {
    if (!ctx || !out) return CTC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ShapeDev sh;
    int rc = check_shape(ctx, shape, &sh); if (rc) return rc;
    out[0] = out[1] = out[2] = 0;
    if (n == 0) return CTC_OK;
    if (!d_xyz) return fail(ctx, CTC_ERR_INVALID_ARGUMENT, "NULL point buffer");
    CK(cudaSetDevice(ctx->device));
    CK(ctx->pts_out.ensure(3 * sizeof(unsigned long long)));
    CK(cudaMemsetAsync(ctx->pts_out.p, 0, 3 * sizeof(unsigned long long), ctx->stream));
    const unsigned blocks = (unsigned)((n + kThreads - 1) / kThreads);
    unsigned long long* o = ctx->pts_out.as<unsigned long long>();
    const int variant = shape_variant(shape);
    if (variant == kVarP8) iteration_stats_points_kernel<kVarP8><<<blocks, kThreads, 0, ctx->stream>>>(sh, d_xyz, n, o);
    else if (variant == kVarGeneric) iteration_stats_points_kernel<kVarGeneric><<<blocks, kThreads, 0, ctx->stream>>>(sh, d_xyz, n, o);
    else iteration_stats_points_kernel<kVarSphere><<<blocks, kThreads, 0, ctx->stream>>>(sh, d_xyz, n, o);
    ctx->launches++;
    CK(cudaGetLastError());
    unsigned long long h[3];
    CK(cudaMemcpyAsync(h, ctx->pts_out.p, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    out[0] = h[0]; out[1] = h[1]; out[2] = h[2];
    return CTC_OK;
}

int ctc_ctx_set_fast_band(ctc_ctx* ctx, float kappa) {
    if (!ctx) return CTC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!(kappa >= 0.0f)) return fail(ctx, CTC_ERR_INVALID_ARGUMENT, "kappa must be >= 0");
    ctx->kappa = kappa > 0.0f ? kappa : kDefaultKappa;
    return CTC_OK;
}

int ctc_mesh_fixups(ctc_ctx* ctx, uint64_t* suspects, uint64_t* sign_fixups) {
    if (!ctx) return CTC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!ctx->h_state.p) return fail(ctx, CTC_ERR_INVALID_ARGUMENT, "no mesh result has been fetched on this context");
    const MeshState* st = static_cast<const MeshState*>(ctx->h_state.p);
    if (suspects) *suspects = st->suspects;
    if (sign_fixups) *sign_fixups = st->sign_fixups;
    return CTC_OK;
}

int ctc_sample_signs(ctc_ctx* ctx, const ctc_shape* shape, const ctc_span* spans, size_t nspans, uint32_t resolution,
                     uint32_t* planes) {
    if (!ctx) return CTC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ShapeDev sh; uint32_t lg;
    int rc = check_shape(ctx, shape, &sh); if (rc) return rc;
    rc = check_spans(ctx, spans, nspans, resolution, &lg); if (rc) return rc;
    if (nspans == 0) return CTC_OK;
    if (!planes) return fail(ctx, CTC_ERR_INVALID_ARGUMENT, "planes is NULL");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    rc = upload_geom(ctx, spans, nspans, resolution); if (rc) return rc;
    const size_t n3 = (size_t)(resolution + 1) * (resolution + 1) * (resolution + 1);
    const size_t words = (n3 + 31) / 32;
    const uint32_t sign_stride = (uint32_t)(words + 1);
    size_t per = (1ull << 30) / (n3 * 4);
    if (per < 1) per = 1;
    if (per > nspans) per = nspans;
    if (per > 32768) per = 32768;
    CK(ctx->grids.ensure(per * n3 * 4));
    CK(ctx->sign_bits.ensure(per * (size_t)sign_stride * 4));
    const bool fast = (shape->flags & CTC_MATH_FAST) != 0;
    const int variant = shape_variant(shape);
    const bool listed = fast && variant == kVarP8;
    SuspectList sl{nullptr, nullptr, 0u};
    if (listed) {
        const size_t cap = suspect_cap(per, n3);
        CK(ctx->suspects.ensure(cap * sizeof(uint2)));
        CK(ctx->suspect_count.ensure(2 * sizeof(unsigned int)));
        sl = SuspectList{ctx->suspects.as<uint2>(), ctx->suspect_count.as<unsigned int>(), (uint32_t)cap};
    }
    for (size_t s0 = 0; s0 < nspans; s0 += per) {
        const uint32_t cnt = (uint32_t)((nspans - s0) < per ? (nspans - s0) : per);
        const SpanGeom* geom = ctx->geom.as<SpanGeom>() + s0;
        CK(cudaMemsetAsync(ctx->sign_bits.p, 0, (size_t)cnt * sign_stride * 4, ctx->stream));
        if (listed) CK(cudaMemsetAsync(sl.count, 0, sizeof(unsigned int), ctx->stream));
#define CALL(F, V) launch_sample<F, V>(ctx, sh, geom, resolution, lg, ctx->grids.as<float>(), n3, cnt, n3, ctx->sign_bits.as<uint32_t>(), sign_stride, sl, ctx->stream)
        DISPATCH(fast, variant, CALL);
#undef CALL
        if (listed)
            launch_fixup(ctx, sh, variant, geom, resolution, ctx->grids.as<float>(), n3, ctx->sign_bits.as<uint32_t>(), sign_stride, sl, nullptr, ctx->stream);
        CK(cudaGetLastError());
        CK(cudaMemcpy2DAsync(planes + s0 * words, words * 4, ctx->sign_bits.p, (size_t)sign_stride * 4, words * 4, cnt,
                             cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return CTC_OK;
}

int ctc_fast_sign_probe(ctc_ctx* ctx, const ctc_shape* shape, const ctc_span* spans, size_t nspans, uint32_t resolution,
                        uint64_t* out, size_t out_words) {
    if (!ctx || !out) return CTC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ShapeDev sh; uint32_t lg;
    int rc = check_shape(ctx, shape, &sh); if (rc) return rc;
    rc = check_spans(ctx, spans, nspans, resolution, &lg); if (rc) return rc;
    if (shape_variant(shape) != kVarP8) return fail(ctx, CTC_ERR_INVALID_ARGUMENT, "the sign probe covers the power-8 path");
    constexpr size_t kDumpWords = (size_t)kProbeDump * 6;      // 12 floats per record
    constexpr size_t kTotal = (size_t)kProbeWords + 1 + kDumpWords;
    if (out_words < kTotal) return fail(ctx, CTC_ERR_INVALID_ARGUMENT, "out too small (ctc_fast_sign_probe needs 489 words)");
    for (size_t i = 0; i < out_words; ++i) out[i] = 0;
    if (nspans == 0) return CTC_OK;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    rc = upload_geom(ctx, spans, nspans, resolution); if (rc) return rc;
    CK(ctx->pts_out.ensure(kTotal * sizeof(unsigned long long)));
    CK(cudaMemsetAsync(ctx->pts_out.p, 0, kTotal * sizeof(unsigned long long), ctx->stream));
    unsigned long long* d_out = ctx->pts_out.as<unsigned long long>();
    unsigned int* d_cnt = reinterpret_cast<unsigned int*>(d_out + kProbeWords);
    float* d_dump = reinterpret_cast<float*>(d_out + kProbeWords + 1);
    const size_t n3 = (size_t)(resolution + 1) * (resolution + 1) * (resolution + 1);
    for (size_t s0 = 0; s0 < nspans; s0 += 32768) {
        const uint32_t cnt = (uint32_t)((nspans - s0) < 32768 ? (nspans - s0) : 32768);
        dim3 grid((unsigned)((n3 + kThreads - 1) / kThreads), cnt);
        fast_sign_probe_kernel<<<grid, kThreads, 0, ctx->stream>>>(sh, ctx->geom.as<SpanGeom>() + s0, resolution, lg,
                                                                   1.0f / (float)resolution, d_out, d_dump, d_cnt);
        ctx->launches++;
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, ctx->pts_out.p, kTotal * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return CTC_OK;
}

int ctc_fp32_peak_probe(ctc_ctx* ctx, double* tflops, int* num_sms) {
    if (!ctx || !tflops) return CTC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    const unsigned blocks = (unsigned)ctx->num_sms * 8u;
    const uint32_t iters = 4096;
    CK(ctx->pts_out.ensure((size_t)blocks * kThreads * sizeof(float)));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        CK(cudaEventRecord(a, ctx->stream));
        fma_peak_kernel<<<blocks, kThreads, 0, ctx->stream>>>(ctx->pts_out.as<float>(), iters);
        ctx->launches++;
        CK(cudaEventRecord(b, ctx->stream));
        CK(cudaEventSynchronize(b));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, a, b));
        const double flops = 2.0 * 8 * 16 * (double)iters * blocks * kThreads;
        if (rep >= 2 && ms > 0.f) { const double t = flops / (ms * 1e-3) / 1e12; if (t > best) best = t; }
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    *tflops = best;
    if (num_sms) *num_sms = ctx->num_sms;
    return CTC_OK;
}

}  // extern "C"

#include "multi.inc"
#include "interop.inc"
