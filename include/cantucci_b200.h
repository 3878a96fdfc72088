/*
 * cantucci_b200.h -- C ABI of the B200-native cantucci hot path.
 *
 * The reference (LukasKalbertodt/cantucci, Rust) has no FFI layer; the seam
 * this library plugs into is the `Shape` trait plus one associated function:
 *
 *   trait Shape::min_distance_from / batch_min_distance_from
 *                                   src/shape/mod.rs:37, :89  (impl src/shape/mandelbulb.rs:59-79)
 *   MeshBuffer::generate_for_box    src/mesh/buffer.rs:30-42  (called from src/mesh/mod.rs:141-148)
 *   mesh::Vertex / index buffers    src/mesh/mod.rs:255-261, consumed by src/mesh/view.rs:23-41
 *
 * Every entry point is `extern "C"`, takes plain pointers and sizes, and
 * returns an int status (the reference panics; the Rust shim in
 * INTEGRATION.md turns a non-zero status back into a panic).  There is no CPU
 * fallback: without a CUDA device every call fails with CTC_ERR_CUDA /
 * CTC_ERR_NO_DEVICE.
 *
 * Threading: a ctc_ctx may be used from any thread (Shape: Sync + Send).
 * Concurrent ctc_mesh_spans calls on ONE context (the reference meshes one
 * leaf per thread-pool job, src/mesh/mod.rs:141-148) are coalesced by an
 * internal submission queue into batched launches; every other entry point
 * serialises on the context's mutex.
 */
#ifndef CANTUCCI_B200_H
#define CANTUCCI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTC_VERSION 100

#if defined(__GNUC__)
#define CTC_API __attribute__((visibility("default")))
#else
#define CTC_API
#endif

/* ---- status codes ------------------------------------------------------ */
enum {
    CTC_OK = 0,
    /* the reference's argument asserts: span.start < span.end per axis,
     * resolution a non-zero power of two (src/mesh/buffer.rs:35-39),
     * GridTable size >= 2 (src/util/grid.rs:25), max_iters >= 1
     * (src/shape/mandelbulb.rs:20) */
    CTC_ERR_INVALID_ARGUMENT = 1,
    CTC_ERR_CUDA = 2,
    /* output capacity too small; required totals are still written to
     * v_off[nspans] / i_off[nspans] so the caller can re-allocate and retry */
    CTC_ERR_OVERFLOW = 3,
    /* a lerp factor left [0,1] (NaN/inf distance): the reference's worker
     * would have panicked at src/math.rs:19 */
    CTC_ERR_LERP_ASSERT = 4,
    CTC_ERR_NO_DEVICE = 5,
};

/* ---- plain-data records ------------------------------------------------ */

/* octree::Span = Range<Point3<f32>> (src/octree/mod.rs:13); Range is not
 * repr(C), so the shim copies start/end into this. */
typedef struct { float start[3]; float end[3]; } ctc_span;

/* mesh::Vertex (src/mesh/mod.rs:255-261): #[repr(C)], 28 bytes, no padding. */
typedef struct { float position[3]; float normal[3]; float distance_from_surface; } ctc_vertex;

enum { CTC_SHAPE_MANDELBULB = 0, CTC_SHAPE_SPHERE = 1 };

/* ctc_shape.flags */
enum {
    /* IEEE mul/add/div/sqrt in the reference's evaluation order, no FMA
     * contraction, glibc's logf algorithm: power-8 distances are bit-identical
     * to the reference's CPU path. */
    CTC_MATH_EXACT = 0,
    /* FMA contraction + MUFU approximations (rsqrt/rcp/lg2).  Distances agree
     * to <= 1e-5 relative away from the escape boundary; see DESIGN.md. */
    CTC_MATH_FAST = 1,
};

/* Mandelbulb<P>{max_iters, bailout} (src/shape/mandelbulb.rs:13-16) or
 * Sphere{center, radius} (src/shape/sphere.rs:7-10). */
typedef struct {
    int32_t  kind;        /* CTC_SHAPE_* */
    uint32_t power;       /* Mandelbulb's const generic P (u8 in the reference) */
    uint64_t max_iters;
    float    bailout;
    float    center[3];   /* Sphere */
    float    radius;      /* Sphere */
    uint32_t flags;       /* CTC_MATH_* */
} ctc_shape;

/* mesh::buffer::Timings (src/mesh/buffer.rs:398-405): the three pass
 * durations (device time, CUDA events) and the totals over the call. */
typedef struct {
    double   first_ms;    /* pass 1: sample grids            (buffer.rs:77-83)   */
    double   second_ms;   /* pass 2: classify + vertices     (buffer.rs:113-275) */
    double   third_ms;    /* pass 3: quads / index emission  (buffer.rs:288-372) */
    uint64_t vertices;
    uint64_t faces;       /* indices / 6, as buffer.rs:380 */
} ctc_timings;

typedef struct ctc_ctx ctc_ctx;

/* ---- context ------------------------------------------------------------ */

CTC_API int ctc_version(void);
/* Number of CUDA devices, or 0. */
CTC_API int ctc_device_count(void);
CTC_API int ctc_ctx_create(int device, ctc_ctx **out);
CTC_API void ctc_ctx_destroy(ctc_ctx *ctx);
/* Adopt an external cudaStream_t (e.g. the host framework's current stream);
 * NULL restores the context's own stream. */
CTC_API int ctc_ctx_set_stream(ctc_ctx *ctx, void *cuda_stream);
/* Spans per internal launch group (0 = automatic: about 512 MiB of sample grids per group, at least four
 * groups per large call so that copies and extraction pipeline behind the next group's DE kernel). */
CTC_API int ctc_ctx_set_group_spans(ctc_ctx *ctx, uint32_t spans_per_group);
/* 1 (default): a launch group's extraction kernels run on a second, high-
 * priority stream concurrently with the next group's DE kernel.  0: strictly
 * serial kernels (per-pass timings then do not overlap; used for profiling). */
CTC_API int ctc_ctx_set_overlap(ctc_ctx *ctx, int enable);
/* Measurement runs: 1 = an event pair around EVERY kernel of the following mesh calls (off by default:
 * the three pass timers are always on).  ctc_mesh_kernel_times then returns, for the last call whose
 * result was fetched, the device milliseconds per kernel summed over the launch groups (index CTC_K_*).
 * Meaningful per kernel only with ctc_ctx_set_overlap(0); with overlap on they share the SMs. */
enum { CTC_K_SAMPLE_GRIDS = 0, CTC_K_FIXUP = 1, CTC_K_CLASSIFY = 2, CTC_K_SCAN = 3, CTC_K_PREFIX = 4 /* emit_lists */,
       CTC_K_VERTEX = 5, CTC_K_QUADS = 6, CTC_NUM_KERNELS = 8 };
CTC_API int ctc_ctx_set_kernel_timing(ctc_ctx *ctx, int enable);
CTC_API int ctc_mesh_kernel_times(ctc_ctx *ctx, double *ms, size_t n);
CTC_API int ctc_ctx_synchronize(ctc_ctx *ctx);
/* Human-readable description of the last failure on this context.  The pointer stays valid until the
 * next failing call on the context: with several threads on ONE context use ctc_last_error_copy, which
 * copies the message out under the context's lock (returns its full length; buf may be NULL). */
CTC_API const char *ctc_last_error(const ctc_ctx *ctx);
CTC_API size_t ctc_last_error_copy(ctc_ctx *ctx, char *buf, size_t len);
/* Kernels launched by this context since creation (all entry points). */
CTC_API uint64_t ctc_kernel_launches(const ctc_ctx *ctx);

/* ---- Shape::batch_min_distance_from (src/shape/mod.rs:89) ---------------- */

/* xyz: n packed Point3<f32> (12-byte stride); out: n f32.  Host pointers. */
CTC_API int ctc_de_batch(ctc_ctx *ctx, const ctc_shape *shape, const float *xyz, size_t n, float *out);
/* Same with device pointers; asynchronous on the context's stream. */
CTC_API int ctc_de_batch_device(ctc_ctx *ctx, const ctc_shape *shape, const float *d_xyz, size_t n, float *d_out);

/* ---- pass 1 only: sample grids (src/mesh/buffer.rs:64-83) ---------------- */

/* grids: nspans x (R+1)^3 f32, index x*(R+1)^2 + y*(R+1) + z
 * (src/util/grid.rs:45-48), sampled over the skirt-expanded span.
 * `spans` is a HOST pointer in both variants (24 bytes per span). */
CTC_API int ctc_sample_grids(ctc_ctx *ctx, const ctc_shape *shape, const ctc_span *spans, size_t nspans,
                     uint32_t resolution, float *grids);
CTC_API int ctc_sample_grids_device(ctc_ctx *ctx, const ctc_shape *shape, const ctc_span *spans, size_t nspans,
                            uint32_t resolution, float *d_grids);

/* ---- MeshBuffer::generate_for_box x nspans (src/mesh/buffer.rs:30-391) ---- */

/* Meshes every span.  Vertices of span s are v[v_off[s] .. v_off[s+1]) and
 * its indices idx[i_off[s] .. i_off[s+1]); indices are span-local (they
 * index into the span's own vertex range), ordered exactly as the reference
 * emits them.  v_off / i_off have nspans+1 entries.  v, idx, v_off, i_off are
 * HOST pointers (pinned memory copies fastest) or any other address the
 * device can copy to (mapped peer-GPU memory, see ctc_ipc_open).  The copy of
 * each launch group's slice overlaps the next group's compute.
 * timings may be NULL. */
CTC_API int ctc_mesh_spans(ctc_ctx *ctx, const ctc_shape *shape, const ctc_span *spans, size_t nspans,
                   uint32_t resolution,
                   ctc_vertex *v, size_t vcap, uint32_t *idx, size_t icap,
                   uint64_t *v_off, uint64_t *i_off, ctc_timings *timings);

/* Concurrent small ctc_mesh_spans calls (<= 64 spans, host destinations) of one context are coalesced:
 * the caller that finds the context idle launches; calls that arrive meanwhile queue up and are served
 * by ONE batched launch per shape/resolution (results and per-call error statuses are exactly those of
 * separate calls; pass timings of a batch are apportioned by span count).  0 disables it.
 * ctc_ctx_coalescing_stats: batched launches so far and the requests they served. */
CTC_API int ctc_ctx_set_coalescing(ctc_ctx *ctx, int enable);
CTC_API int ctc_ctx_coalescing_stats(ctc_ctx *ctx, uint64_t *batches, uint64_t *requests);

/* Device-resident variant: d_v / d_idx / d_v_off / d_i_off are device
 * pointers, `spans` stays a host pointer.  Asynchronous on the context's
 * stream; the status only covers argument checks and launch errors.  Call
 * ctc_mesh_result() afterwards to synchronise and fetch totals, overflow and
 * lerp-assert status. */
CTC_API int ctc_mesh_spans_device(ctc_ctx *ctx, const ctc_shape *shape, const ctc_span *spans, size_t nspans,
                          uint32_t resolution,
                          ctc_vertex *d_v, size_t vcap, uint32_t *d_idx, size_t icap,
                          uint64_t *d_v_off, uint64_t *d_i_off);
/* Synchronises the stream and reports the last ctc_mesh_spans_device call:
 * total vertices / indices REQUIRED (even on overflow), pass timings.
 * Returns CTC_OK, CTC_ERR_OVERFLOW or CTC_ERR_LERP_ASSERT. */
CTC_API int ctc_mesh_result(ctc_ctx *ctx, uint64_t *n_vertices, uint64_t *n_indices, ctc_timings *timings);

/* ---- focus rays: ShapeMesh::get_focii's sphere tracing (src/mesh/mod.rs:229-241) ---- */

/* N4: the ray-marcher the reference plans (README.md:12-16, GLSL distance estimator
 * src/shape/mandelbulb.frag) as one kernel: pixel (i, j), i < width, j < height, looks from `eye` through
 * top_left + (i + 0.5) du + (j + 0.5) dv and is sphere-traced exactly like a focus ray of ShapeMesh::get_focii
 * (src/mesh/mod.rs:229-241: pos += dir * d until d < epsilon, at most max_steps steps).  out: width * height
 * records of four floats, row-major: the final position and the distance travelled, negative when the ray did
 * not hit.  The ray set-up is plain IEEE f32 arithmetic in a fixed order (normalise = v * (1 / sqrt((x*x+y*y)+z*z)),
 * as cgmath), so a host can form the same rays bit for bit; with an exact-mode shape a pixel equals ctc_ray_march
 * on its ray.  ctc_render copies the image to host memory and returns synchronised; ctc_render_device leaves it in
 * device memory (16-byte aligned), asynchronous on the context's stream. */
typedef struct ctc_camera_rays {
    float eye[3];
    float top_left[3];
    float du[3];   /* one pixel to the right */
    float dv[3];   /* one pixel down */
} ctc_camera_rays;
CTC_API int ctc_render(ctc_ctx *ctx, const ctc_shape *shape, const ctc_camera_rays *cam, uint32_t width, uint32_t height,
                       uint32_t max_steps, float epsilon, float *out);
CTC_API int ctc_render_device(ctc_ctx *ctx, const ctc_shape *shape, const ctc_camera_rays *cam, uint32_t width,
                              uint32_t height, uint32_t max_steps, float epsilon, float *d_out);

/* DE-bound span culling (not in the reference).  keep[i] = 0 when span i needs no meshing: the
 * distance estimate at its centre (exact arithmetic) exceeds `safety` times the half-diagonal of the skirt-
 * expanded span, i.e. the surface cannot reach it (Shape::min_distance_from is "a lower bound of the distance",
 * shape/mod.rs:26-37); for shapes with an upper bound too (Sphere, max_distance_from = min_distance_from,
 * sphere.rs:37-39) also when the span lies entirely inside.  The Mandelbulb estimate is not a rigorous bound:
 * safety >= 1 is the caller's margin (the host mirror uses 2); tests check every culled span of the BASELINE
 * configs against the CPU oracle.  Culling is never applied implicitly by the mesh calls. */
CTC_API int ctc_cull_spans(ctc_ctx *ctx, const ctc_shape *shape, const ctc_span *spans, size_t nspans,
                           uint32_t resolution, float safety, uint8_t *keep);
/* Cost-aware span order: order[k] = index of the span to mesh k-th, ascending |DE(centre)| / reach --
 * the spans most likely to hold surface first, the provably empty ones last (stable; NaN counts as 0).  A mesh
 * call made in this order produces its bytes early, so the copy pipeline behind the launch groups (PCIe to the
 * host, or the NVLink puts of the multi-GPU gather) is busy from the first group on and the groups computed last
 * leave nothing to copy after the kernels end.  The meshes do not depend on the order.  Host pointers; synchronous
 * (one DE evaluation per span). */
CTC_API int ctc_order_spans(ctc_ctx *ctx, const ctc_shape *shape, const ctc_span *spans, size_t nspans,
                            uint32_t resolution, uint32_t *order);

/* n rays (origin, unit direction; packed xyz, HOST pointers).  Each ray repeats
 * `d = DE(pos); pos += dir * d; if d < epsilon { hit }` up to max_steps times
 * (the reference uses EPSILON = 1e-6, MAX_ITERS = 100).  out_pos: n packed xyz,
 * out_hit: n u32 (1 = Some(pos), 0 = None). */
CTC_API int ctc_ray_march(ctc_ctx *ctx, const ctc_shape *shape, const float *origin, const float *dir, size_t n,
                          uint32_t max_steps, float epsilon, float *out_pos, uint32_t *out_hit);

/* ---- span scheduler over the GPUs of one box (replaces ThreadPool::new(num_cpus) +
 * mpsc channel, src/mesh/mod.rs:60-62, 129-161, for a single-process host) ------------------------------
 *
 * A ctc_multi owns one context and one worker thread per device.  A call deals the spans round-robin
 * (span s -> device s % ngpus; neighbouring spans cost alike, so the deal balances the devices), every
 * device meshes its share, and each device's launch groups are copied into ITS region of the caller's
 * buffers while it computes the following groups: over NVLink into device memory of devices[0]
 * (..._multi_device: "gather to rank 0"), or over each device's own PCIe link into host memory
 * (..._multi).  Regions have fixed capacities (ctc_multi_shard_plan: proportional to span counts), so
 * no count exchange and no collective is needed.
 *
 * Result layout: the vertices of span s are v[span_v[2s] .. span_v[2s+1]) and its indices
 * idx[span_i[2s] .. span_i[2s+1]) (span-local ids, reference order); span_v / span_i are HOST arrays of
 * 2 * nspans entries.  Completion order is free in the reference (results are keyed by span.center(),
 * src/mesh/mod.rs:110-120, 147), so the buffers are device-major, not span-major.
 * need (may be NULL): need[0] / need[1] = a vcap / icap with which the call cannot overflow (valid
 * after CTC_OK or CTC_ERR_OVERFLOW: re-allocate and retry).  timings: per-pass maxima over the devices
 * (they run concurrently), vertex / face totals. */
typedef struct ctc_multi ctc_multi;
/* devices: ngpus CUDA device ordinals, or NULL for 0 .. ngpus-1; ngpus <= 0: every device of the box. */
CTC_API int ctc_multi_create(const int *devices, int ngpus, ctc_multi **out);
CTC_API void ctc_multi_destroy(ctc_multi *m);
CTC_API int ctc_multi_ngpus(const ctc_multi *m);
/* The context of device i (ctc_ctx_set_group_spans, ctc_ctx_set_fast_band, ... apply per device). */
CTC_API ctc_ctx *ctc_multi_ctx(ctc_multi *m, int i);
CTC_API size_t ctc_multi_last_error(ctc_multi *m, char *buf, size_t len);
CTC_API int ctc_mesh_spans_multi(ctc_multi *m, const ctc_shape *shape, const ctc_span *spans, size_t nspans,
                                 uint32_t resolution, ctc_vertex *v, size_t vcap, uint32_t *idx, size_t icap,
                                 uint64_t *span_v, uint64_t *span_i, uint64_t need[2], ctc_timings *timings);
/* d_v / d_idx: device memory of devices[0]. */
CTC_API int ctc_mesh_spans_multi_device(ctc_multi *m, const ctc_shape *shape, const ctc_span *spans, size_t nspans,
                                        uint32_t resolution, ctc_vertex *d_v, size_t vcap, uint32_t *d_idx, size_t icap,
                                        uint64_t *span_v, uint64_t *span_i, uint64_t need[2], ctc_timings *timings);
/* The sharding and region tables of a call, without touching a GPU (also what the entry points use):
 * first_v / first_i [ngpus + 1] = first vertex / index of every device's region (64-element aligned),
 * owner [nspans] (may be NULL) = device position of every span. */
CTC_API int ctc_multi_shard_plan(size_t nspans, int ngpus, size_t vcap, size_t icap, uint64_t *first_v,
                                 uint64_t *first_i, uint32_t *owner);

/* ---- peer memory for the multi-GPU gather -------------------------------- */

/* Plain cudaMalloc/cudaFree on the context's device (IPC handles need the base
 * pointer of an allocation, which pooled allocators do not give). */
CTC_API int ctc_device_alloc(ctc_ctx *ctx, size_t bytes, void **d_ptr);
CTC_API int ctc_device_free(ctc_ctx *ctx, void *d_ptr);
/* cudaIpcGetMemHandle / cudaIpcOpenMemHandle / cudaIpcCloseMemHandle: lets
 * another process (one per GPU) map a buffer of this device.  `handle` is 64
 * bytes.  The destination pointers of ctc_mesh_spans may be such mapped peer
 * memory (or pinned host memory): each launch group's slice of the mesh is
 * copied there as soon as the group finishes, overlapping the copy over NVLink
 * with the next group's compute. */
CTC_API int ctc_ipc_export(ctc_ctx *ctx, const void *d_ptr, unsigned char handle[64]);
CTC_API int ctc_ipc_open(ctc_ctx *ctx, const unsigned char handle[64], void **d_ptr);
CTC_API int ctc_ipc_close(ctc_ctx *ctx, void *d_ptr);

/* ---- interop buffers: the upload step after the path ------- */

/* Replaces the host round trip of `MeshView::new` (src/mesh/view.rs:23-41: `create_buffer_init` from
 * `bytemuck::cast_slice(vertices)` / `(indices)`, i.e. host memory -> Vulkan buffer).  An interop buffer is
 * device memory of the context's GPU with a POSIX file-descriptor handle (CUDA virtual memory management):
 * the renderer imports `fd` as VkDeviceMemory (VkImportMemoryFdInfoKHR, OPAQUE_FD, VK_KHR_external_memory_fd;
 * allocationSize = *allocated_bytes) and binds its vertex / index VkBuffers to it; ctc_mesh_spans_device,
 * ctc_mesh_spans_multi_device and ctc_render_device take `*d_ptr` like any device pointer, so the meshes are
 * written straight into the memory the draw calls read -- no PCIe bytes but the per-span offset tables.
 * Hand-over: ctc_ctx_synchronize (or the stream's own event) before the renderer's submit.
 * The caller owns `fd` (close(2) it after the import; the memory lives while any mapping or import does).
 * ctc_interop_import maps such a descriptor in another context or process (CUDA consumer, tests).
 * ctc_interop_free unmaps a buffer obtained from either call; ctc_ctx_destroy frees what is left. */
CTC_API int ctc_interop_alloc(ctc_ctx *ctx, size_t bytes, void **d_ptr, int *fd, size_t *allocated_bytes);
CTC_API int ctc_interop_import(ctc_ctx *ctx, int fd, size_t allocated_bytes, void **d_ptr);
CTC_API int ctc_interop_free(ctc_ctx *ctx, void *d_ptr);
/* Synchronous read of device memory (interop or not) into host memory, ordered behind the context's stream:
 * the offset tables of a ctc_mesh_spans_device call, or a look at an imported buffer. */
CTC_API int ctc_device_read(ctc_ctx *ctx, void *host_dst, const void *d_src, size_t bytes);
/* The other direction (the destination may be a mapped peer buffer, ctc_ipc_open): small host-side tables that
 * travel with a gathered result, e.g. the order a rank meshed its spans in (ctc_order_spans). */
CTC_API int ctc_device_write(ctc_ctx *ctx, void *d_dst, const void *host_src, size_t bytes);

/* Index wire format of ctc_mesh_spans (the host / peer destination variant only).  0 (default): six
 * u32 indices per quad, the reference's layout.  1: one packed 8-byte record per quad (four 16-bit
 * span-local vertex ids; v0 < v1 always holds, so swapping the first two encodes the winding) written
 * densely at QUAD offsets into `idx` -- a third of the bytes for the multi-GPU gather.  i_off keeps
 * counting indices (6 per quad).  A span with >= 65536 vertices makes the call fail with CTC_ERR_OVERFLOW.
 * ctc_expand_quads widens nquads records into 6 * nquads u32 indices (device pointers, asynchronous
 * on the context's stream). */
CTC_API int ctc_ctx_set_index_wire(ctc_ctx *ctx, int packed_quads);
CTC_API int ctc_expand_quads(ctc_ctx *ctx, const void *d_records, size_t nquads, uint32_t *d_idx);
/* The same widening on the calling HOST thread (host pointers; no context, no GPU): what the host-side pool of
 * ctc_mesh_spans runs per piece.  A consumer that asked for packed records (ctc_ctx_set_index_wire) can widen
 * them itself with this. */
CTC_API int ctc_expand_quads_host(const void *records, size_t nquads, uint32_t *idx);
/* Packed index wire towards a PEER GPU (ctc_ctx_set_index_wire(ctx, 1), device destination): after every launch
 * group's records ctc_mesh_spans also puts one 64-bit progress word {done:1 | call epoch:23 | quads so far:40}
 * at d_progress_word (memory of the destination GPU), in order behind the records, and a last one with the
 * done bit when the call's records are all there.  The destination can widen the slices that have landed
 * (ctc_expand_quads) while the sender is still computing instead of after the step.  The epoch counts the
 * sender's calls since the word was set (first call = 1).  NULL switches it off. */
CTC_API int ctc_ctx_set_wire_progress(ctc_ctx *ctx, void *d_progress_word);
/* HOST destinations of ctc_mesh_spans (what the reference's caller passes, mesh/mod.rs:141-148): by default
 * (1) the indices cross PCIe as the packed 8-byte records above and a pool of host threads widens every launch
 * group's slice into the caller's u32 index buffer while later groups are still computing -- the caller sees
 * exactly the reference's six u32 per quad, the device->host copy carries a third of the index bytes.  A span
 * with >= 65536 vertices makes the call repeat itself with u32 indices on the wire (counted in `fallbacks`).
 * Calls of fewer than 128 spans are not PCIe-bound and keep six u32 per quad (2: packed records for every call).
 * 0: always copy six u32 per quad.  Independently of this switch, buffers in PAGEABLE host memory are filled
 * through pinned landing buffers by the same host threads (calls of more than 8 spans).  The pool has all but two
 * hardware threads (shared out over the devices of a ctc_multi); CANTUCCI_B200_EXPAND_THREADS overrides the count. */
CTC_API int ctc_ctx_set_host_index_wire(ctc_ctx *ctx, int packed_quads);
/* The packed wire loads the HOST's memory system (the widening threads read 8 and write 24 bytes per quad beside the
 * DMA writes), the u32 wire loads the PCIe link.  Towards page-locked index buffers the call can mix them by
 * launch group: of every `den` groups `num` travel packed, the others as six u32 per quad straight into the caller's
 * buffer (num = den, the default: all packed -- the fastest split on the measured hosts, 10.5 / 11.2 / 11.8 / 13.7 /
 * 13.0 ms per benched volume for 4/4, 3/4, 2/4, 1/4, 0/4; num = 0: none).  CANTUCCI_B200_HOST_WIRE_SHARE="num/den" sets the
 * default of new contexts.  ctc_mesh_d2h_bytes: the mesh bytes the last host-pointer call copied device -> host. */
CTC_API int ctc_ctx_set_host_wire_share(ctc_ctx *ctx, uint32_t num, uint32_t den);
CTC_API int ctc_mesh_d2h_bytes(ctc_ctx *ctx, uint64_t *bytes);
CTC_API int ctc_ctx_host_index_wire_stats(ctc_ctx *ctx, uint64_t *calls, uint64_t *fallbacks, uint32_t *threads);

/* Page-lock / unlock host memory the caller owns (cudaHostRegister, portable), e.g. a POSIX shared-
 * memory segment that several one-GPU processes fill with ctc_mesh_spans, each over its own PCIe link. */
CTC_API int ctc_host_register(ctc_ctx *ctx, void *ptr, size_t bytes);
CTC_API int ctc_host_unregister(ctc_ctx *ctx, void *ptr);

/* ---- measurement aids (not part of the reference's interface) ------------- */

/* Iteration statistics of the sample lattices (exact arithmetic == the
 * reference's counts): out[0] = sum of completed iterations, out[1] = samples
 * that left through `r > bailout`, out[2] = samples.  Used for the algorithmic
 * flop count flops(sample) = 75 k + 6 [bailed] + 10. Host pointers; synchronous. */
CTC_API int ctc_iteration_stats(ctc_ctx *ctx, const ctc_shape *shape, const ctc_span *spans, size_t nspans,
                                uint32_t resolution, uint64_t out[3]);
/* ---- fast mode's sign repair (CTC_MATH_FAST) ----------------------------------
 * Fast mode is SIGN-EXACT: a sample whose sign the fast arithmetic cannot guarantee (its last radius
 * is within kappa * max dr * polar stretch of 1, an iterate touched the z axis, or anything went NaN)
 * is re-evaluated with the exact arithmetic, so that the inside/outside classification -- and with it
 * every index buffer -- equals exact mode's (for power 8: the reference's).
 * ctc_ctx_set_fast_band overrides the calibrated band (0 = default); ctc_mesh_fixups reports, for the
 * last mesh call whose result was fetched, how many samples were re-evaluated and how many of those
 * changed sign. */
CTC_API int ctc_ctx_set_fast_band(ctc_ctx *ctx, float kappa);
CTC_API int ctc_mesh_fixups(ctc_ctx *ctx, uint64_t *suspects, uint64_t *sign_fixups);
/* Parity aid: the sign bit-planes of the spans' sample grids as pass 1 produces them (fast mode: after
 * the sign repair).  planes: nspans x ((R+1)^3 + 31) / 32 u32, HOST pointer; bit j of a span's plane is
 * !f32::is_sign_positive(grid[j]), j = x*(R+1)^2 + y*(R+1) + z -- everything the mesher's topology
 * depends on (src/mesh/buffer.rs:130-147, 299-350).  Synchronous. */
CTC_API int ctc_sample_signs(ctc_ctx *ctx, const ctc_shape *shape, const ctc_span *spans, size_t nspans,
                             uint32_t resolution, uint32_t *planes);
/* Calibration aid: evaluates every sample of the spans' lattices with the raw fast path AND the exact
 * path and counts sign mismatches, suspects and uncovered mismatches for kappa = 2^-8 .. 2^-31
 * (layout: csrc/kernels.cuh, fast_sign_probe_kernel; out_words >= 489).  Power 8 only. */
CTC_API int ctc_fast_sign_probe(ctc_ctx *ctx, const ctc_shape *shape, const ctc_span *spans, size_t nspans,
                                uint32_t resolution, uint64_t *out, size_t out_words);

/* The same statistics for n packed xyz points in DEVICE memory (the 7 evaluation points per vertex of
 * pass 2, for its algorithmic flop count). */
CTC_API int ctc_iteration_stats_points(ctc_ctx *ctx, const ctc_shape *shape, const float *d_xyz, size_t n,
                                       uint64_t out[3]);
/* Sustained FP32 FMA rate of the device in TFLOP/s (FMA = 2 flops), measured
 * with dependent-free FFMA streams on every SM. */
CTC_API int ctc_fp32_peak_probe(ctc_ctx *ctx, double *tflops, int *num_sms);

#ifdef __cplusplus
}
#endif
#endif /* CANTUCCI_B200_H */
