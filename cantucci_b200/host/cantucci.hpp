// cantucci.hpp -- C++ host-side mirror of the reference's Rust interface for the hot path,
// over the C ABI (include/cantucci_b200.h).  Header-only; link with libcantucci_b200.so.
//
// The reference's own toolchain (rustc) is not available in this image, and the reference is
// compiled code, so this is the compiled-language mirror of:
//   trait Shape            /root/reference/src/shape/mod.rs:24-105
//   Mandelbulb<P>          src/shape/mandelbulb.rs:13-33
//   Sphere                 src/shape/sphere.rs:7-41
//   octree::Span           src/octree/mod.rs:13-32
//   mesh::Vertex           src/mesh/mod.rs:255-261
//   MeshBuffer / Timings   src/mesh/buffer.rs:24-42, 398-405
//   MeshView               src/mesh/view.rs:14-41   (the upload step, without the host round trip: MeshViews)
// Same names, argument meaning and error behaviour: where the reference panics (assert!), these
// throw cantucci::Panic.  The Rust binding itself is shown in INTEGRATION.md.
#pragma once

#include <array>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/cantucci_b200.h"

namespace cantucci {

struct Panic : std::runtime_error { using std::runtime_error::runtime_error; };

using Point3 = std::array<float, 3>;           // cgmath::Point3<f32>, 12 bytes
static_assert(sizeof(Point3) == 12, "Point3<f32> must be 3 packed f32");

struct Span { Point3 start, end; };            // Range<Point3<f32>>

inline Point3 center(const Span& s) {          // SpanExt::center, octree/mod.rs:21-23
    return {s.start[0] + (s.end[0] - s.start[0]) / 2.0f, s.start[1] + (s.end[1] - s.start[1]) / 2.0f,
            s.start[2] + (s.end[2] - s.start[2]) / 2.0f};
}

using Vertex = ctc_vertex;                     // #[repr(C)] Vertex, 28 bytes
static_assert(sizeof(Vertex) == 28, "mesh::Vertex is 28 bytes without padding");

// One CUDA device + stream + workspace.  Shape: Sync + Send -> callable from any thread.
class Context {
public:
    explicit Context(int device = 0) {
        const int rc = ctc_ctx_create(device, &ctx_);
        if (rc != CTC_OK) throw std::runtime_error("ctc_ctx_create failed (no CUDA device; there is no CPU fallback)");
    }
    ~Context() { ctc_ctx_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    ctc_ctx* get() const { return ctx_; }
    void check(int rc) const {
        if (rc == CTC_OK) return;
        const std::string msg = ctc_last_error(ctx_);
        if (rc == CTC_ERR_INVALID_ARGUMENT || rc == CTC_ERR_LERP_ASSERT) throw Panic(msg);   // the reference's assert!s
        throw std::runtime_error(msg);
    }
private:
    ctc_ctx* ctx_ = nullptr;
};

// trait Shape (shape/mod.rs:24)
class Shape {
public:
    virtual ~Shape() = default;
    virtual ctc_shape descriptor() const = 0;
    virtual Span bounding_box() const = 0;

    // shape/mod.rs:89 (impl_batch_methods!, shape/util.rs:3-5)
    std::vector<float> batch_min_distance_from(const Context& ctx, const std::vector<Point3>& points) const {
        std::vector<float> out(points.size());
        const ctc_shape d = descriptor();
        ctx.check(ctc_de_batch(ctx.get(), &d, points.empty() ? nullptr : points[0].data(), points.size(), out.data()));
        return out;
    }
    // shape/mod.rs:37
    float min_distance_from(const Context& ctx, Point3 p) const { return batch_min_distance_from(ctx, {p})[0]; }
    // shape/mod.rs:78-80
    bool contains(const Context& ctx, Point3 p) const { return min_distance_from(ctx, p) < 0.0f; }
};

// Mandelbulb<const P: u8> (mandelbulb.rs:13-27)
class Mandelbulb final : public Shape {
public:
    Mandelbulb(uint8_t power, uint64_t max_iters, float bailout, bool fast = false)
        : power_(power), max_iters_(max_iters), bailout_(bailout), fast_(fast) {
        if (!(max_iters >= 1)) throw Panic("assertion failed: max_iters >= 1");        // mandelbulb.rs:20
    }
    static Mandelbulb classic(uint64_t max_iters, float bailout, bool fast = false) {   // mandelbulb.rs:29-33
        return Mandelbulb(8, max_iters, bailout, fast);
    }
    Span bounding_box() const override { return {{-1.2f, -1.2f, -1.2f}, {1.2f, 1.2f, 1.2f}}; }   // mandelbulb.rs:53-57
    ctc_shape descriptor() const override {
        ctc_shape s{};
        s.kind = CTC_SHAPE_MANDELBULB; s.power = power_; s.max_iters = max_iters_; s.bailout = bailout_;
        s.flags = fast_ ? CTC_MATH_FAST : CTC_MATH_EXACT;
        return s;
    }
private:
    uint8_t power_; uint64_t max_iters_; float bailout_; bool fast_;
};

// Sphere (sphere.rs:7-41)
class Sphere final : public Shape {
public:
    Sphere(Point3 center, float radius) : center_(center), radius_(radius) {}
    Span bounding_box() const override {                                                // sphere.rs:28-31
        return {{center_[0] + -radius_, center_[1] + -radius_, center_[2] + -radius_},
                {center_[0] + radius_, center_[1] + radius_, center_[2] + radius_}};
    }
    ctc_shape descriptor() const override {
        ctc_shape s{};
        s.kind = CTC_SHAPE_SPHERE; s.center[0] = center_[0]; s.center[1] = center_[1]; s.center[2] = center_[2];
        s.radius = radius_;
        return s;
    }
private:
    Point3 center_; float radius_;
};

// Timings (buffer.rs:398-405); durations are device milliseconds.
struct Timings { double first = 0, second = 0, third = 0; uint32_t vertices = 0, faces = 0; };

// MeshBuffer (buffer.rs:24-27)
struct MeshBuffer {
    std::vector<Vertex> vertices;
    std::vector<uint32_t> indices;

    // MeshBuffer::generate_for_box (buffer.rs:30-42): panics (throws Panic) on span.start >= span.end,
    // resolution not a non-zero power of two, and on the lerp assert (math.rs:19).
    static std::pair<MeshBuffer, Timings> generate_for_box(const Context& ctx, const Span& span, const Shape& shape,
                                                           uint32_t resolution) {
        auto r = generate_for_boxes(ctx, {span}, shape, resolution);
        return {std::move(r.first[0]), r.second};
    }

    // The batching point of mesh/mod.rs:129-161: all empty leaves in ONE call.
    static std::pair<std::vector<MeshBuffer>, Timings> generate_for_boxes(const Context& ctx, const std::vector<Span>& spans,
                                                                         const Shape& shape, uint32_t resolution) {
        static_assert(sizeof(Span) == sizeof(ctc_span), "Span must flatten to 6 f32");
        const size_t n = spans.size();
        const ctc_shape d = shape.descriptor();
        size_t vcap = std::max<size_t>(1024, n * 8 * size_t(resolution) * resolution), icap = 6 * vcap;
        std::vector<Vertex> v; std::vector<uint32_t> idx;
        std::vector<uint64_t> v_off(n + 1), i_off(n + 1);
        ctc_timings t{};
        for (int attempt = 0;; ++attempt) {
            v.resize(vcap); idx.resize(icap);
            const int rc = ctc_mesh_spans(ctx.get(), &d, reinterpret_cast<const ctc_span*>(spans.data()), n, resolution,
                                          v.data(), vcap, idx.data(), icap, v_off.data(), i_off.data(), &t);
            if (rc == CTC_ERR_OVERFLOW && attempt == 0) { vcap = v_off[n]; icap = i_off[n]; continue; }
            ctx.check(rc);
            break;
        }
        std::vector<MeshBuffer> out(n);
        for (size_t k = 0; k < n; ++k) {
            out[k].vertices.assign(v.begin() + v_off[k], v.begin() + v_off[k + 1]);
            out[k].indices.assign(idx.begin() + i_off[k], idx.begin() + i_off[k + 1]);
        }
        return {std::move(out), Timings{t.first_ms, t.second_ms, t.third_ms, uint32_t(t.vertices), uint32_t(t.faces)}};
    }
};

// Cost-aware span order (ctc_order_spans; not in the reference, whose pool takes the jobs in tree order,
// mesh/mod.rs:129-161): indices of `spans`, the ones most likely to hold surface first.
inline std::vector<uint32_t> order_spans(const Context& ctx, const std::vector<Span>& spans, const Shape& shape, uint32_t resolution) {
    std::vector<uint32_t> order(spans.size());
    const ctc_shape d = shape.descriptor();
    ctx.check(ctc_order_spans(ctx.get(), &d, reinterpret_cast<const ctc_span*>(spans.data()), spans.size(), resolution, order.data()));
    return order;
}

// Device memory the renderer imports by file descriptor (ctc_interop_alloc): VkImportMemoryFdInfoKHR on the Vulkan
// side, ctc_interop_import in another CUDA context or process.
class InteropBuffer {
public:
    InteropBuffer(const Context& ctx, size_t bytes) : ctx_(&ctx) { ctx.check(ctc_interop_alloc(ctx.get(), bytes, &ptr_, &fd_, &bytes_)); }
    ~InteropBuffer() { if (ptr_) ctc_interop_free(ctx_->get(), ptr_); }      // (the descriptor stays the caller's to close)
    InteropBuffer(const InteropBuffer&) = delete;
    InteropBuffer& operator=(const InteropBuffer&) = delete;
    void* ptr() const { return ptr_; }
    int fd() const { return fd_; }
    size_t bytes() const { return bytes_; }        // the allocated size: VkMemoryAllocateInfo::allocationSize
private:
    const Context* ctx_; void* ptr_ = nullptr; int fd_ = -1; size_t bytes_ = 0;
};

// MeshView{vbuf, ibuf, num_indices} (view.rs:14-18) as byte ranges of the batch's two interop buffers: the vertex /
// index buffer bindings of the span's draw call (view.rs:76-79; indices are span-local, base vertex 0).
struct MeshView { uint64_t vertex_offset, index_offset; uint32_t num_vertices, num_indices; };

struct MeshViews {
    std::vector<uint64_t> v_off, i_off;            // per-span tables, in vertices / indices
    MeshView view(size_t k) const {
        return {v_off[k] * sizeof(Vertex), i_off[k] * 4, uint32_t(v_off[k + 1] - v_off[k]), uint32_t(i_off[k + 1] - i_off[k])};
    }
    size_t size() const { return v_off.empty() ? 0 : v_off.size() - 1; }

    // generate_for_box + MeshView::new for every span without leaving the GPU (mesh/mod.rs:141-146): the meshes are
    // written straight into `vbuf` / `ibuf`, only the two offset tables come back.  Throws std::length_error when a
    // buffer is too small (what(): the required vertices and indices).
    static std::pair<MeshViews, Timings> generate(const Context& ctx, const std::vector<Span>& spans, const Shape& shape,
                                                  uint32_t resolution, const InteropBuffer& vbuf, const InteropBuffer& ibuf) {
        const size_t n = spans.size();
        const ctc_shape d = shape.descriptor();
        void* tables = nullptr;
        ctx.check(ctc_device_alloc(ctx.get(), 2 * (n + 1) * 8, &tables));
        struct Free { const Context& c; void* p; ~Free() { ctc_device_free(c.get(), p); } } guard{ctx, tables};
        uint64_t* d_v_off = static_cast<uint64_t*>(tables);
        ctx.check(ctc_mesh_spans_device(ctx.get(), &d, reinterpret_cast<const ctc_span*>(spans.data()), n, resolution,
                                        static_cast<ctc_vertex*>(vbuf.ptr()), vbuf.bytes() / sizeof(Vertex),
                                        static_cast<uint32_t*>(ibuf.ptr()), ibuf.bytes() / 4, d_v_off, d_v_off + n + 1));
        uint64_t nv = 0, ni = 0;
        ctc_timings t{};
        const int rc = ctc_mesh_result(ctx.get(), &nv, &ni, &t);
        if (rc == CTC_ERR_OVERFLOW) throw std::length_error("interop buffers too small: need " + std::to_string(nv) + " vertices, " + std::to_string(ni) + " indices");
        ctx.check(rc);
        MeshViews out;
        std::vector<uint64_t> both(2 * (n + 1));
        ctx.check(ctc_device_read(ctx.get(), both.data(), tables, both.size() * 8));
        out.v_off.assign(both.begin(), both.begin() + n + 1);
        out.i_off.assign(both.begin() + n + 1, both.end());
        return {std::move(out), Timings{t.first_ms, t.second_ms, t.third_ms, uint32_t(t.vertices), uint32_t(t.faces)}};
    }
};

}  // namespace cantucci
