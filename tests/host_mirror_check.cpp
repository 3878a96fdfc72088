// Exercises the C++ host mirror (cantucci_b200/host/cantucci.hpp).  Usage: host_mirror_check [gpu]
//  without "gpu": checks the no-device error path and the panic mirroring of constructor asserts;
//  with    "gpu": meshes one span and prints counts + an FNV hash the pytest compares with the oracle.
#include <cstdio>
#include <cstring>
#include "../cantucci_b200/host/cantucci.hpp"

static uint64_t fnv(const void* p, size_t n, uint64_t h = 1469598103934665603ull) {
    const unsigned char* b = static_cast<const unsigned char*>(p);
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

int main(int argc, char** argv) {
    using namespace cantucci;
    bool threw = false;
    try { Mandelbulb::classic(0, 2.5f); } catch (const Panic&) { threw = true; }
    if (!threw) { std::puts("FAIL: max_iters >= 1 assert not mirrored"); return 1; }
    if (argc < 2 || std::strcmp(argv[1], "gpu") != 0) {
        if (ctc_device_count() == 0) {
            try { Context c(0); std::puts("FAIL: context without a device"); return 1; }
            catch (const std::runtime_error&) { std::puts("ok: no device -> error, no CPU fallback"); }
        } else std::puts("ok: device present");
        return 0;
    }
    Context ctx(0);
    const Mandelbulb bulb = Mandelbulb::classic(6, 2.5f);
    auto [mesh, t] = MeshBuffer::generate_for_box(ctx, Span{{0.f, 0.f, 0.f}, {0.6f, 0.6f, 0.6f}}, bulb, 16);
    std::printf("vertices %zu indices %zu vhash %016llx ihash %016llx\n", mesh.vertices.size(), mesh.indices.size(),
                (unsigned long long)fnv(mesh.vertices.data(), mesh.vertices.size() * sizeof(Vertex)),
                (unsigned long long)fnv(mesh.indices.data(), mesh.indices.size() * 4));
    threw = false;
    try { MeshBuffer::generate_for_box(ctx, Span{{0.f, 0.f, 0.f}, {1.f, 1.f, 1.f}}, bulb, 12); } catch (const Panic&) { threw = true; }
    if (!threw) { std::puts("FAIL: power-of-two assert not mirrored"); return 1; }
    const float d = bulb.min_distance_from(ctx, {0.3f, 0.2f, 0.1f});
    uint32_t bits; std::memcpy(&bits, &d, 4);
    std::printf("de %08x\n", bits);
    // N2: the same span (and an empty one) meshed straight into interop buffers; read back through the C ABI
    {
        const std::vector<Span> spans{Span{{0.9f, 0.9f, 0.9f}, {1.2f, 1.2f, 1.2f}}, Span{{0.f, 0.f, 0.f}, {0.6f, 0.6f, 0.6f}}};
        InteropBuffer vbuf(ctx, 1 << 20), ibuf(ctx, 1 << 20);
        auto [views, tv] = MeshViews::generate(ctx, spans, bulb, 16, vbuf, ibuf);
        const MeshView w = views.view(1), e = views.view(0);
        std::vector<Vertex> v(w.num_vertices); std::vector<uint32_t> i(w.num_indices);
        ctx.check(ctc_device_read(ctx.get(), v.data(), static_cast<char*>(vbuf.ptr()) + w.vertex_offset, v.size() * sizeof(Vertex)));
        ctx.check(ctc_device_read(ctx.get(), i.data(), static_cast<char*>(ibuf.ptr()) + w.index_offset, i.size() * 4));
        const bool same = v.size() == mesh.vertices.size() && i == mesh.indices &&
                          std::memcmp(v.data(), mesh.vertices.data(), v.size() * sizeof(Vertex)) == 0;
        const std::vector<uint32_t> order = order_spans(ctx, spans, bulb, 16);
        std::printf("interop fd %d empty %u same %d order %u %u\n", vbuf.fd() >= 0 ? 1 : 0, e.num_indices, same ? 1 : 0, order[0], order[1]);
        // 64 copies of the surface span at R = 32 do not fit one allocation granule: the required sizes are reported
        int small = 0;
        InteropBuffer tiny(ctx, 28);
        try { MeshViews::generate(ctx, std::vector<Span>(64, spans[1]), bulb, 32, tiny, ibuf); }
        catch (const std::length_error&) { small = 1; }
        catch (const std::exception&) { small = -1; }
        std::printf("overflow %d\n", small);
    }
    return 0;
}
