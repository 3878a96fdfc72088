"""Host-side mirror of the upload step, `MeshView::new` (/root/reference/src/mesh/view.rs:14-41), without the
host round trip.

The reference builds one `MeshView {vbuf, ibuf, num_indices}` per span from host slices
(`create_buffer_init(bytemuck::cast_slice(vertices))`).  Here a whole batch of spans is meshed straight into two
INTEROP buffers -- device memory with a POSIX file-descriptor handle (ctc_interop_alloc) that the renderer imports
as VkDeviceMemory -- and a `MeshView` is a pair of byte ranges in them: the vertex / index buffer bindings of the
span's draw call (`set_vertex_buffer(0, vbuf.slice(..))`, `set_index_buffer(ibuf.slice(..))`,
`draw_indexed(0..num_indices, 0, 0..1)`, view.rs:76-79).  Indices are span-local, so base vertex 0 with the
vertex binding at the span's first vertex is exactly the reference's per-span buffer pair.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import _lib
from .mesh import Timings, VERTEX_DTYPE, _check_args
from .octree import spans_array
from .shape import Shape


class InteropBuffer:
    """Device memory of `ctx`'s GPU that another API or process can import through `fd` (ctc_interop_alloc)."""

    def __init__(self, ctx: _lib.Context, nbytes: int):
        self.ctx = ctx
        p, fd, size = C.c_void_p(), C.c_int(-1), C.c_size_t(0)
        ctx.check(_lib.lib().ctc_interop_alloc(ctx.handle, nbytes, C.byref(p), C.byref(fd), C.byref(size)))
        self.ptr, self.fd, self.nbytes = p.value, fd.value, size.value

    @classmethod
    def from_fd(cls, ctx: _lib.Context, fd: int, nbytes: int) -> "InteropBuffer":
        """Map a buffer another context or process exported (`nbytes` = its `nbytes`: the allocated size)."""
        self = cls.__new__(cls)
        self.ctx = ctx
        p = C.c_void_p()
        ctx.check(_lib.lib().ctc_interop_import(ctx.handle, fd, nbytes, C.byref(p)))
        self.ptr, self.fd, self.nbytes = p.value, -1, nbytes
        return self

    def read(self, offset: int, nbytes: int, dtype=np.uint8) -> np.ndarray:
        """Synchronous device -> host read of a byte range (tests, non-renderer consumers)."""
        assert 0 <= offset and offset + nbytes <= self.nbytes
        out = np.empty(nbytes, dtype=np.uint8)
        self.ctx.check(_lib.lib().ctc_device_read(self.ctx.handle, out.ctypes.data, self.ptr + offset, nbytes))
        return out.view(dtype)

    def close(self):
        if getattr(self, "ptr", None):
            self.ctx.check(_lib.lib().ctc_interop_free(self.ctx.handle, self.ptr))
            self.ptr = None
        if getattr(self, "fd", -1) >= 0:
            os.close(self.fd)
            self.fd = -1

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


@dataclass(frozen=True)
class MeshView:
    """MeshView{vbuf, ibuf, num_indices} (view.rs:14-18) as ranges of the batch's two interop buffers."""
    vertex_offset: int       # bytes into the vertex interop buffer
    num_vertices: int
    index_offset: int        # bytes into the index interop buffer
    num_indices: int


@dataclass
class MeshViews:
    """Every span's MeshView of one batched call + the two buffers they point into."""
    vbuf: InteropBuffer
    ibuf: InteropBuffer
    v_off: np.ndarray        # u64 [nspans + 1], vertices
    i_off: np.ndarray        # u64 [nspans + 1], indices

    def __len__(self) -> int:
        return len(self.v_off) - 1

    def view(self, k: int) -> MeshView:
        v0, v1, i0, i1 = int(self.v_off[k]), int(self.v_off[k + 1]), int(self.i_off[k]), int(self.i_off[k + 1])
        return MeshView(v0 * VERTEX_DTYPE.itemsize, v1 - v0, i0 * 4, i1 - i0)

    def download(self, k: int):
        """(vertices, indices) of span k read back from the interop buffers (what the renderer would draw)."""
        w = self.view(k)
        return (self.vbuf.read(w.vertex_offset, w.num_vertices * VERTEX_DTYPE.itemsize, VERTEX_DTYPE),
                self.ibuf.read(w.index_offset, w.num_indices * 4, np.uint32))


def _device_tables(ctx: _lib.Context, nbytes: int) -> int:
    """Device scratch for the two offset tables of a call, kept with the context and grown on demand."""
    have = getattr(ctx, "_view_tables", None)
    if have is None or have[1] < nbytes:
        L = _lib.lib()
        if have is not None:
            ctx.check(L.ctc_device_free(ctx.handle, have[0]))
        p = C.c_void_p()
        ctx.check(L.ctc_device_alloc(ctx.handle, 2 * nbytes, C.byref(p)))
        ctx._view_tables = have = (p.value, 2 * nbytes)
    return have[0]


def generate_views(spans, shape: Shape, resolution: int, ctx: _lib.Context | None = None,
                   vbuf: InteropBuffer | None = None, ibuf: InteropBuffer | None = None):
    """`generate_for_box` + `MeshView::new` for every span, in one call and without leaving the GPU
    (mesh/mod.rs:141-146) -> (MeshViews, Timings).  The only bytes that cross PCIe are the spans (24 B each) and the
    two offset tables (16 B per span).  Buffers that are missing are allocated (8 R^2 vertices per span, like
    generate_for_boxes); a buffer that turns out too small is replaced by one of the required size and the call repeated."""
    ctx = ctx or _lib.default_context()
    L = _lib.lib()
    arr = spans_array(spans)
    _check_args(arr, resolution)
    ns = arr.shape[0]
    vcap = max(1024, ns * 8 * resolution * resolution)
    icap = 6 * vcap
    sh = shape._ctc_shape()
    tables = _device_tables(ctx, 2 * (ns + 1) * 8)
    for _attempt in range(2):
        if vbuf is None:
            vbuf = InteropBuffer(ctx, vcap * VERTEX_DTYPE.itemsize)
        if ibuf is None:
            ibuf = InteropBuffer(ctx, icap * 4)
        d_voff, d_ioff = tables, tables + (ns + 1) * 8
        rc = L.ctc_mesh_spans_device(ctx.handle, C.byref(sh), arr.ctypes.data, ns, resolution, vbuf.ptr,
                                     vbuf.nbytes // VERTEX_DTYPE.itemsize, ibuf.ptr, ibuf.nbytes // 4, d_voff, d_ioff)
        if rc == _lib.CTC_ERR_INVALID_ARGUMENT:
            raise AssertionError(ctx.last_error())
        ctx.check(rc)
        nv, ni, t = C.c_uint64(), C.c_uint64(), _lib.CtcTimings()
        rc = L.ctc_mesh_result(ctx.handle, C.byref(nv), C.byref(ni), C.byref(t))
        if rc == _lib.CTC_ERR_OVERFLOW:          # the library reports what it needs: re-allocate what is too small
            vcap, icap = int(nv.value), int(ni.value)
            if vbuf.nbytes < vcap * VERTEX_DTYPE.itemsize:
                vbuf = None
            if ibuf.nbytes < icap * 4:
                ibuf = None
            continue
        if rc == _lib.CTC_ERR_LERP_ASSERT:
            raise AssertionError(ctx.last_error())
        ctx.check(rc)
        break
    else:
        raise _lib.CantucciError(_lib.CTC_ERR_OVERFLOW, "output still too small after retry")
    off = np.empty(2 * (ns + 1), dtype=np.uint64)
    ctx.check(L.ctc_device_read(ctx.handle, off.ctypes.data, tables, off.nbytes))
    timings = Timings(t.first_ms, t.second_ms, t.third_ms, int(t.vertices), int(t.faces))
    return MeshViews(vbuf, ibuf, off[:ns + 1].copy(), off[ns + 1:].copy()), timings
