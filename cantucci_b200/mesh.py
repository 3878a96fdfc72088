"""Host-side mirror of `MeshBuffer::generate_for_box` backed by the CUDA path.

Mirrors /root/reference/src/mesh/buffer.rs:24-42 (MeshBuffer), :398-434
(Timings) and the batching point of src/mesh/mod.rs:129-161 (all empty leaves
known up front -> one batched call).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from .octree import Span, spans_array
from .shape import Shape

VERTEX_DTYPE = _lib.VERTEX_DTYPE


@dataclass
class Timings:
    """mesh::buffer::Timings (buffer.rs:398-405); durations in milliseconds of device time."""
    first: float = 0.0
    second: float = 0.0
    third: float = 0.0
    vertices: int = 0
    faces: int = 0

    def __add__(self, o: "Timings") -> "Timings":     # buffer.rs:423-434
        return Timings(self.first + o.first, self.second + o.second, self.third + o.third,
                       self.vertices + o.vertices, self.faces + o.faces)


@dataclass
class MeshBuffer:
    """MeshBuffer{vertices: Vec<Vertex>, indices: Vec<u32>} (buffer.rs:24-27)."""
    vertices: np.ndarray     # VERTEX_DTYPE, 28-byte records
    indices: np.ndarray      # u32, span-local

    @staticmethod
    def generate_for_box(span: Span, shape: Shape, resolution: int, ctx: _lib.Context | None = None):
        """buffer.rs:30-42 -> (MeshBuffer, Timings).  Asserts like the reference."""
        batch, t = generate_for_boxes([span], shape, resolution, ctx=ctx)
        return batch.mesh(0), t


@dataclass
class MeshBatch:
    """The meshes of many spans in two flat buffers + offset tables."""
    vertices: np.ndarray
    indices: np.ndarray
    v_off: np.ndarray        # u64 [nspans+1]
    i_off: np.ndarray        # u64 [nspans+1]
    slot: np.ndarray | None = None   # meshed in another order than the caller's: slot[k] = place of span k in the tables

    def __len__(self) -> int:
        return len(self.v_off) - 1

    def mesh(self, k: int) -> MeshBuffer:
        if self.slot is not None:
            k = int(self.slot[k])
        return MeshBuffer(self.vertices[int(self.v_off[k]):int(self.v_off[k + 1])],
                          self.indices[int(self.i_off[k]):int(self.i_off[k + 1])])


def _check_args(spans: np.ndarray, resolution: int):
    # buffer.rs:35-39 and grid.rs:25
    assert np.all(spans[:, 0:3] < spans[:, 3:6]), "span.start < span.end"
    assert resolution != 0
    assert resolution & (resolution - 1) == 0, "resolution.is_power_of_two()"
    assert resolution >= 2, "GridTable size >= 2"


def sample_grids(spans, shape: Shape, resolution: int, ctx: _lib.Context | None = None) -> np.ndarray:
    """Pass 1 only (buffer.rs:64-83): [nspans, (R+1)^3] f32, x-major z-fastest."""
    ctx = ctx or _lib.default_context()
    arr = spans_array(spans)
    _check_args(arr, resolution)
    n = resolution + 1
    out = np.empty((arr.shape[0], n * n * n), dtype=np.float32)
    sh = shape._ctc_shape()
    ctx.check(_lib.lib().ctc_sample_grids(ctx.handle, C.byref(sh), arr.ctypes.data, arr.shape[0], resolution,
                                          out.ctypes.data))
    return out


def sample_signs(spans, shape: Shape, resolution: int, ctx: _lib.Context | None = None) -> np.ndarray:
    """Sign bit-planes of the spans' sample grids: [nspans, ((R+1)^3 + 31) // 32] u32, bit j of a plane =
    not is_sign_positive(grid[j]) (parity aid: the mesher's topology is a function of these bits)."""
    ctx = ctx or _lib.default_context()
    arr = spans_array(spans)
    _check_args(arr, resolution)
    words = ((resolution + 1) ** 3 + 31) // 32
    out = np.zeros((arr.shape[0], words), dtype=np.uint32)
    sh = shape._ctc_shape()
    ctx.check(_lib.lib().ctc_sample_signs(ctx.handle, C.byref(sh), arr.ctypes.data, arr.shape[0], resolution,
                                          out.ctypes.data))
    return out


def cull_spans(spans, shape: Shape, resolution: int, ctx: _lib.Context | None = None, safety: float = 2.0) -> np.ndarray:
    """DE-bound span culling (ctc_cull_spans): boolean mask of the spans that still need meshing.  A span is
    dropped when the distance estimate at its centre exceeds `safety` x the half-diagonal of its skirt-expanded
    box (and, for shapes with an upper bound, when it lies entirely inside).  Not part of the reference; never
    applied implicitly."""
    ctx = ctx or _lib.default_context()
    arr = spans_array(spans)
    _check_args(arr, resolution)
    keep = np.ones(arr.shape[0], dtype=np.uint8)
    sh = shape._ctc_shape()
    ctx.check(_lib.lib().ctc_cull_spans(ctx.handle, C.byref(sh), arr.ctypes.data, arr.shape[0], resolution,
                                        C.c_float(safety), keep.ctypes.data))
    return keep.astype(bool)


def order_spans(spans, shape: Shape, resolution: int, ctx: _lib.Context | None = None) -> np.ndarray:
    """Cost-aware span order (ctc_order_spans): indices of the spans, the ones most likely to hold surface first
    (ascending |DE(centre)| / half-diagonal), the provably empty ones last.  Meshing in this order keeps the copy
    pipeline behind the launch groups busy from the start and leaves no copy tail.  Not part of the reference, whose
    thread pool takes the jobs in tree order (mesh/mod.rs:129-161); never applied implicitly."""
    ctx = ctx or _lib.default_context()
    arr = spans_array(spans)
    _check_args(arr, resolution)
    order = np.zeros(arr.shape[0], dtype=np.uint32)
    sh = shape._ctc_shape()
    ctx.check(_lib.lib().ctc_order_spans(ctx.handle, C.byref(sh), arr.ctypes.data, arr.shape[0], resolution, order.ctypes.data))
    return order.astype(np.int64)


def generate_for_boxes(spans, shape: Shape, resolution: int, ctx: _lib.Context | None = None,
                       vcap: int | None = None, icap: int | None = None, out_v: np.ndarray | None = None,
                       out_i: np.ndarray | None = None, cull: bool = False, surface_first: bool = False):
    """generate_for_box for every span in ONE batched call -> (MeshBatch, Timings).

    Output capacity is guessed from the resolution and retried with the exact
    required size when the library reports CTC_ERR_OVERFLOW.  A lerp-factor
    failure raises AssertionError like the reference's panic (math.rs:19).
    cull = True (not in the reference): spans the DE bound proves empty (cull_spans) are not meshed at all;
    they come back as empty meshes.
    surface_first = True (not in the reference): the spans are meshed in order_spans' order; `mesh(k)` still is the
    mesh of the caller's k-th span (MeshBatch.slot maps it to its place in the buffers)."""
    ctx = ctx or _lib.default_context()
    arr = spans_array(spans)
    _check_args(arr, resolution)
    if surface_first and arr.shape[0] > 1:
        order = order_spans(arr, shape, resolution, ctx)
        sub, t = generate_for_boxes(np.ascontiguousarray(arr[order]), shape, resolution, ctx, vcap, icap, out_v, out_i, cull)
        slot = np.empty_like(order)
        slot[order] = np.arange(order.shape[0])
        return MeshBatch(sub.vertices, sub.indices, sub.v_off, sub.i_off, slot), t
    if cull and arr.shape[0]:
        keep = cull_spans(arr, shape, resolution, ctx)
        if not keep.all():
            sub, t = generate_for_boxes(np.ascontiguousarray(arr[keep]), shape, resolution, ctx, vcap, icap, out_v, out_i)
            # offsets of a culled span = those of the next kept one (an empty range)
            pos = np.concatenate([[0], np.cumsum(keep)]).astype(np.int64)
            return MeshBatch(sub.vertices, sub.indices, sub.v_off[pos], sub.i_off[pos]), t
    ns = arr.shape[0]
    if vcap is None:
        vcap = max(1024, ns * 8 * resolution * resolution)
    if icap is None:
        icap = 6 * max(1024, ns * 8 * resolution * resolution)
    sh = shape._ctc_shape()
    v_off = np.zeros(ns + 1, dtype=np.uint64)
    i_off = np.zeros(ns + 1, dtype=np.uint64)
    t = _lib.CtcTimings()
    for _attempt in range(2):
        v = out_v if (out_v is not None and len(out_v) >= vcap) else np.empty(vcap, dtype=VERTEX_DTYPE)
        idx = out_i if (out_i is not None and len(out_i) >= icap) else np.empty(icap, dtype=np.uint32)
        rc = _lib.lib().ctc_mesh_spans(ctx.handle, C.byref(sh), arr.ctypes.data, ns, resolution,
                                       v.ctypes.data, vcap, idx.ctypes.data, icap,
                                       v_off.ctypes.data, i_off.ctypes.data, C.byref(t))
        if rc == _lib.CTC_ERR_OVERFLOW:
            vcap, icap = int(v_off[ns]), int(i_off[ns])
            out_v = out_i = None
            continue
        if rc == _lib.CTC_ERR_LERP_ASSERT:
            raise AssertionError(ctx.last_error())
        if rc == _lib.CTC_ERR_INVALID_ARGUMENT:
            raise AssertionError(ctx.last_error())
        ctx.check(rc)
        break
    else:
        raise _lib.CantucciError(_lib.CTC_ERR_OVERFLOW, "output still too small after retry")
    nv, ni = int(v_off[ns]), int(i_off[ns])
    timings = Timings(t.first_ms, t.second_ms, t.third_ms, int(t.vertices), int(t.faces))
    return MeshBatch(v[:nv], idx[:ni], v_off, i_off), timings


@dataclass
class MultiMeshBatch:
    """The meshes of many spans as ctc_mesh_spans_multi delivers them: flat buffers that are DEVICE-major
    (every GPU fills its own region) and per-span [begin, end) tables in the caller's span order."""
    vertices: np.ndarray     # VERTEX_DTYPE (host destination) or None (device destination)
    indices: np.ndarray
    span_v: np.ndarray       # u64 [nspans, 2]
    span_i: np.ndarray       # u64 [nspans, 2]

    def __len__(self) -> int:
        return self.span_v.shape[0]

    def mesh(self, k: int) -> MeshBuffer:
        return MeshBuffer(self.vertices[int(self.span_v[k, 0]):int(self.span_v[k, 1])],
                          self.indices[int(self.span_i[k, 0]):int(self.span_i[k, 1])])


def generate_for_boxes_multi(spans, shape: Shape, resolution: int, multi: "_lib.MultiContext",
                             vcap: int | None = None, icap: int | None = None):
    """generate_for_box for every span, sharded over the GPUs of `multi` (ONE process: the span scheduler
    behind the C ABI, ctc_mesh_spans_multi) -> (MultiMeshBatch, Timings).  Host destination; capacity is
    guessed and the call is retried with the `need` the library reports on CTC_ERR_OVERFLOW."""
    arr = spans_array(spans)
    _check_args(arr, resolution)
    ns = arr.shape[0]
    if vcap is None:
        vcap = max(4096, ns * 8 * resolution * resolution)
    if icap is None:
        icap = 6 * max(4096, ns * 8 * resolution * resolution)
    sh = shape._ctc_shape()
    span_v = np.zeros((ns, 2), dtype=np.uint64)
    span_i = np.zeros((ns, 2), dtype=np.uint64)
    need = (C.c_uint64 * 2)()
    t = _lib.CtcTimings()
    for _attempt in range(2):
        v = np.empty(vcap, dtype=VERTEX_DTYPE)
        idx = np.empty(icap, dtype=np.uint32)
        rc = _lib.lib().ctc_mesh_spans_multi(multi.handle, C.byref(sh), arr.ctypes.data, ns, resolution,
                                             v.ctypes.data, vcap, idx.ctypes.data, icap,
                                             span_v.ctypes.data, span_i.ctypes.data, need, C.byref(t))
        if rc == _lib.CTC_ERR_OVERFLOW:
            vcap, icap = int(need[0]), int(need[1])
            continue
        if rc in (_lib.CTC_ERR_LERP_ASSERT, _lib.CTC_ERR_INVALID_ARGUMENT):
            raise AssertionError(multi.last_error())
        multi.check(rc)
        break
    else:
        raise _lib.CantucciError(_lib.CTC_ERR_OVERFLOW, "output still too small after retry")
    timings = Timings(t.first_ms, t.second_ms, t.third_ms, int(t.vertices), int(t.faces))
    return MultiMeshBatch(v, idx, span_v, span_i), timings
