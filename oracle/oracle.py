"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product package cantucci_b200
never does.  PARITY UNPINNED -- see oracle/cantucci_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcantucci_oracle.so")

VERTEX_DTYPE = np.dtype(
    [("position", "<f4", (3,)), ("normal", "<f4", (3,)), ("distance_from_surface", "<f4")]
)
assert VERTEX_DTYPE.itemsize == 28


class Span(C.Structure):
    _fields_ = [("start", C.c_float * 3), ("end", C.c_float * 3)]


class Shape(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("power", C.c_uint32),
        ("max_iters", C.c_uint64),
        ("bailout", C.c_float),
        ("center", C.c_float * 3),
        ("radius", C.c_float),
    ]


class DeInfo(C.Structure):
    _fields_ = [
        ("iters", C.c_uint32),
        ("bailed", C.c_uint32),
        ("r", C.c_float),
        ("dr", C.c_float),
        ("min_margin", C.c_float),
    ]


class Mesh(C.Structure):
    _fields_ = [
        ("vertices", C.c_void_p),
        ("indices", C.c_void_p),
        ("n_vertices", C.c_uint64),
        ("n_indices", C.c_uint64),
        ("first_s", C.c_double),
        ("second_s", C.c_double),
        ("third_s", C.c_double),
        ("panicked", C.c_int32),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (recipe: oracle/Makefile)."""
    src = os.path.join(_HERE, "cantucci_oracle.c")
    hdr = os.path.join(_HERE, "cantucci_oracle.h")
    stale = (not os.path.exists(_SO)) or any(
        os.path.exists(f) and os.path.getmtime(f) > os.path.getmtime(_SO) for f in (src, hdr)
    )
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libcantucci_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        fp = C.POINTER(C.c_float)
        L.orc_min_distance_from.restype = C.c_float
        L.orc_min_distance_from.argtypes = [C.POINTER(Shape), fp]
        L.orc_min_distance_from_info.restype = C.c_float
        L.orc_min_distance_from_info.argtypes = [C.POINTER(Shape), fp, C.POINTER(DeInfo)]
        L.orc_batch_min_distance_from.restype = None
        L.orc_batch_min_distance_from.argtypes = [C.POINTER(Shape), C.c_void_p, C.c_size_t, C.c_void_p]
        for name in ("orc_rotate_p8_scalar",):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [fp, fp]
        for name in ("orc_rotate_generic", "orc_rotate"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [C.c_uint32, fp, fp]
        L.orc_logf_glibc_fma.restype = C.c_float
        L.orc_logf_glibc_fma.argtypes = [C.c_float]
        L.orc_sample_grid.restype = C.c_int
        L.orc_sample_grid.argtypes = [C.POINTER(Shape), C.POINTER(Span), C.c_uint32, C.c_void_p]
        L.orc_sample_grid_info.restype = C.c_int
        L.orc_sample_grid_info.argtypes = [C.POINTER(Shape), C.POINTER(Span), C.c_uint32, C.c_void_p,
                                           C.c_void_p, C.c_void_p]
        L.orc_sample_grid_iters.restype = C.c_int
        L.orc_sample_grid_iters.argtypes = [C.POINTER(Shape), C.POINTER(Span), C.c_uint32, C.c_void_p, C.c_void_p]
        L.orc_generate_for_box.restype = C.c_int
        L.orc_generate_for_box.argtypes = [C.POINTER(Shape), C.POINTER(Span), C.c_uint32, C.POINTER(Mesh)]
        L.orc_mesh_free.restype = None
        L.orc_mesh_free.argtypes = [C.POINTER(Mesh)]
        L.orc_generate_for_boxes_mt.restype = C.c_double
        L.orc_generate_for_boxes_mt.argtypes = [C.POINTER(Shape), C.c_void_p, C.c_size_t, C.c_uint32,
                                                C.c_int, C.c_void_p]
        L.orc_generate_for_boxes_signs_mt.restype = C.c_double
        L.orc_generate_for_boxes_signs_mt.argtypes = [C.POINTER(Shape), C.c_void_p, C.c_size_t, C.c_uint32,
                                                      C.c_int, C.c_void_p, C.c_void_p]
        L.orc_sample_grids_mt.restype = C.c_double
        L.orc_sample_grids_mt.argtypes = [C.POINTER(Shape), C.c_void_p, C.c_size_t, C.c_uint32, C.c_int,
                                          C.POINTER(C.c_double)]
        L.orc_span_center.restype = None
        L.orc_span_center.argtypes = [C.POINTER(Span), fp]
        L.orc_create_spans.restype = None
        L.orc_create_spans.argtypes = [C.POINTER(Span), C.POINTER(Span)]
        L.orc_hardware_threads.restype = C.c_int
        _lib = L
    return _lib


# --------------------------------------------------------------------------
# pythonic helpers
# --------------------------------------------------------------------------

def mandelbulb(power: int = 8, max_iters: int = 6, bailout: float = 2.5) -> Shape:
    assert max_iters >= 1  # Mandelbulb::new (mandelbulb.rs:20)
    s = Shape()
    s.kind, s.power, s.max_iters, s.bailout = 0, power, max_iters, bailout
    return s


def sphere(center=(0.0, 0.0, 0.0), radius: float = 1.0) -> Shape:
    s = Shape()
    s.kind = 1
    s.center[:] = center
    s.radius = radius
    return s


def make_span(start, end) -> Span:
    sp = Span()
    sp.start[:] = [float(np.float32(v)) for v in start]
    sp.end[:] = [float(np.float32(v)) for v in end]
    return sp


def spans_to_array(spans) -> np.ndarray:
    """list[Span] | ndarray -> float32 [n, 6] (start xyz, end xyz)."""
    if isinstance(spans, np.ndarray):
        return np.ascontiguousarray(spans, dtype=np.float32).reshape(-1, 6)
    return np.array([[*s.start, *s.end] for s in spans], dtype=np.float32).reshape(-1, 6)


def _f3(p):
    return (C.c_float * 3)(*[float(np.float32(v)) for v in p])


def min_distance_from(shape: Shape, p) -> float:
    return lib().orc_min_distance_from(C.byref(shape), _f3(p))


def min_distance_from_info(shape: Shape, p):
    info = DeInfo()
    d = lib().orc_min_distance_from_info(C.byref(shape), _f3(p), C.byref(info))
    return d, info


def batch_min_distance_from(shape: Shape, xyz: np.ndarray) -> np.ndarray:
    xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
    out = np.empty(xyz.shape[0], dtype=np.float32)
    lib().orc_batch_min_distance_from(C.byref(shape), xyz.ctypes.data, xyz.shape[0], out.ctypes.data)
    return out


def rotate(power: int, p, variant: str = "dispatch") -> np.ndarray:
    out = (C.c_float * 3)()
    if variant == "p8_scalar":
        lib().orc_rotate_p8_scalar(_f3(p), out)
    elif variant == "generic":
        lib().orc_rotate_generic(power, _f3(p), out)
    else:
        lib().orc_rotate(power, _f3(p), out)
    return np.array(out[:], dtype=np.float32)


def sample_grid(shape: Shape, span: Span, resolution: int, with_info: bool = False):
    n = resolution + 1
    out = np.empty(n * n * n, dtype=np.float32)
    if with_info:
        hist = np.zeros(int(shape.max_iters) + 1, dtype=np.uint64)
        nb = np.zeros(1, dtype=np.uint64)
        rc = lib().orc_sample_grid_info(C.byref(shape), C.byref(span), resolution, out.ctypes.data,
                                        hist.ctypes.data, nb.ctypes.data)
        if rc:
            raise AssertionError("generate_for_box argument assertion (buffer.rs:35-39)")
        return out, hist, int(nb[0])
    rc = lib().orc_sample_grid(C.byref(shape), C.byref(span), resolution, out.ctypes.data)
    if rc:
        raise AssertionError("generate_for_box argument assertion (buffer.rs:35-39)")
    return out


def sample_grid_iters(shape: Shape, span: Span, resolution: int):
    """-> (distances, completed iterations per sample [u8, saturating])."""
    n = resolution + 1
    out = np.empty(n * n * n, dtype=np.float32)
    it = np.empty(n * n * n, dtype=np.uint8)
    rc = lib().orc_sample_grid_iters(C.byref(shape), C.byref(span), resolution, out.ctypes.data, it.ctypes.data)
    if rc:
        raise AssertionError("generate_for_box argument assertion (buffer.rs:35-39)")
    return out, it


def _mesh_out(m: Mesh):
    if m.panicked:
        return None
    nv, ni = int(m.n_vertices), int(m.n_indices)
    v = np.empty(nv, dtype=VERTEX_DTYPE)
    i = np.empty(ni, dtype=np.uint32)
    if nv:
        C.memmove(v.ctypes.data, m.vertices, nv * 28)
    if ni:
        C.memmove(i.ctypes.data, m.indices, ni * 4)
    return v, i, (m.first_s, m.second_s, m.third_s)


def generate_for_box(shape: Shape, span: Span, resolution: int):
    """-> (vertices[VERTEX_DTYPE], indices[u32], timings) ; raises like the
    reference panics (argument asserts, lerp assert)."""
    m = Mesh()
    rc = lib().orc_generate_for_box(C.byref(shape), C.byref(span), resolution, C.byref(m))
    if rc:
        raise AssertionError("generate_for_box argument assertion (buffer.rs:35-39)")
    try:
        out = _mesh_out(m)
    finally:
        lib().orc_mesh_free(C.byref(m))
    if out is None:
        raise AssertionError("lerp factor assertion (math.rs:19)")
    return out


def generate_for_boxes_mt(shape: Shape, spans, resolution: int, nthreads: int | None = None):
    """Thread-pool run (mesh/mod.rs:61-62,141). -> (list of meshes|None, wall seconds)."""
    arr = spans_to_array(spans)
    n = arr.shape[0]
    meshes = (Mesh * n)()
    if nthreads is None:
        nthreads = lib().orc_hardware_threads()
    secs = lib().orc_generate_for_boxes_mt(C.byref(shape), arr.ctypes.data, n, resolution, nthreads,
                                           C.addressof(meshes))
    out = []
    for k in range(n):
        out.append(_mesh_out(meshes[k]))
        lib().orc_mesh_free(C.byref(meshes[k]))
    return out, secs


def generate_for_boxes_flat_mt(shape: Shape, spans, resolution: int, nthreads: int | None = None, signs: bool = True):
    """Thread-pool run returning FLAT arrays (the parity gate's checker): vertices [V] VERTEX_DTYPE,
    indices [I] u32, v_off / i_off [n+1] (spans that panicked contribute nothing and are listed in
    `panicked`), the sign bit-planes [n, words] u32 (or None) and the wall seconds of the pool."""
    arr = spans_to_array(spans)
    n = arr.shape[0]
    meshes = (Mesh * n)()
    if nthreads is None:
        nthreads = lib().orc_hardware_threads()
    words = ((resolution + 1) ** 3 + 31) // 32
    planes = np.zeros((n, words), dtype=np.uint32) if signs else None
    secs = lib().orc_generate_for_boxes_signs_mt(C.byref(shape), arr.ctypes.data, n, resolution, nthreads,
                                                 C.addressof(meshes), planes.ctypes.data if signs else None)
    nv = np.array([m.n_vertices for m in meshes], dtype=np.int64)
    ni = np.array([m.n_indices for m in meshes], dtype=np.int64)
    v_off = np.concatenate([[0], np.cumsum(nv)])
    i_off = np.concatenate([[0], np.cumsum(ni)])
    v = np.empty(int(v_off[-1]), dtype=VERTEX_DTYPE)
    i = np.empty(int(i_off[-1]), dtype=np.uint32)
    panicked = []
    for k in range(n):
        m = meshes[k]
        if m.panicked:
            panicked.append(k)
        if m.n_vertices:
            C.memmove(v.ctypes.data + int(v_off[k]) * 28, m.vertices, int(m.n_vertices) * 28)
        if m.n_indices:
            C.memmove(i.ctypes.data + int(i_off[k]) * 4, m.indices, int(m.n_indices) * 4)
        lib().orc_mesh_free(C.byref(meshes[k]))
    return v, i, v_off, i_off, planes, panicked, secs


def sample_grids_mt(shape: Shape, spans, resolution: int, nthreads: int | None = None):
    arr = spans_to_array(spans)
    if nthreads is None:
        nthreads = lib().orc_hardware_threads()
    cs = C.c_double(0.0)
    secs = lib().orc_sample_grids_mt(C.byref(shape), arr.ctypes.data, arr.shape[0], resolution, nthreads,
                                     C.byref(cs))
    return secs, cs.value


def create_spans(parent: Span):
    out = (Span * 8)()
    lib().orc_create_spans(C.byref(parent), out)
    return [make_span(s.start[:], s.end[:]) for s in out]


def span_center(span: Span) -> np.ndarray:
    out = (C.c_float * 3)()
    lib().orc_span_center(C.byref(span), out)
    return np.array(out[:], dtype=np.float32)


def hardware_threads() -> int:
    return lib().orc_hardware_threads()
