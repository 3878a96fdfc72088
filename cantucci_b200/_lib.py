"""ctypes binding of libcantucci_b200.so (the C ABI in include/cantucci_b200.h).

There is no CPU fallback: if the CUDA library has not been built, or no CUDA
device is present, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# CANTUCCI_B200_LIB selects another build of the same library (A/B kernel experiments)
LIB_PATH = os.environ.get("CANTUCCI_B200_LIB") or os.path.join(_HERE, "libcantucci_b200.so")

CTC_OK = 0
CTC_ERR_INVALID_ARGUMENT = 1
CTC_ERR_CUDA = 2
CTC_ERR_OVERFLOW = 3
CTC_ERR_LERP_ASSERT = 4
CTC_ERR_NO_DEVICE = 5

CTC_SHAPE_MANDELBULB = 0
CTC_SHAPE_SPHERE = 1
CTC_MATH_EXACT = 0
CTC_MATH_FAST = 1

# mesh::Vertex (src/mesh/mod.rs:255-261)
VERTEX_DTYPE = np.dtype(
    [("position", "<f4", (3,)), ("normal", "<f4", (3,)), ("distance_from_surface", "<f4")]
)
assert VERTEX_DTYPE.itemsize == 28


class CtcSpan(C.Structure):
    _fields_ = [("start", C.c_float * 3), ("end", C.c_float * 3)]


class CtcShape(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("power", C.c_uint32),
        ("max_iters", C.c_uint64),
        ("bailout", C.c_float),
        ("center", C.c_float * 3),
        ("radius", C.c_float),
        ("flags", C.c_uint32),
    ]


class CtcTimings(C.Structure):
    _fields_ = [
        ("first_ms", C.c_double),
        ("second_ms", C.c_double),
        ("third_ms", C.c_double),
        ("vertices", C.c_uint64),
        ("faces", C.c_uint64),
    ]


# every symbol include/cantucci_b200.h declares
EXPORTS = (
    "ctc_version", "ctc_device_count", "ctc_ctx_create", "ctc_ctx_destroy", "ctc_ctx_set_stream",
    "ctc_ctx_set_group_spans", "ctc_ctx_set_overlap", "ctc_ctx_synchronize", "ctc_last_error", "ctc_kernel_launches",
    "ctc_de_batch", "ctc_de_batch_device", "ctc_sample_grids", "ctc_sample_grids_device",
    "ctc_mesh_spans", "ctc_mesh_spans_device", "ctc_mesh_result",
    "ctc_iteration_stats", "ctc_fp32_peak_probe",
    "ctc_ray_march", "ctc_device_alloc", "ctc_device_free", "ctc_ipc_export", "ctc_ipc_open", "ctc_ipc_close",
    "ctc_host_register", "ctc_host_unregister", "ctc_ctx_set_index_wire", "ctc_expand_quads",
    "ctc_ctx_set_fast_band", "ctc_mesh_fixups", "ctc_fast_sign_probe", "ctc_sample_signs",
    "ctc_ctx_set_kernel_timing", "ctc_mesh_kernel_times", "ctc_iteration_stats_points",
    "ctc_ctx_set_coalescing", "ctc_ctx_coalescing_stats", "ctc_ctx_set_host_index_wire", "ctc_ctx_host_index_wire_stats", "ctc_ctx_set_wire_progress", "ctc_cull_spans", "ctc_expand_quads_host", "ctc_render", "ctc_render_device",
    "ctc_last_error_copy", "ctc_multi_create", "ctc_multi_destroy", "ctc_multi_ngpus", "ctc_multi_ctx",
    "ctc_multi_last_error", "ctc_mesh_spans_multi", "ctc_mesh_spans_multi_device", "ctc_multi_shard_plan",
    "ctc_interop_alloc", "ctc_interop_import", "ctc_interop_free", "ctc_device_read", "ctc_device_write", "ctc_order_spans", "ctc_ctx_set_host_wire_share", "ctc_mesh_d2h_bytes",
)

_lib = None


class CantucciError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"cantucci_b200 error {code}: {msg}")
        self.code = code


def lib() -> C.CDLL:
    """Load the CUDA library; fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m cantucci_b200.build` "
            "(there is no CPU fallback)"
        )
    L = C.CDLL(LIB_PATH)
    vp, sz, u32, u64p = C.c_void_p, C.c_size_t, C.c_uint32, C.POINTER(C.c_uint64)
    shp, spn = C.POINTER(CtcShape), C.c_void_p
    L.ctc_version.restype = C.c_int
    L.ctc_device_count.restype = C.c_int
    L.ctc_ctx_create.restype = C.c_int
    L.ctc_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.ctc_ctx_destroy.restype = None
    L.ctc_ctx_destroy.argtypes = [vp]
    L.ctc_ctx_set_stream.restype = C.c_int
    L.ctc_ctx_set_stream.argtypes = [vp, vp]
    L.ctc_ctx_set_group_spans.restype = C.c_int
    L.ctc_ctx_set_group_spans.argtypes = [vp, u32]
    L.ctc_ctx_set_overlap.restype = C.c_int
    L.ctc_ctx_set_overlap.argtypes = [vp, C.c_int]
    L.ctc_ctx_synchronize.restype = C.c_int
    L.ctc_ctx_synchronize.argtypes = [vp]
    L.ctc_last_error.restype = C.c_char_p
    L.ctc_last_error.argtypes = [vp]
    L.ctc_kernel_launches.restype = C.c_uint64
    L.ctc_kernel_launches.argtypes = [vp]
    for name in ("ctc_de_batch", "ctc_de_batch_device"):
        f = getattr(L, name)
        f.restype = C.c_int
        f.argtypes = [vp, shp, vp, sz, vp]
    for name in ("ctc_sample_grids", "ctc_sample_grids_device"):
        f = getattr(L, name)
        f.restype = C.c_int
        f.argtypes = [vp, shp, spn, sz, u32, vp]
    L.ctc_mesh_spans.restype = C.c_int
    L.ctc_mesh_spans.argtypes = [vp, shp, spn, sz, u32, vp, sz, vp, sz, vp, vp, C.POINTER(CtcTimings)]
    L.ctc_mesh_spans_device.restype = C.c_int
    L.ctc_mesh_spans_device.argtypes = [vp, shp, spn, sz, u32, vp, sz, vp, sz, vp, vp]
    L.ctc_mesh_result.restype = C.c_int
    L.ctc_mesh_result.argtypes = [vp, u64p, u64p, C.POINTER(CtcTimings)]
    L.ctc_iteration_stats.restype = C.c_int
    L.ctc_iteration_stats.argtypes = [vp, shp, spn, sz, u32, u64p]
    L.ctc_iteration_stats_points.restype = C.c_int
    L.ctc_iteration_stats_points.argtypes = [vp, shp, vp, sz, u64p]
    L.ctc_ray_march.restype = C.c_int
    L.ctc_ray_march.argtypes = [vp, shp, vp, vp, sz, u32, C.c_float, vp, vp]
    L.ctc_device_alloc.restype = C.c_int
    L.ctc_device_alloc.argtypes = [vp, sz, C.POINTER(vp)]
    L.ctc_device_free.restype = C.c_int
    L.ctc_device_free.argtypes = [vp, vp]
    L.ctc_ipc_export.restype = C.c_int
    L.ctc_ipc_export.argtypes = [vp, vp, C.c_char_p]
    L.ctc_ipc_open.restype = C.c_int
    L.ctc_ipc_open.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
    L.ctc_ipc_close.restype = C.c_int
    L.ctc_ipc_close.argtypes = [vp, vp]
    L.ctc_interop_alloc.restype = C.c_int
    L.ctc_interop_alloc.argtypes = [vp, sz, C.POINTER(vp), C.POINTER(C.c_int), C.POINTER(sz)]
    L.ctc_interop_import.restype = C.c_int
    L.ctc_interop_import.argtypes = [vp, C.c_int, sz, C.POINTER(vp)]
    L.ctc_interop_free.restype = C.c_int
    L.ctc_interop_free.argtypes = [vp, vp]
    L.ctc_order_spans.restype = C.c_int
    L.ctc_order_spans.argtypes = [vp, shp, spn, sz, u32, vp]
    L.ctc_ctx_set_host_wire_share.restype = C.c_int
    L.ctc_ctx_set_host_wire_share.argtypes = [vp, u32, u32]
    L.ctc_mesh_d2h_bytes.restype = C.c_int
    L.ctc_mesh_d2h_bytes.argtypes = [vp, u64p]
    L.ctc_device_write.restype = C.c_int
    L.ctc_device_write.argtypes = [vp, vp, vp, sz]
    L.ctc_device_read.restype = C.c_int
    L.ctc_device_read.argtypes = [vp, vp, vp, sz]
    L.ctc_ctx_set_index_wire.restype = C.c_int
    L.ctc_ctx_set_index_wire.argtypes = [vp, C.c_int]
    L.ctc_expand_quads.restype = C.c_int
    L.ctc_expand_quads.argtypes = [vp, vp, sz, vp]
    L.ctc_host_register.restype = C.c_int
    L.ctc_host_register.argtypes = [vp, vp, sz]
    L.ctc_host_unregister.restype = C.c_int
    L.ctc_host_unregister.argtypes = [vp, vp]
    L.ctc_ctx_set_fast_band.restype = C.c_int
    L.ctc_ctx_set_fast_band.argtypes = [vp, C.c_float]
    L.ctc_mesh_fixups.restype = C.c_int
    L.ctc_mesh_fixups.argtypes = [vp, u64p, u64p]
    L.ctc_sample_signs.restype = C.c_int
    L.ctc_sample_signs.argtypes = [vp, shp, spn, sz, u32, vp]
    L.ctc_fast_sign_probe.restype = C.c_int
    L.ctc_fast_sign_probe.argtypes = [vp, shp, spn, sz, u32, u64p, sz]
    L.ctc_ctx_set_coalescing.restype = C.c_int
    L.ctc_ctx_set_coalescing.argtypes = [vp, C.c_int]
    L.ctc_render.restype = C.c_int
    L.ctc_render.argtypes = [vp, shp, vp, u32, u32, u32, C.c_float, vp]
    L.ctc_render_device.restype = C.c_int
    L.ctc_render_device.argtypes = [vp, shp, vp, u32, u32, u32, C.c_float, vp]
    L.ctc_expand_quads_host.restype = C.c_int
    L.ctc_expand_quads_host.argtypes = [vp, sz, vp]
    L.ctc_cull_spans.restype = C.c_int
    L.ctc_cull_spans.argtypes = [vp, vp, vp, sz, C.c_uint32, C.c_float, vp]
    L.ctc_ctx_set_wire_progress.restype = C.c_int
    L.ctc_ctx_set_wire_progress.argtypes = [vp, vp]
    L.ctc_ctx_set_host_index_wire.restype = C.c_int
    L.ctc_ctx_set_host_index_wire.argtypes = [vp, C.c_int]
    L.ctc_ctx_host_index_wire_stats.restype = C.c_int
    L.ctc_ctx_host_index_wire_stats.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
    L.ctc_ctx_coalescing_stats.restype = C.c_int
    L.ctc_ctx_coalescing_stats.argtypes = [vp, u64p, u64p]
    L.ctc_ctx_set_kernel_timing.restype = C.c_int
    L.ctc_ctx_set_kernel_timing.argtypes = [vp, C.c_int]
    L.ctc_mesh_kernel_times.restype = C.c_int
    L.ctc_mesh_kernel_times.argtypes = [vp, C.POINTER(C.c_double), sz]
    L.ctc_last_error_copy.restype = C.c_size_t
    L.ctc_last_error_copy.argtypes = [vp, C.c_char_p, sz]
    L.ctc_multi_create.restype = C.c_int
    L.ctc_multi_create.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]
    L.ctc_multi_destroy.restype = None
    L.ctc_multi_destroy.argtypes = [vp]
    L.ctc_multi_ngpus.restype = C.c_int
    L.ctc_multi_ngpus.argtypes = [vp]
    L.ctc_multi_ctx.restype = vp
    L.ctc_multi_ctx.argtypes = [vp, C.c_int]
    L.ctc_multi_last_error.restype = C.c_size_t
    L.ctc_multi_last_error.argtypes = [vp, C.c_char_p, sz]
    for name in ("ctc_mesh_spans_multi", "ctc_mesh_spans_multi_device"):
        f = getattr(L, name)
        f.restype = C.c_int
        f.argtypes = [vp, shp, spn, sz, u32, vp, sz, vp, sz, vp, vp, u64p, C.POINTER(CtcTimings)]
    L.ctc_multi_shard_plan.restype = C.c_int
    L.ctc_multi_shard_plan.argtypes = [sz, C.c_int, sz, sz, u64p, u64p, vp]
    L.ctc_fp32_peak_probe.restype = C.c_int
    L.ctc_fp32_peak_probe.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_int)]
    _lib = L
    return L


class Context:
    """Owns one ctc_ctx (one CUDA device + stream + workspace)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        rc = lib().ctc_ctx_create(device, C.byref(self._h))
        if rc != CTC_OK:
            self._h = C.c_void_p()
            raise CantucciError(rc, "ctc_ctx_create failed (no CUDA device? there is no CPU fallback)")
        self.device = device

    @property
    def handle(self):
        return self._h

    def check(self, rc: int):
        if rc != CTC_OK:
            raise CantucciError(rc, self.last_error())

    def last_error(self) -> str:
        buf = C.create_string_buffer(512)
        lib().ctc_last_error_copy(self._h, buf, 512)      # copied out under the context's lock
        return buf.value.decode(errors="replace")

    def set_stream(self, cuda_stream: int | None):
        self.check(lib().ctc_ctx_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def set_group_spans(self, n: int):
        self.check(lib().ctc_ctx_set_group_spans(self._h, n))

    def set_overlap(self, enable: bool):
        self.check(lib().ctc_ctx_set_overlap(self._h, 1 if enable else 0))

    def set_host_index_wire(self, mode):
        """Host destinations of ctc_mesh_spans: packed quad records over PCIe, widened by host threads.
        0 / False: off; 1 / True (default): calls of >= 128 spans; 2: every call."""
        self.check(lib().ctc_ctx_set_host_index_wire(self._h, int(mode)))

    def set_host_wire_share(self, num: int, den: int):
        """Of every `den` launch groups of a large host-buffer call, `num` ship their indices as packed records."""
        self.check(lib().ctc_ctx_set_host_wire_share(self._h, num, den))

    def mesh_d2h_bytes(self) -> int:
        b = C.c_uint64(0)
        self.check(lib().ctc_mesh_d2h_bytes(self._h, C.byref(b)))
        return int(b.value)

    def host_index_wire_stats(self):
        """(calls that used the packed wire, calls that fell back to u32 indices, widening threads)."""
        a, b, n = C.c_uint64(0), C.c_uint64(0), C.c_uint32(0)
        self.check(lib().ctc_ctx_host_index_wire_stats(self._h, C.byref(a), C.byref(b), C.byref(n)))
        return int(a.value), int(b.value), int(n.value)

    def synchronize(self):
        self.check(lib().ctc_ctx_synchronize(self._h))

    KERNELS = ("sample_grids", "fixup_suspects", "classify_count", "span_scan", "emit_lists", "vertex", "quads")

    def set_coalescing(self, enable: bool):
        self.check(lib().ctc_ctx_set_coalescing(self._h, 1 if enable else 0))

    def coalescing_stats(self) -> tuple[int, int]:
        """(batched launches so far, requests they served)."""
        a, b = C.c_uint64(0), C.c_uint64(0)
        self.check(lib().ctc_ctx_coalescing_stats(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def set_kernel_timing(self, enable: bool):
        self.check(lib().ctc_ctx_set_kernel_timing(self._h, 1 if enable else 0))

    def kernel_times(self) -> dict:
        """Device ms per kernel of the last fetched mesh call (needs set_kernel_timing(True))."""
        ms = (C.c_double * 8)()
        self.check(lib().ctc_mesh_kernel_times(self._h, ms, 8))
        return {name: float(ms[k]) for k, name in enumerate(self.KERNELS)}

    def set_fast_band(self, kappa: float = 0.0):
        """Fast mode's sign-trust band (0 = the calibrated default)."""
        self.check(lib().ctc_ctx_set_fast_band(self._h, kappa))

    def mesh_fixups(self) -> tuple[int, int]:
        """(samples re-evaluated exactly, of which changed sign) of the last fetched mesh result."""
        a, b = C.c_uint64(0), C.c_uint64(0)
        self.check(lib().ctc_mesh_fixups(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def kernel_launches(self) -> int:
        return int(lib().ctc_kernel_launches(self._h))

    def close(self):
        if self._h:
            lib().ctc_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx: dict[int, Context] = {}


def default_context(device: int = 0) -> Context:
    ctx = _default_ctx.get(device)
    if ctx is None:
        ctx = _default_ctx[device] = Context(device)
    return ctx


def shard_plan(nspans: int, ngpus: int, vcap: int, icap: int):
    """ctc_multi_shard_plan: (first_v [ngpus+1], first_i [ngpus+1], owner [nspans]) -- no GPU needed."""
    first_v = np.zeros(ngpus + 1, dtype=np.uint64)
    first_i = np.zeros(ngpus + 1, dtype=np.uint64)
    owner = np.zeros(max(nspans, 1), dtype=np.uint32)
    rc = lib().ctc_multi_shard_plan(nspans, ngpus, vcap, icap, first_v.ctypes.data_as(C.POINTER(C.c_uint64)),
                                    first_i.ctypes.data_as(C.POINTER(C.c_uint64)), owner.ctypes.data)
    if rc != CTC_OK:
        raise CantucciError(rc, "ctc_multi_shard_plan: invalid argument")
    return first_v, first_i, owner[:nspans]


class MultiContext:
    """Owns a ctc_multi: one context + worker thread per GPU of this box, all in THIS process."""

    def __init__(self, devices=None, ngpus: int = 0):
        self._h = C.c_void_p()
        arr = None
        if devices is not None:
            ngpus = len(devices)
            arr = (C.c_int * ngpus)(*devices)
        rc = lib().ctc_multi_create(arr, ngpus, C.byref(self._h))
        if rc != CTC_OK:
            self._h = C.c_void_p()
            raise CantucciError(rc, "ctc_multi_create failed (no CUDA device? there is no CPU fallback)")
        self.ngpus = int(lib().ctc_multi_ngpus(self._h))

    @property
    def handle(self):
        return self._h

    def last_error(self) -> str:
        buf = C.create_string_buffer(512)
        lib().ctc_multi_last_error(self._h, buf, 512)
        return buf.value.decode(errors="replace")

    def check(self, rc: int):
        if rc != CTC_OK:
            raise CantucciError(rc, self.last_error())

    def close(self):
        if self._h:
            lib().ctc_multi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
