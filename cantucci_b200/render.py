"""The ray-marcher the reference plans (README.md:12-16) on top of ctc_render: one sphere-traced ray per pixel,
the same loop as ShapeMesh::get_focii (src/mesh/mod.rs:229-241).  SURVEY 8f, N4."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .shape import Shape


class CtcCameraRays(C.Structure):
    _fields_ = [("eye", C.c_float * 3), ("top_left", C.c_float * 3), ("du", C.c_float * 3), ("dv", C.c_float * 3)]


def look_at_rays(eye, target, up, fov_y_deg: float, width: int, height: int) -> CtcCameraRays:
    """A pinhole camera as the four vectors ctc_render takes: the image plane sits one unit in front of the eye."""
    eye, target, up = (np.asarray(v, dtype=np.float64) for v in (eye, target, up))
    f = target - eye; f /= np.linalg.norm(f)
    r = np.cross(f, up); r /= np.linalg.norm(r)
    u = np.cross(r, f)
    half_h = np.tan(np.radians(fov_y_deg) / 2.0); half_w = half_h * width / height
    cam = CtcCameraRays()
    cam.eye[:] = eye.astype(np.float32)
    cam.top_left[:] = (eye + f - half_w * r + half_h * u).astype(np.float32)
    cam.du[:] = (2.0 * half_w / width * r).astype(np.float32)
    cam.dv[:] = (-2.0 * half_h / height * u).astype(np.float32)
    return cam


def pixel_rays(cam: CtcCameraRays, width: int, height: int):
    """The rays ctc_render forms, bit for bit, in numpy f32 (row-major pixels): (origins [n,3], directions [n,3])."""
    f = np.float32
    i = (np.arange(width, dtype=np.float32) + f(0.5))[None, :, None]
    j = (np.arange(height, dtype=np.float32) + f(0.5))[:, None, None]
    tl, du, dv, eye = (np.array(list(v), dtype=np.float32)[None, None, :] for v in (cam.top_left, cam.du, cam.dv, cam.eye))
    q = (tl + i * du).astype(np.float32) + (j * dv).astype(np.float32)
    d = (q - eye).astype(np.float32)
    n2 = ((d[..., 0] * d[..., 0]).astype(np.float32) + (d[..., 1] * d[..., 1]).astype(np.float32)).astype(np.float32)
    n2 = (n2 + (d[..., 2] * d[..., 2]).astype(np.float32)).astype(np.float32)
    inv = (f(1.0) / np.sqrt(n2, dtype=np.float32)).astype(np.float32)
    dirs = (d * inv[..., None]).astype(np.float32).reshape(-1, 3)
    origins = np.broadcast_to(eye.reshape(1, 3), dirs.shape).astype(np.float32).copy()
    return origins, dirs


def render(shape: Shape, cam: CtcCameraRays, width: int, height: int, max_steps: int = 100, epsilon: float = 1e-6,
           ctx: _lib.Context | None = None) -> np.ndarray:
    """[height, width, 4] f32: final position xyz and distance travelled (negative: no hit)."""
    ctx = ctx or _lib.default_context()
    out = np.empty((height, width, 4), dtype=np.float32)
    sh = shape._ctc_shape()
    ctx.check(_lib.lib().ctc_render(ctx.handle, C.byref(sh), C.byref(cam), width, height, max_steps, C.c_float(epsilon),
                                    out.ctypes.data))
    return out


def shade(gbuffer: np.ndarray, light=(0.5, 0.8, 0.6)) -> np.ndarray:
    """A Lambert image [height, width] u8 from the position buffer (screen-space normals): a viewer's job, here for demos."""
    pos, t = gbuffer[..., :3].astype(np.float64), gbuffer[..., 3]
    dx = np.zeros_like(pos); dy = np.zeros_like(pos)
    dx[:, 1:-1] = pos[:, 2:] - pos[:, :-2]; dy[1:-1] = pos[2:] - pos[:-2]
    n = np.cross(dx, dy); n /= np.maximum(np.linalg.norm(n, axis=-1, keepdims=True), 1e-30)
    l = np.asarray(light, dtype=np.float64); l /= np.linalg.norm(l)
    img = np.clip(np.abs(n @ l), 0.0, 1.0) * (t >= 0)
    return (255 * img).astype(np.uint8)
