"""Host-side mirror of the reference's camera-driven refinement, used to build BASELINE config 3
("octree refined to depth 6 around the camera, every leaf meshed at 64^3").

Mirrors /root/reference/src/mesh/mod.rs:86-104 (split rule), :205-243 (get_focii: FOCUS_POINTS^2 rays
through the near plane, sphere-traced with the shape's DE), src/camera.rs:90-103,186-189 (near plane) and
the default orbit camera of src/control/orbit.rs:49-55 / src/app.rs:98.  The DE calls go through the CUDA
batch entry point (Shape::batch_min_distance_from); the tree logic is host code, as in the reference.
View-space algebra is done in float64 (the reference inverts a 4x4 f32 matrix); the camera here is
synthetic, so only which leaves get split matters.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .octree import Octree, Span, startup_tree
from .shape import Shape

FOCUS_POINTS = 5          # mesh/mod.rs:86
EPSILON = 0.000_001       # mesh/mod.rs:206
MAX_ITERS = 100           # mesh/mod.rs:207


@dataclass
class Camera:
    """camera.rs:15-19 + Projection (camera.rs:114-127)."""
    position: np.ndarray
    direction: np.ndarray
    fov: float = 1.0            # app.rs:98: Rad(1.0)
    aspect_ratio: float = 1.0
    near_plane: float = 0.000_04

    @classmethod
    def default_orbit(cls) -> "Camera":
        """Orbit::around(origin) (orbit.rs:49-55): distance 3, looking +x."""
        return cls(np.array([-3.0, 0.0, 0.0]), np.array([1.0, 0.0, 0.0]))

    def near_plane_dimension(self):
        h = 2.0 * self.near_plane * np.tan(self.fov * 0.5)      # camera.rs:186-189
        return h * self.aspect_ratio, h

    def basis(self):
        f = self.direction / np.linalg.norm(self.direction)
        s = np.cross(f, np.array([0.0, 0.0, 1.0])); s /= np.linalg.norm(s)     # look_at_rh, up = +z
        u = np.cross(s, f)
        return s, u, f


def get_focii(shape: Shape, camera: Camera, focus_points: int = FOCUS_POINTS, ctx=None) -> np.ndarray:
    """ShapeMesh::get_focii (mesh/mod.rs:205-243): the surface hits of focus_points^2 rays."""
    w, h = camera.near_plane_dimension()
    s, u, f = camera.basis()
    size_h, size_v = w / focus_points, h / focus_points
    pts = []
    for x in range(focus_points):            # iter::square: x outer, y inner (util/iter.rs:54-88)
        for y in range(focus_points):
            cx = -w / 2 + x * size_h + w / (2.0 * focus_points)
            cy = -h / 2 + y * size_v + h / (2.0 * focus_points)
            pts.append(camera.position + s * cx + u * cy + f * camera.near_plane)
    pts = np.array(pts)
    dirs = pts - camera.position
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    pos = np.repeat(camera.position[None, :], len(pts), axis=0).astype(np.float32)
    dirs = dirs.astype(np.float32)
    if type(shape).batch_min_distance_from is Shape.batch_min_distance_from:
        # the CUDA shape: all rays are sphere-traced in one kernel launch (ctc_ray_march)
        import ctypes as C
        from . import _lib
        ctx = ctx or _lib.default_context()
        pos = np.ascontiguousarray(pos); dirs = np.ascontiguousarray(dirs)
        out = np.empty_like(pos); hit = np.zeros(len(pts), dtype=np.uint32)
        sh = shape._ctc_shape()
        ctx.check(_lib.lib().ctc_ray_march(ctx.handle, C.byref(sh), pos.ctypes.data, dirs.ctypes.data, len(pts),
                                           MAX_ITERS, EPSILON, out.ctypes.data, hit.ctypes.data))
        return out[hit != 0]
    done = np.zeros(len(pts), dtype=bool)
    for _ in range(MAX_ITERS):
        live = ~done
        if not live.any():
            break
        d = shape.batch_min_distance_from(pos[live], ctx)
        pos[live] = (pos[live] + dirs[live] * d[:, None]).astype(np.float32)      # pos += dir * distance
        idx = np.nonzero(live)[0]
        done[idx[d < EPSILON]] = True
    return pos[done]


def refine_step(tree: Octree, camera: Camera, focii: np.ndarray, max_depth: int) -> int:
    """The split rule of ShapeMesh::update (mesh/mod.rs:91-104), treating every leaf as Ready."""
    splits = 0
    for focus in focii:
        leaf = tree.leaf_around(focus)
        if leaf is None or leaf.depth >= max_depth:
            continue
        dist = float(np.linalg.norm(camera.position.astype(np.float32) - focus.astype(np.float32)))
        threshold = 2.0 * abs(float(leaf.span.end[0]) - float(leaf.span.start[0]))
        if dist < threshold:
            Octree.split(leaf)
            splits += 1
    return splits


def config3_spans(shape: Shape, max_depth: int = 6, ctx=None) -> tuple[np.ndarray, Camera]:
    """BASELINE config 3: start from the 64-leaf startup tree, put the camera on the default orbit
    ray at its surface hit (backed off by one near plane), and apply the reference's split rule until
    nothing within `max_depth` splits any more.  Returns every leaf span in iter order."""
    tree = startup_tree(shape.bounding_box())
    cam = Camera.default_orbit()
    hit = get_focii(shape, cam, 1, ctx)
    assert len(hit) == 1, "the central ray must hit the shape"
    cam = Camera(hit[0].astype(np.float64) - cam.direction * 1e-3, cam.direction)
    for _ in range(4 * max_depth):
        focii = get_focii(shape, cam, FOCUS_POINTS, ctx)
        if refine_step(tree, cam, focii, max_depth) == 0:
            break
    from .octree import spans_array
    return spans_array([n.span for n in tree.leaves()]), cam
