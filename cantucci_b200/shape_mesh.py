"""Host-side mirror of the reference's `ShapeMesh` (the CALLER of the hot path).

Mirrors /root/reference/src/mesh/mod.rs:24-178 without the wgpu parts: the octree of MeshStatus,
`update()` = refine around the camera's focus points (:91-104), then start a mesh job for every empty
leaf (:129-161).  Where the reference hands one `generate_for_box` closure per leaf to a thread pool
(:141-148), this mirror hands ALL empty leaves of the frame to the GPU in one batched call -- the
batching point named in SURVEY 8b.  Jobs complete synchronously here, so a leaf goes from `None`
straight to `Ready` within the same `update()`.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _lib
from .mesh import MeshBuffer, Timings, generate_for_boxes
from .octree import Octree, spans_array, startup_tree
from .refine import FOCUS_POINTS, Camera, get_focii
from .shape import Shape

RESOLUTION = 64      # mesh/mod.rs:133


@dataclass
class Ready:
    """MeshStatus::Ready(view) (mesh/mod.rs:246-251); the view is the raw MeshBuffer here."""
    mesh: MeshBuffer


class ShapeMesh:
    def __init__(self, shape: Shape, ctx: _lib.Context | None = None, resolution: int = RESOLUTION):
        # ShapeMesh::new (mesh/mod.rs:45-78): startup tree of 64 leaves
        self.shape, self.ctx, self.resolution = shape, ctx, resolution
        self.tree: Octree = startup_tree(shape.bounding_box())
        self.batch_timings = Timings()
        self.finished_jobs = 0

    def get_focii(self, camera: Camera, focus_points: int = FOCUS_POINTS) -> np.ndarray:
        """mesh/mod.rs:205-243 (all rays sphere-traced in one kernel launch)."""
        return get_focii(self.shape, camera, focus_points, self.ctx)

    def update(self, camera: Camera) -> int:
        """mesh/mod.rs:82-178.  Returns the number of leaves meshed in this call."""
        # refine: split the Ready leaf around each focus point when the camera is near enough (:91-104)
        for focus in self.get_focii(camera):
            leaf = self.tree.leaf_around(focus)
            if leaf is not None and isinstance(leaf.data, Ready):
                dist = float(np.linalg.norm(camera.position.astype(np.float32) - focus.astype(np.float32)))
                threshold = 2.0 * abs(float(leaf.span.end[0]) - float(leaf.span.start[0]))
                if dist < threshold:
                    Octree.split(leaf)
        # one mesh job per empty leaf (:129-161) -- batched
        empty = [n for n in self.tree.leaves() if n.data is None]
        if not empty:
            return 0
        batch, timings = generate_for_boxes(spans_array([n.span for n in empty]), self.shape, self.resolution, self.ctx)
        for k, leaf in enumerate(empty):
            leaf.data = Ready(batch.mesh(k))
        self.finished_jobs += len(empty)
        self.batch_timings = self.batch_timings + timings
        return len(empty)

    def ready_meshes(self):
        """What ShapeMesh::draw walks (mesh/mod.rs:181-201): (span, MeshBuffer) of every Ready leaf."""
        return [(n.span, n.data.mesh) for n in self.tree.leaves() if isinstance(n.data, Ready)]
