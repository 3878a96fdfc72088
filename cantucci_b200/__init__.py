"""cantucci_b200 -- B200-native (sm_100a) implementation of cantucci's hot path.

Mandelbulb distance-estimator sampling over octree-leaf span grids and
naive-surface-nets mesh extraction, as hand-written CUDA kernels behind a C ABI
(include/cantucci_b200.h).  This package is the host-side mirror of the
reference's `Shape` / `MeshBuffer` / `octree::Span` interface over that ABI.
"""
from ._lib import (CantucciError, Context, MultiContext, VERTEX_DTYPE, default_context, lib, shard_plan, LIB_PATH)
from .mesh import (MeshBatch, MeshBuffer, MultiMeshBatch, Timings, cull_spans, generate_for_boxes,
                   generate_for_boxes_multi, order_spans, sample_grids, sample_signs)
from .octree import Octree, Span, create_spans, spans_array, startup_tree, tile_volume
from .shape import Mandelbulb, Shape, Sphere
from .render import CtcCameraRays, look_at_rays, pixel_rays, render, shade
from .shape_mesh import ShapeMesh
from .view import InteropBuffer, MeshView, MeshViews, generate_views

__all__ = [
    "CantucciError", "Context", "MultiContext", "VERTEX_DTYPE", "default_context", "lib", "shard_plan", "LIB_PATH",
    "MeshBatch", "MeshBuffer", "MultiMeshBatch", "Timings", "cull_spans", "generate_for_boxes", "generate_for_boxes_multi",
    "order_spans", "sample_grids", "sample_signs",
    "Octree", "Span", "create_spans", "spans_array", "startup_tree", "tile_volume",
    "Mandelbulb", "Shape", "Sphere", "ShapeMesh",
    "CtcCameraRays", "look_at_rays", "pixel_rays", "render", "shade",
    "InteropBuffer", "MeshView", "MeshViews", "generate_views",
]
