"""Host-side mirror of the reference's octree span maths.

Spans are the unit of work handed to the GPU; the tree itself is host logic
(SURVEY.md section 8 a14).  Mirrors /root/reference/src/octree/mod.rs:13-32,
296-329 and the DFS order of src/octree/iter.rs:88-106.  All arithmetic is
IEEE f32 via numpy, in the reference's evaluation order.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Iterator, Optional

import numpy as np

f32 = np.float32


@dataclass(frozen=True)
class Span:
    """octree::Span = Range<Point3<f32>> (octree/mod.rs:13)."""
    start: tuple
    end: tuple

    def __post_init__(self):
        object.__setattr__(self, "start", tuple(f32(v) for v in self.start))
        object.__setattr__(self, "end", tuple(f32(v) for v in self.end))

    def center(self) -> tuple:
        # SpanExt::center (octree/mod.rs:21-23): start + (end - start) / 2.0
        return tuple(f32(s + f32(f32(e - s) / f32(2.0))) for s, e in zip(self.start, self.end))

    def contains(self, p) -> bool:
        # SpanExt::contains (octree/mod.rs:25-31): half-open box
        return all(s <= f32(q) < e for s, q, e in zip(self.start, p, self.end))

    def as_row(self) -> np.ndarray:
        return np.array([*self.start, *self.end], dtype=np.float32)


def create_spans(parent: Span) -> list[Span]:
    """create_spans (octree/mod.rs:315-329): child i = (x,y,z) bits, z lowest."""
    c = parent.center()
    out = []
    for i in range(8):
        hi = [(i >> (2 - k)) & 1 for k in range(3)]
        start = tuple(c[k] if hi[k] else parent.start[k] for k in range(3))
        end = tuple(parent.end[k] if hi[k] else c[k] for k in range(3))
        out.append(Span(start, end))
    return out


@dataclass
class _Node:
    span: Span
    depth: int = 0
    data: Any = None                      # Leaf(Option<L>)
    children: Optional[list] = None       # SubTree

    @property
    def is_leaf(self) -> bool:
        return self.children is None


class Octree:
    """Octree<L, I> (octree/mod.rs:39-100) restricted to what the span scheduler needs."""

    def __init__(self, span: Span):
        self.root = _Node(span)           # Octree::spanning

    @staticmethod
    def split(node: _Node):
        """NodeEntryMut::split (octree/mod.rs:296-310); returns the old leaf data."""
        assert node.is_leaf
        old, node.data = node.data, None
        node.children = [_Node(s, node.depth + 1) for s in create_spans(node.span)]
        return old

    def iter(self) -> Iterator[_Node]:
        """IterMut (octree/iter.rs:88-106): LIFO stack, so children come out 7 -> 0."""
        stack = [self.root]
        while stack:
            n = stack.pop()
            if not n.is_leaf:
                stack.extend(n.children)
            yield n

    def leaves(self) -> list[_Node]:
        return [n for n in self.iter() if n.is_leaf]

    def leaf_around(self, p) -> Optional[_Node]:
        """Octree::leaf_around_mut (octree/mod.rs:75-89)."""
        node = self.root
        if not node.span.contains(p):
            return None
        while not node.is_leaf:
            node = next(c for c in node.children if c.span.contains(p))
        return node


def startup_tree(bounding_box: Span) -> Octree:
    """ShapeMesh::new (mesh/mod.rs:52-56): split the root and its 8 children -> 64 leaves."""
    tree = Octree(bounding_box)
    Octree.split(tree.root)
    for child in tree.root.children:
        Octree.split(child)
    return tree


def spans_array(spans) -> np.ndarray:
    """Iterable[Span] | ndarray -> contiguous float32 [n, 6] (ctc_span layout)."""
    if isinstance(spans, np.ndarray):
        return np.ascontiguousarray(spans, dtype=np.float32).reshape(-1, 6)
    rows = [s.as_row() if isinstance(s, Span) else np.asarray(s, dtype=np.float32).reshape(6) for s in spans]
    if not rows:
        return np.zeros((0, 6), dtype=np.float32)
    return np.ascontiguousarray(np.stack(rows), dtype=np.float32)


def tile_volume(bounding_box: Span, tiles_per_axis: int) -> np.ndarray:
    """A dense volume as tiles^3 equal spans (config 5: 4096^3 = 64^3 spans of R=64).

    Edges are start + (end-start) * (i / tiles) in f32, so neighbouring spans
    share their faces exactly (tiles_per_axis is a power of two for the
    BASELINE configs, making i/tiles exact)."""
    t = tiles_per_axis
    axes = []
    for k in range(3):
        s, e = bounding_box.start[k], bounding_box.end[k]
        axes.append((s + (e - s) * (np.arange(t + 1, dtype=np.float32) / f32(t))).astype(np.float32))
    ix, iy, iz = np.meshgrid(np.arange(t), np.arange(t), np.arange(t), indexing="ij")
    ix, iy, iz = ix.ravel(), iy.ravel(), iz.ravel()
    out = np.empty((t ** 3, 6), dtype=np.float32)
    out[:, 0], out[:, 1], out[:, 2] = axes[0][ix], axes[1][iy], axes[2][iz]
    out[:, 3], out[:, 4], out[:, 5] = axes[0][ix + 1], axes[1][iy + 1], axes[2][iz + 1]
    return out
