"""Render statistics from the GPU's item / dist maps (reference: src/mosaic/stats.rs).

The match kernel already emits what `RenderStats::push_tile` collects per block (rendering.rs:211-214):
the chosen tile id and its distance.  These are cheap host-side reductions over those two maps.
"""
from __future__ import annotations

import sys
from typing import List, Optional, Sequence

import numpy as np


class RenderStats:
    """``RenderStats<D>`` (stats.rs:30-195) with the reference's interface: a map from the coordinates a tile was placed at
    to (tile id, distance); ``push_tile`` on the same coordinates replaces the entry, like ``HashMap::insert``.  The renderers
    fill it from the GPU's maps with ``from_maps`` instead of one ``push_tile`` per block under a mutex (rendering.rs:211-214)."""

    def __init__(self):
        self.tiles = {}          # (x, y) -> (signed 1-based tile id, distance); insertion-ordered

    @classmethod
    def from_maps(cls, item: np.ndarray, dist: np.ndarray, step: int) -> "RenderStats":
        """Blocks of an item / dist map as entries at (bx * step, by * step): step = dim for render_nto1 (source coordinates,
        rendering.rs:211-214), step = tile_size for render_nto1_no_repeat (output coordinates, :352-365).  Unplaced blocks
        (item 0) have no entry."""
        s = cls()
        bh, bw = item.shape
        for by in range(bh):
            for bx in range(bw):
                if item[by, bx] != 0:
                    s.tiles[(bx * step, by * step)] = (int(item[by, bx]), int(dist[by, bx]))
        return s

    def push_tile(self, x: int, y: int, tile_idx: int, distance: int):
        """stats.rs:56-64."""
        self.tiles[(int(x), int(y))] = (int(tile_idx), int(distance))

    def tile_count(self) -> int:
        return len(self.tiles)

    def summarise(self, paths: Optional[Sequence[str]] = None, file=sys.stderr) -> dict:
        """stats.rs:87-139."""
        if not self.tiles:
            print("No tiles recorded in statistics", file=file)
            return {}
        v = list(self.tiles.values())
        return summarise(np.array([t for t, _ in v], np.int64), np.array([d for _, d in v], np.uint64), paths, file)

    def render(self, tile_size: int) -> np.ndarray:
        """stats.rs:154-195: one grey pixel per (x / tile_size, y / tile_size), brightness (d / max_d * 255) as u8 in f64.  Entries
        that land on the same pixel overwrite each other in the order of the map (the reference: HashMap iteration order)."""
        if not self.tiles:
            raise ValueError("Cannot render visualization: no tiles recorded")
        if tile_size == 0:
            raise ValueError("Tile size must be greater than 0")
        max_x = max(x for x, _ in self.tiles)
        max_y = max(y for _, y in self.tiles)
        md = float(max(d for _, d in self.tiles.values()))
        img = np.zeros((max_y // tile_size + 1, max_x // tile_size + 1, 3), np.uint8)
        for (x, y), (_, d) in self.tiles.items():
            nd = d / md if md > 0.0 else 0.0
            img[y // tile_size, x // tile_size] = int(nd * 255.0)
        return img


def summarise(item: np.ndarray, dist: np.ndarray, paths: Optional[Sequence[str]] = None, file=sys.stderr, ctx=None) -> dict:
    """stats.rs:87-139: totals, unique images, average distance, top-10 usage, worst-10 matches.  With `ctx` (a Context whose
    library is the one the maps refer to) the counts and sums are reduced on the GPU (emo_stats); the two ordered top-10
    lists are host work either way.  item 0 (an unplaced no-repeat block) has no entry."""
    item = np.asarray(item)
    keep = item.reshape(-1) != 0
    if not keep.any():
        print("No tiles recorded in statistics", file=file)
        return {}
    ids = np.abs(item.reshape(-1)[keep])
    d = np.asarray(dist).reshape(-1)[keep].astype(np.uint64)
    if ctx is not None:
        sums, usage = ctx.stats(item, dist)
        uniq = np.nonzero(usage)[0] + 1
        counts = usage[uniq - 1].astype(np.int64)
        total, dsum = sums["placed"], sums["total_distance"]
    else:
        uniq, counts = np.unique(ids, return_counts=True)
        total, dsum = int(ids.size), int(d.sum())
    order = np.lexsort((uniq, -counts))[:10]
    worst = np.argsort(-d.astype(np.int64), kind="stable")[:10]
    name = (lambda i: paths[i - 1]) if paths is not None else (lambda i: f"tile #{i}")
    out = {
        "total": int(total), "unique": int(uniq.size), "average_distance": float(dsum) / total,
        "top": [(name(int(uniq[i])), int(counts[i])) for i in order],
        "worst": [(name(int(ids[i])), int(d[i])) for i in worst],
    }
    print("Mosaic Statistics:", file=file)
    print(f"  Total tiles placed: {out['total']}", file=file)
    print(f"  Unique images used: {out['unique']}", file=file)
    print(f"  Average color distance: {out['average_distance']:.3f}", file=file)
    print("\nTop 10 most used tiles:", file=file)
    for i, (p, c) in enumerate(out["top"]):
        print(f"  {i + 1}. {p} ({c} times)", file=file)
    print("\nWorst 10 color matches:", file=file)
    for i, (p, dd) in enumerate(out["worst"]):
        print(f"  {i + 1}. {p} (distance: {dd})", file=file)
    return out


def render(dist: np.ndarray, dim: int, tile_size: int, placed: Optional[np.ndarray] = None) -> np.ndarray:
    """stats.rs:154-195 over a whole dist map (vectorised ``RenderStats.from_maps(item, dist, dim).render(tile_size)``):
    grey-scale quality map, brightness = (d / max_d * 255) as u8 in f64.

    `dim` is the distance between the recorded coordinates of neighbouring blocks: render_nto1 records SOURCE coordinates
    (x = bx*dim, rendering.rs:211-214), so with tile_size > dim several blocks land on one pixel and the reference's
    survivor depends on HashMap iteration order — here the last block in row-major order wins (deterministic);
    render_nto1_no_repeat records OUTPUT coordinates (x = bx*tile_size, rendering.rs:352-365): pass dim = tile_size and the
    image has one pixel per block.  `placed` (bool map) masks blocks without a tile (no entry in the reference's map)."""
    if placed is None:
        placed = np.ones(dist.shape, bool)
    if dist.size == 0 or not placed.any():
        raise ValueError("Cannot render visualization: no tiles recorded")
    if tile_size == 0:
        raise ValueError("Tile size must be greater than 0")
    bh, bw = dist.shape
    ys_p, xs_p = np.nonzero(placed)
    max_x, max_y = int(xs_p.max()) * dim, int(ys_p.max()) * dim
    img = np.zeros((max_y // tile_size + 1, max_x // tile_size + 1, 3), np.uint8)
    md = float(dist[placed].max())
    norm = dist.astype(np.float64) / md if md > 0 else np.zeros(dist.shape)
    b = (norm * 255.0).astype(np.uint8)
    img[(ys_p * dim) // tile_size, (xs_p * dim) // tile_size] = b[ys_p, xs_p][:, None]
    return img
