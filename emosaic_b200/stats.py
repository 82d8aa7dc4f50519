"""Render statistics from the GPU's item / dist maps (reference: src/mosaic/stats.rs).

The match kernel already emits what `RenderStats::push_tile` collects per block (rendering.rs:211-214):
the chosen tile id and its distance.  These are cheap host-side reductions over those two maps.
"""
from __future__ import annotations

import sys
from typing import List, Optional, Sequence

import numpy as np


def summarise(item: np.ndarray, dist: np.ndarray, paths: Optional[Sequence[str]] = None, file=sys.stderr) -> dict:
    """stats.rs:87-139: totals, unique images, average distance, top-10 usage, worst-10 matches."""
    if item.size == 0:
        print("No tiles recorded in statistics", file=file)
        return {}
    ids = np.abs(item).reshape(-1)
    d = dist.reshape(-1).astype(np.uint64)
    uniq, counts = np.unique(ids, return_counts=True)
    order = np.lexsort((uniq, -counts))[:10]
    worst = np.argsort(-d.astype(np.int64), kind="stable")[:10]
    name = (lambda i: paths[i - 1]) if paths is not None else (lambda i: f"tile #{i}")
    out = {
        "total": int(ids.size), "unique": int(uniq.size), "average_distance": float(d.sum()) / ids.size,
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


def render(dist: np.ndarray, dim: int, tile_size: int) -> np.ndarray:
    """stats.rs:154-195: grey-scale quality map, brightness = (d / max_d * 255) as u8 in f64.

    The reference keys its map by SOURCE coordinates (x = bx*dim, y = by*dim, rendering.rs:211-214) and then
    divides by tile_size, so with tile_size > dim several blocks land on one pixel and the survivor depends on
    HashMap iteration order.  Here the last block in row-major order wins (deterministic)."""
    if dist.size == 0:
        raise ValueError("Cannot render visualization: no tiles recorded")
    if tile_size == 0:
        raise ValueError("Tile size must be greater than 0")
    bh, bw = dist.shape
    max_x, max_y = (bw - 1) * dim, (bh - 1) * dim
    img = np.zeros((max_y // tile_size + 1, max_x // tile_size + 1, 3), np.uint8)
    md = float(dist.max())
    norm = dist.astype(np.float64) / md if md > 0 else np.zeros(dist.shape)
    b = (norm * 255.0).astype(np.uint8)
    ys = (np.arange(bh) * dim) // tile_size
    xs = (np.arange(bw) * dim) // tile_size
    img[ys[:, None], xs[None, :]] = b[:, :, None]
    return img
