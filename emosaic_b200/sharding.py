"""Row-stripe partition of the source image across GPUs.

The match/compose path has no exchange step: block rows are independent, so rank r of `world`
takes a contiguous range of block rows, holds the whole (replicated) library, and writes its own
output slab; slabs are concatenated on the host.  The analysis build shards tiles the same way.
"""
from __future__ import annotations

from typing import List, Tuple


def stripe_bounds(n_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [start, stop) of `n_rows` units for `rank`; sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} / world {world}")
    base, extra = divmod(n_rows, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def all_stripes(n_rows: int, world: int) -> List[Tuple[int, int]]:
    return [stripe_bounds(n_rows, world, r) for r in range(world)]


def source_stripe(src, dim: int, world: int, rank: int):
    """Rows of `src` ([H,W,3]) owned by `rank`: whole block rows (dim source rows each)."""
    bh = src.shape[0] // dim
    a, b = stripe_bounds(bh, world, rank)
    return src[a * dim:b * dim], (a, b)
