"""Row-stripe partition of the source image across GPUs.

The match/compose path has no exchange step: block rows are independent, so rank r of `world`
takes a contiguous range of block rows, holds the whole (replicated) library, and writes its own
output slab; slabs are concatenated on the host.  The analysis build and tile preparation shard tiles the same way.
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


def gather_analysis(local, T: int, bytes_per_tile: int, world: int, rank: int, group=None):
    """Analysis cache build across GPUs: rank r analysed the tiles of `stripe_bounds(T, world, r)`;
    all-gather the per-tile results (`bytes_per_tile` = 3*N for colours, 3*ts*ts for tiles prepared by emo_resize) so every rank holds [T, bytes_per_tile].

    `local` is a flat uint8 torch tensor (CUDA with the NCCL backend, CPU with gloo) of this rank's
    (stop - start) * bytes_per_tile bytes.  Ranges differ by at most one tile, so shards are padded to the
    largest and trimmed after the collective — one all_gather, no other exchange."""
    import torch
    import torch.distributed as dist

    a, b = stripe_bounds(T, world, rank)
    if local.numel() != (b - a) * bytes_per_tile:
        raise ValueError(f"rank {rank}: expected {(b - a) * bytes_per_tile} bytes, got {local.numel()}")
    if world == 1:
        return local
    cap = (T + world - 1) // world * bytes_per_tile
    padded = torch.zeros(cap, dtype=torch.uint8, device=local.device)
    padded[:local.numel()] = local
    out = torch.empty(world * cap, dtype=torch.uint8, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    if T % world == 0:
        return out
    parts = []
    for r in range(world):
        ra, rb = stripe_bounds(T, world, r)
        parts.append(out[r * cap:r * cap + (rb - ra) * bytes_per_tile])
    return torch.cat(parts)
