"""emosaic_b200 — B200 (sm_100a) implementation of emosaic's data-parallel core.

The product is ``libemosaic_cuda.so`` (hand-written CUDA behind the C ABI of
``include/emosaic_cuda.h``); this package is the Python host-side mirror of the reference's
interface for that path (``analyse``, ``TileSet.build_kiddo``, ``render_nto1``, tint, the
``.emosaic_*`` cache) on top of that ABI.  No torch, no CPU fallback.
"""
from ._lib import EmosaicError, LIB_PATH, load  # noqa: F401
from .api import (  # noqa: F401
    Context, Group, stripe_bounds, host_register, host_unregister, Tile, TileSet, RenderResult, analyse, analyse_tiles, get_img_colors, flipped_coords,
    render_nto1, render_nto1_no_repeat, apply_tint, tint_alpha, adjust_source_dims, resize_source, prepare_view, prepare_tile, rotate,
)
from .cache import cache_file_name, serialize_tile_set, deserialize_tile_set  # noqa: F401

__all__ = [
    "EmosaicError", "Context", "Group", "stripe_bounds", "host_register", "host_unregister", "Tile", "TileSet", "RenderResult", "analyse", "analyse_tiles", "get_img_colors",
    "flipped_coords", "render_nto1", "render_nto1_no_repeat", "apply_tint", "tint_alpha", "adjust_source_dims", "resize_source", "prepare_view", "prepare_tile", "rotate", "cache_file_name",
    "serialize_tile_set", "deserialize_tile_set",
]
