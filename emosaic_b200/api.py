"""Host-side mirror of the reference's interface for the hot path, on top of the C ABI.

Names, argument meaning and error behaviour follow the reference (paths under the reference
root): ``analyse`` (src/mosaic/analysis.rs:5), ``get_img_colors`` (analysis.rs:23), ``Tile`` /
``Tile.coords`` (tiles/tile.rs), ``flipped_coords`` (tiles/utils.rs:18), ``TileSet`` /
``build_kiddo`` / ``get_tile`` / ``get_image`` (tiles/tileset.rs), ``render_nto1``
(rendering.rs:124) and the tint block (main.rs:447-478).  Where the reference panics or exits,
these raise ``EmosaicError`` (status code + the reference's message).

All compute goes through ``libemosaic_cuda.so``; numpy is only used to hold host buffers.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import EmosaicError, check

EMO_ERR_ARG = -1
EMO_ERR_UNSUPPORTED = -5


def _u8(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _isqrt_exact(N: int) -> int:
    d = math.isqrt(N)
    if d * d != N or N < 1:
        raise EmosaicError(EMO_ERR_ARG, f"N={N} is not a square number of cells")
    return d


MATCH_AUTO, MATCH_SCAN, MATCH_INDEX, MATCH_INDEX_WIDE, MATCH_INDEX_COMPACT = 0, 1, 2, 3, 4
MATCH_MODES = {"auto": MATCH_AUTO, "scan": MATCH_SCAN, "index": MATCH_INDEX, "index_wide": MATCH_INDEX_WIDE,
               "index_compact": MATCH_INDEX_COMPACT}


class Context:
    """One per GPU (one process per GPU under torchrun).  Wraps ``emo_ctx``."""

    def __init__(self, device: int = 0, _handle=None):
        self._lib = _lib.load()
        self._owned = _handle is None
        if _handle is None:
            h = C.c_void_p()
            check(self._lib.emo_create(int(device), C.byref(h)))
        else:  # a member of a Group: the group owns the emo_ctx
            h = C.c_void_p(_handle)
        self._h = h
        self.device = device
        self.N = 0
        self.dim = 0
        self.ts = 0
        self.T = 0

    def close(self):
        if getattr(self, "_h", None):
            if self._owned:
                self._lib.emo_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- plumbing -------------------------------------------------------------------------
    def set_stream(self, cuda_stream: int | None):
        check(self._lib.emo_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def sync(self):
        check(self._lib.emo_sync(self._h))

    def device_info(self):
        name = C.create_string_buffer(256)
        sm, ma, mi = C.c_int(), C.c_int(), C.c_int()
        check(self._lib.emo_device_info(self._h, name, 256, C.byref(sm), C.byref(ma), C.byref(mi)))
        return {"name": name.value.decode(), "sm_count": sm.value, "cc": (ma.value, mi.value)}

    def launch_count(self) -> int:
        return int(self._lib.emo_launch_count(self._h))

    def timer_start(self):
        check(self._lib.emo_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        check(self._lib.emo_timer_stop(self._h, C.byref(ms)))
        return float(ms.value)

    def mark(self, slot: int):
        check(self._lib.emo_mark(self._h, slot))

    def mark_elapsed(self, a: int, b: int) -> float:
        ms = C.c_float()
        check(self._lib.emo_mark_elapsed(self._h, a, b, C.byref(ms)))
        return float(ms.value)

    def dev_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        check(self._lib.emo_dev_alloc(self._h, nbytes, C.byref(p)))
        return int(p.value)

    def dev_free(self, p: int):
        check(self._lib.emo_dev_free(self._h, C.c_void_p(p)))

    def host_alloc(self, nbytes: int) -> np.ndarray:
        """Pinned host buffer as a uint8 numpy array (freed with host_free)."""
        p = C.c_void_p()
        check(self._lib.emo_host_alloc(self._h, nbytes, C.byref(p)))
        buf = (C.c_uint8 * nbytes).from_address(p.value)
        arr = np.frombuffer(buf, dtype=np.uint8)
        arr.flags.writeable = True
        return arr

    def host_free(self, arr: np.ndarray):
        check(self._lib.emo_host_free(self._h, C.c_void_p(arr.ctypes.data)))

    def h2d(self, dst: int, src: np.ndarray):
        check(self._lib.emo_copy_h2d(self._h, C.c_void_p(dst), _ptr(src), src.nbytes))

    def d2h(self, dst: np.ndarray, src: int):
        check(self._lib.emo_copy_d2h(self._h, _ptr(dst), C.c_void_p(src), dst.nbytes))

    # -- multi-GPU, one process per GPU (include/emosaic_cuda.h "multi-GPU" (a)) -----------
    @staticmethod
    def comm_unique_id() -> bytes:
        """ncclGetUniqueId: 128 bytes rank 0 hands to every rank (torchrun's store, a file, ...)."""
        buf = C.create_string_buffer(128)
        check(_lib.load().emo_comm_unique_id(buf))
        return buf.raw

    def comm_init_rank(self, uid: bytes, rank: int, world: int):
        if len(uid) != 128:
            raise EmosaicError(EMO_ERR_ARG, f"the communicator id has 128 bytes, got {len(uid)}")
        check(self._lib.emo_comm_init_rank(self._h, C.c_char_p(uid), rank, world))
        self.rank, self.world = rank, world

    def comm_info(self):
        r, w, v = C.c_int(), C.c_int(), C.c_int()
        check(self._lib.emo_comm_info(self._h, C.byref(r), C.byref(w), C.byref(v)))
        return {"rank": r.value, "world": w.value, "nccl_version": v.value}

    def comm_set_library(self, colors=None, tile_px=None, root: int = 0):
        """emo_set_library on every rank from the root's arrays (NCCL broadcast); the other ranks pass nothing."""
        T = N = ts = 0
        if colors is not None:
            colors = _u8(colors)
            if colors.ndim != 3 or colors.shape[2] != 3:
                raise EmosaicError(EMO_ERR_ARG, f"colors must be [T,N,3], got {colors.shape}")
            T, N = colors.shape[:2]
            if tile_px is not None:
                tile_px = _u8(tile_px)
                ts = tile_px.shape[1]
        check(self._lib.emo_comm_set_library(self._h, _ptr(colors), _ptr(tile_px), T, N, ts, root))
        self._refresh_library_info()

    def comm_set_library_dev(self, colors_dev: int, tile_px_dev: int, T: int, N: int, ts: int, root: int = 0):
        check(self._lib.emo_comm_set_library_dev(self._h, C.c_void_p(colors_dev or 0), C.c_void_p(tile_px_dev or 0), T, N, ts, root))
        self._refresh_library_info()

    def library_info(self):
        """(T, N, ts) of the resident library (zeros when none)."""
        T, N, ts = C.c_uint32(), C.c_uint32(), C.c_uint32()
        check(self._lib.emo_library_info(self._h, C.byref(T), C.byref(N), C.byref(ts)))
        return T.value, N.value, ts.value

    def _refresh_library_info(self):
        self.T, self.N, self.ts = self.library_info()
        self.dim = _isqrt_exact(self.N) if self.N else 0

    def comm_broadcast_dev(self, buf_dev: int, nbytes: int, root: int = 0):
        check(self._lib.emo_comm_broadcast_dev(self._h, C.c_void_p(buf_dev), nbytes, root))

    def comm_allgather_analysis_dev(self, local_dev: int, T: int, bytes_per_tile: int, all_dev: int):
        check(self._lib.emo_comm_allgather_analysis_dev(self._h, C.c_void_p(local_dev or 0), T, bytes_per_tile, C.c_void_p(all_dev)))

    # -- (0) Lanczos3 resize --------------------------------------------------------------
    def resize(self, images, nw: int, nh: int, view=None) -> np.ndarray:
        """image 0.25.2 ``imageops::resize(view, nw, nh, Lanczos3)`` (main.rs:595, tiles/utils.rs:188-189).
        images [H,W,3] or a batch [n,H,W,3]; view = (x0, y0, cw, ch) applied to every image, default the whole image."""
        images = _u8(images)
        single = images.ndim == 3
        if single:
            images = images[None]
        if images.ndim != 4 or images.shape[3] != 3:
            raise EmosaicError(EMO_ERR_ARG, f"images must be [H,W,3] or [n,H,W,3], got {images.shape}")
        n, H, W = images.shape[:3]
        x0, y0, cw, ch = view if view is not None else (0, 0, W, H)
        out = np.zeros((n, nh, nw, 3), np.uint8)
        check(self._lib.emo_resize(self._h, _ptr(images), n, W, H, x0, y0, cw, ch, nw, nh, _ptr(out)))
        return out[0] if single else out

    def resize_dev(self, images_dev: int, n: int, W: int, H: int, view, nw: int, nh: int, out_dev: int):
        x0, y0, cw, ch = view if view is not None else (0, 0, W, H)
        check(self._lib.emo_resize_dev(self._h, C.c_void_p(images_dev), n, W, H, x0, y0, cw, ch, nw, nh, C.c_void_p(out_dev)))

    # -- (1) analysis ---------------------------------------------------------------------
    def analyse_tiles(self, tiles, dim: int) -> np.ndarray:
        """tiles [T,ts,ts,3] -> [T,dim*dim,3]  (analysis.rs:5-20 over the library, main.rs:786-794)."""
        tiles = _u8(tiles)
        if tiles.ndim != 4 or tiles.shape[1] != tiles.shape[2] or tiles.shape[3] != 3:
            raise EmosaicError(EMO_ERR_ARG, f"tiles must be [T,ts,ts,3], got {tiles.shape}")
        T, ts = tiles.shape[0], tiles.shape[1]
        out = np.zeros((T, dim * dim, 3), np.uint8)
        check(self._lib.emo_analyse(self._h, _ptr(tiles), T, ts, dim, _ptr(out)))
        return out

    def analyse_tiles_fused(self, tiles):
        """One pass -> (1to1 [T,1,3], 4to1 [T,4,3])."""
        tiles = _u8(tiles)
        T, ts = tiles.shape[0], tiles.shape[1]
        o1 = np.zeros((T, 1, 3), np.uint8)
        o4 = np.zeros((T, 4, 3), np.uint8)
        check(self._lib.emo_analyse_fused(self._h, _ptr(tiles), T, ts, _ptr(o1), _ptr(o4)))
        return o1, o4

    def analyse_dev(self, tiles_dev: int, T: int, ts: int, dim: int, out_dev: int):
        check(self._lib.emo_analyse_dev(self._h, C.c_void_p(tiles_dev), T, ts, dim, C.c_void_p(out_dev)))

    def analyse_fused_dev(self, tiles_dev: int, T: int, ts: int, out1_dev: int, out4_dev: int):
        check(self._lib.emo_analyse_fused_dev(self._h, C.c_void_p(tiles_dev), T, ts, C.c_void_p(out1_dev),
                                              C.c_void_p(out4_dev)))

    # -- (2) library ----------------------------------------------------------------------
    def set_library(self, colors, tile_px=None):
        """colors [T,N,3], tile_px [T,ts,ts,3] or None  (tileset.rs:178-190 build_kiddo)."""
        colors = _u8(colors)
        if colors.ndim != 3 or colors.shape[2] != 3:
            raise EmosaicError(EMO_ERR_ARG, f"colors must be [T,N,3], got {colors.shape}")
        T, N = colors.shape[0], colors.shape[1]
        ts = 0
        if tile_px is not None:
            tile_px = _u8(tile_px)
            if tile_px.ndim != 4 or tile_px.shape[0] != T or tile_px.shape[1] != tile_px.shape[2] or tile_px.shape[3] != 3:
                raise EmosaicError(EMO_ERR_ARG, f"tile_px must be [T,ts,ts,3] with T={T}, got {tile_px.shape}")
            ts = tile_px.shape[1]
        check(self._lib.emo_set_library(self._h, _ptr(colors), _ptr(tile_px), T, N, ts))
        self.T, self.N, self.dim, self.ts = T, N, _isqrt_exact(N), ts

    def set_library_dev(self, colors_dev: int, tile_px_dev: int, T: int, N: int, ts: int):
        check(self._lib.emo_set_library_dev(self._h, C.c_void_p(colors_dev), C.c_void_p(tile_px_dev or 0), T, N, ts))
        self.T, self.N, self.dim, self.ts = T, N, _isqrt_exact(N), ts

    # -- (2b) search index ------------------------------------------------------------------
    def build_index(self):
        """1to1 only: exact nearest-tile table over the colour cube (the KD-tree's job, tileset.rs:178-190)."""
        check(self._lib.emo_build_index(self._h))

    def set_match_mode(self, mode: int | str):
        """'auto' (default) | 'scan' (brute force) | 'index' (colour-cube table when the library supports one) |
        'index_wide' / 'index_compact' (the index with its 64 MiB / 32 MiB form forced)."""
        if isinstance(mode, str):
            mode = MATCH_MODES[mode]
        check(self._lib.emo_set_match_mode(self._h, int(mode)))

    # -- (3) match ------------------------------------------------------------------------
    def match(self, src):
        """src [H,W,3] -> (item [H/dim,W/dim] int32 signed 1-based, dist uint32)  (rendering.rs:158-221)."""
        src = _u8(src)
        if src.ndim != 3 or src.shape[2] != 3:
            raise EmosaicError(EMO_ERR_ARG, f"src must be [H,W,3], got {src.shape}")
        H, W = src.shape[:2]
        d = max(self.dim, 1)
        item = np.zeros((H // d, W // d), np.int32)
        dist = np.zeros((H // d, W // d), np.uint32)
        check(self._lib.emo_match(self._h, _ptr(src), W, H, _ptr(item), _ptr(dist)))
        return item, dist

    def topk(self, src, first: int, k: int, exclude=None):
        """Candidates [first, first+k) of every block's list sorted by (distance, insertion rank): (item [Q,k] int32, dist
        [Q,k] uint32), blocks row-major; past the end item 0 / dist 0xFFFFFFFF  (rendering.rs:307-321 nearest_n).
        exclude: optional [T] uint8, 1 = tile retired (left out of every list, rendering.rs:366-380)."""
        src = _u8(src)
        if src.ndim != 3 or src.shape[2] != 3:
            raise EmosaicError(EMO_ERR_ARG, f"src must be [H,W,3], got {src.shape}")
        H, W = src.shape[:2]
        d = max(self.dim, 1)
        Q = (H // d) * (W // d)
        item = np.zeros((Q, k), np.int32)
        dist = np.zeros((Q, k), np.uint32)
        if exclude is not None:
            exclude = _u8(exclude).reshape(-1)
            if exclude.size != self.T:
                raise EmosaicError(EMO_ERR_ARG, f"exclude must have one byte per tile ({self.T}), got {exclude.size}")
        check(self._lib.emo_topk(self._h, _ptr(src), W, H, first, k, _ptr(exclude), _ptr(item), _ptr(dist)))
        return item, dist

    def topk_dev(self, src_dev: int, W: int, H: int, first: int, k: int, item_dev: int, dist_dev: int, exclude_dev: int = 0):
        check(self._lib.emo_topk_dev(self._h, C.c_void_p(src_dev), W, H, first, k, C.c_void_p(exclude_dev or 0), C.c_void_p(item_dev),
                                     C.c_void_p(dist_dev)))

    def no_repeat(self, src, page: int = 0):
        """emo_no_repeat: the assignment of render_nto1_no_repeat (rendering.rs:262-392): (item [bh,bw] with 0 = unplaced,
        dist [bh,bw], counters dict).  page = candidates per block of the first page (0: automatic)."""
        src = _u8(src)
        if src.ndim != 3 or src.shape[2] != 3:
            raise EmosaicError(EMO_ERR_ARG, f"src must be [H,W,3], got {src.shape}")
        H, W = src.shape[:2]
        d = max(self.dim, 1)
        item = np.zeros((H // d, W // d), np.int32)
        dist = np.zeros((H // d, W // d), np.uint32)
        cnt = (C.c_uint64 * 4)()
        check(self._lib.emo_no_repeat(self._h, _ptr(src), W, H, page, _ptr(item), _ptr(dist), cnt))
        return item, dist, {"refill_launches": int(cnt[0]), "blocks_refilled": int(cnt[1]), "heap_pops": int(cnt[2]), "placed": int(cnt[3])}

    def match_dev(self, src_dev: int, W: int, H: int, item_dev: int, dist_dev: int):
        check(self._lib.emo_match_dev(self._h, C.c_void_p(src_dev), W, H, C.c_void_p(item_dev), C.c_void_p(dist_dev)))

    # -- (4) compose (+tint) ---------------------------------------------------------------
    def compose(self, item, src=None, out_channels: int = 3, tint_alpha: int = 0, out: np.ndarray | None = None):
        """item [bh,bw] (+ src [bh*dim,bw*dim,3] when tinting) -> [bh*ts,bw*ts,out_channels]."""
        item = np.ascontiguousarray(item, dtype=np.int32)
        bh, bw = item.shape
        W, H = bw * self.dim, bh * self.dim
        if src is not None:
            src = _u8(src)
            if src.shape != (H, W, 3):
                raise EmosaicError(EMO_ERR_ARG, f"src must be {(H, W, 3)}, got {src.shape}")
        if out is None:
            out = np.zeros((bh * self.ts, bw * self.ts, out_channels), np.uint8)
        check(self._lib.emo_compose(self._h, _ptr(item), _ptr(src), W, H, out_channels, tint_alpha, _ptr(out)))
        return out

    def compose_overlay(self, item, overlay, tint_alpha: int):
        """Tint with an overlay image of any size (the original image of main.rs:447-478): RGBA result."""
        item = np.ascontiguousarray(item, dtype=np.int32)
        overlay = _u8(overlay)
        bh, bw = item.shape
        out = np.zeros((bh * self.ts, bw * self.ts, 4), np.uint8)
        check(self._lib.emo_compose_overlay(self._h, _ptr(item), bw * self.dim, bh * self.dim, _ptr(overlay), overlay.shape[1],
                                            overlay.shape[0], tint_alpha, _ptr(out)))
        return out

    def compose_dev(self, item_dev: int, src_dev: int, W: int, H: int, out_channels: int, tint_alpha: int, out_dev: int):
        check(self._lib.emo_compose_dev(self._h, C.c_void_p(item_dev), C.c_void_p(src_dev or 0), W, H, out_channels,
                                        tint_alpha, C.c_void_p(out_dev)))

    def reserve(self, W: int, H: int, out_channels: int = 3):
        """emo_reserve: size the staging buffers of the host-pointer calls for sources up to W x H ahead of time."""
        check(self._lib.emo_reserve(self._h, W, H, out_channels))

    def stats(self, item, dist, want_usage: bool = True):
        """emo_stats: the reductions of RenderStats (stats.rs:87-139, :169-175) over a render's maps, on the GPU:
        {"placed", "total_distance", "max_distance"} and usage [T] (blocks per tile, either orientation)."""
        item = np.ascontiguousarray(item, dtype=np.int32)
        dist = np.ascontiguousarray(dist, dtype=np.uint32)
        sums = np.zeros(3, np.uint64)
        usage = np.zeros(self.T, np.uint32) if want_usage else None
        check(self._lib.emo_stats(self._h, _ptr(item), _ptr(dist), item.size, self.T, _ptr(sums), _ptr(usage)))
        return {"placed": int(sums[0]), "total_distance": int(sums[1]), "max_distance": int(sums[2])}, usage

    def stats_dev(self, item_dev: int, dist_dev: int, Q: int, sums_dev: int, usage_dev: int = 0):
        check(self._lib.emo_stats_dev(self._h, C.c_void_p(item_dev), C.c_void_p(dist_dev), Q, self.T, C.c_void_p(sums_dev),
                                      C.c_void_p(usage_dev or 0)))

    def mosaic_dev(self, src_dev: int, W: int, H: int, out_channels: int, tint_alpha: int, item_dev: int, dist_dev: int, out_dev: int):
        check(self._lib.emo_mosaic_dev(self._h, C.c_void_p(src_dev), W, H, out_channels, tint_alpha, C.c_void_p(item_dev),
                                       C.c_void_p(dist_dev), C.c_void_p(out_dev)))

    def mosaic(self, src, out_channels: int = 3, tint_alpha: int = 0, out: np.ndarray | None = None, want_maps=True):
        """Whole path with host buffers: match + compose (+tint)."""
        src = _u8(src)
        H, W = src.shape[:2]
        d = max(self.dim, 1)
        bh, bw = H // d, W // d
        if out is None:
            out = np.zeros((bh * self.ts, bw * self.ts, out_channels), np.uint8)
        item = np.zeros((bh, bw), np.int32) if want_maps else None
        dist = np.zeros((bh, bw), np.uint32) if want_maps else None
        check(self._lib.emo_mosaic(self._h, _ptr(src), W, H, out_channels, tint_alpha, _ptr(item), _ptr(dist), _ptr(out)))
        return out, item, dist


def stripe_bounds(units: int, world: int, rank: int):
    """emo_stripe_bounds: the contiguous [start, stop) of `units` block rows / tiles that part `rank` of `world` owns."""
    a, b = C.c_uint64(), C.c_uint64()
    _lib.load().emo_stripe_bounds(units, world, rank, C.byref(a), C.byref(b))
    return int(a.value), int(b.value)


class Group:
    """Several GPUs driven from one process (``emo_group``): the library is replicated with one NCCL broadcast, the
    source's block rows (rendering.rs:68-89) / the library's tiles (main.rs:760-794) are split into contiguous ranges, and
    every GPU copies its stripe of the result straight into the caller's array.  Same results as one ``Context``."""

    def __init__(self, devices: Sequence[int] | int = 1):
        self._lib = _lib.load()
        if isinstance(devices, int):
            devices = list(range(devices))
        devices = [int(d) for d in devices]
        arr = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        check(self._lib.emo_group_create(arr, len(devices), C.byref(h)))
        self._h = h
        self.devices = devices
        self.members = [Context(d, _handle=self._lib.emo_group_ctx(h, i)) for i, d in enumerate(devices)]
        self.N = self.dim = self.ts = self.T = 0

    def __len__(self):
        return len(self.devices)

    def close(self):
        if getattr(self, "_h", None):
            for m in self.members:
                m.close()
            self._lib.emo_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        for m in self.members:
            m.sync()

    def launch_count(self) -> int:
        return sum(m.launch_count() for m in self.members)

    def set_match_mode(self, mode):
        for m in self.members:
            m.set_match_mode(mode)

    def set_library(self, colors, tile_px=None):
        colors = _u8(colors)
        if colors.ndim != 3 or colors.shape[2] != 3:
            raise EmosaicError(EMO_ERR_ARG, f"colors must be [T,N,3], got {colors.shape}")
        T, N = colors.shape[:2]
        ts = 0
        if tile_px is not None:
            tile_px = _u8(tile_px)
            if tile_px.ndim != 4 or tile_px.shape[0] != T or tile_px.shape[1] != tile_px.shape[2] or tile_px.shape[3] != 3:
                raise EmosaicError(EMO_ERR_ARG, f"tile_px must be [T,ts,ts,3] with T={T}, got {tile_px.shape}")
            ts = tile_px.shape[1]
        check(self._lib.emo_group_set_library(self._h, _ptr(colors), _ptr(tile_px), T, N, ts))
        self.T, self.N, self.dim, self.ts = T, N, _isqrt_exact(N), ts
        for m in self.members:
            m.T, m.N, m.dim, m.ts = self.T, self.N, self.dim, self.ts

    def analyse_tiles(self, tiles, dim: int) -> np.ndarray:
        tiles = _u8(tiles)
        if tiles.ndim != 4 or tiles.shape[1] != tiles.shape[2] or tiles.shape[3] != 3:
            raise EmosaicError(EMO_ERR_ARG, f"tiles must be [T,ts,ts,3], got {tiles.shape}")
        T, ts = tiles.shape[:2]
        out = np.zeros((T, dim * dim, 3), np.uint8)
        check(self._lib.emo_group_analyse(self._h, _ptr(tiles), T, ts, dim, _ptr(out)))
        return out

    def analyse_tiles_fused(self, tiles):
        tiles = _u8(tiles)
        T, ts = tiles.shape[:2]
        o1 = np.zeros((T, 1, 3), np.uint8)
        o4 = np.zeros((T, 4, 3), np.uint8)
        check(self._lib.emo_group_analyse_fused(self._h, _ptr(tiles), T, ts, _ptr(o1), _ptr(o4)))
        return o1, o4

    def mosaic(self, src, out_channels: int = 3, tint_alpha: int = 0, out: np.ndarray | None = None, want_maps=True):
        src = _u8(src)
        H, W = src.shape[:2]
        d = max(self.dim, 1)
        bh, bw = H // d, W // d
        if out is None:
            out = np.zeros((bh * self.ts, bw * self.ts, out_channels), np.uint8)
        item = np.zeros((bh, bw), np.int32) if want_maps else None
        dist = np.zeros((bh, bw), np.uint32) if want_maps else None
        check(self._lib.emo_group_mosaic(self._h, _ptr(src), W, H, out_channels, tint_alpha, _ptr(item), _ptr(dist), _ptr(out)))
        return out, item, dist

    # single-GPU services a renderer needs besides the sharded calls run on the first member
    def resize(self, *a, **k):
        return self.members[0].resize(*a, **k)

    def compose(self, *a, **k):
        return self.members[0].compose(*a, **k)

    def compose_overlay(self, *a, **k):
        return self.members[0].compose_overlay(*a, **k)

    def topk(self, *a, **k):
        return self.members[0].topk(*a, **k)

    def stats(self, *a, **k):
        return self.members[0].stats(*a, **k)


def host_register(arr: np.ndarray):
    """Pin an existing host array (cudaHostRegister) so the per-GPU copies of the group calls run at full PCIe rate."""
    check(_lib.load().emo_host_register(C.c_void_p(arr.ctypes.data), arr.nbytes))


def host_unregister(arr: np.ndarray):
    check(_lib.load().emo_host_unregister(C.c_void_p(arr.ctypes.data)))


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


# ---------------------------------------------------------------------------------------------
# reference-shaped interface
# ---------------------------------------------------------------------------------------------
def analyse(img, N: int, ctx: Context | None = None) -> np.ndarray:
    """``analyse::<N>(img)`` (analysis.rs:5-20): square RGB image -> [N,3] cell means."""
    img = _u8(img)
    if img.ndim != 3 or img.shape[2] != 3 or img.shape[0] != img.shape[1]:
        raise EmosaicError(EMO_ERR_ARG, f"analyse: tiles are square RGB images on this path, got {img.shape}")
    dim = _isqrt_exact(N)
    return (ctx or default_context()).analyse_tiles(img[None], dim)[0]


def analyse_tiles(tiles, N: int, ctx: Context | None = None) -> np.ndarray:
    return (ctx or default_context()).analyse_tiles(tiles, _isqrt_exact(N))


def get_img_colors(x: int, y: int, step: int, source_img, N: int) -> np.ndarray:
    """``get_img_colors::<N>`` (analysis.rs:23-36).  Host-side view; the match kernel does this gather itself."""
    source_img = _u8(source_img)
    out = np.zeros((N, 3), np.uint8)
    for i in range(N):
        out[i] = source_img[y + i // step, x + i % step]
    return out


def flipped_coords(coords: Sequence[int]) -> np.ndarray:
    """tiles/utils.rs:18-43: mirror a flattened [rows*cols*3] coordinate vector row by row."""
    c = np.array(coords).copy()
    n = c.size
    rows = math.isqrt(n // 3)
    cols = rows
    cir = cols * 3
    for i in range(rows):
        for j in range(cols // 2):
            a, b = i * cir + j * 3, (i + 1) * cir - (j + 1) * 3
            for h in range(3):
                c[a + h], c[b + h] = c[b + h], c[a + h]
    return c


@dataclass
class Tile:
    """tiles/tile.rs:9-17."""
    colors: np.ndarray  # [N,3] u8
    idx: int = 0
    flipped: bool = False
    date_taken: Optional[str] = None

    @staticmethod
    def from_colors(colors) -> "Tile":
        return Tile(_u8(colors).reshape(-1, 3), 0)

    def coords(self) -> np.ndarray:
        """tile.rs:106-119: [3N] u32, r,g,b interleaved, mirrored when flipped."""
        c = self.colors.reshape(-1).astype(np.uint32)
        return flipped_coords(c).astype(np.uint32) if self.flipped else c


@dataclass
class TileSet:
    """tiles/tileset.rs:21-26 — colours, paths and (optionally) in-memory tile images."""
    N: int
    colors: List[np.ndarray] = field(default_factory=list)
    paths: List[str] = field(default_factory=list)
    dates: List[Optional[str]] = field(default_factory=list)
    images: dict = field(default_factory=dict)

    def __len__(self):
        return len(self.colors)

    def push_tile(self, path: str, colors, date_taken: Optional[str] = None):
        self.colors.append(_u8(colors).reshape(self.N, 3))
        self.paths.append(path)
        self.dates.append(date_taken)

    def push_tile_with_image(self, path: str, colors, image):
        self.push_tile(path, colors)
        self.images[len(self.colors)] = _u8(image)

    @staticmethod
    def from_arrays(colors, tile_px=None, paths=None, dates=None) -> "TileSet":
        colors = _u8(colors)
        ts = TileSet(N=colors.shape[1])
        ts.colors = [c for c in colors]
        ts.paths = list(paths) if paths is not None else [f"tile{i:07d}.jpg" for i in range(len(colors))]
        ts.dates = list(dates) if dates is not None else [None] * len(colors)
        if tile_px is not None:
            ts._px = _u8(tile_px)
        return ts

    def colors_array(self) -> np.ndarray:
        return np.ascontiguousarray(np.stack(self.colors), dtype=np.uint8) if self.colors else np.zeros((0, self.N, 3), np.uint8)

    def pixels_array(self, tile_size: int) -> np.ndarray:
        if getattr(self, "_px", None) is not None:
            return self._px
        T = len(self)
        if len(self.images) != T:
            raise EmosaicError(EMO_ERR_ARG, "Image not found: tiles without in-memory images must be prepared on the host "
                               "(prepare_tile is outside the accelerated path)")
        px = np.stack([self.images[i + 1] for i in range(T)])
        if px.shape[1] != tile_size or px.shape[2] != tile_size:
            raise EmosaicError(EMO_ERR_ARG, f"tile images are {px.shape[1:3]}, expected {tile_size}x{tile_size}")
        return np.ascontiguousarray(px, dtype=np.uint8)

    def get_tile(self, idx: int) -> Optional[Tile]:
        """tileset.rs:131-143: positive = normal, negative = flipped, 1-based."""
        a = abs(idx)
        if a == 0 or a > len(self):
            return None
        return Tile(self.colors[a - 1], a, idx < 0, self.dates[a - 1])

    def get_path(self, tile: Tile) -> str:
        return self.paths[tile.idx - 1]

    def build_kiddo(self, ctx: Context | None = None, tile_size: int | None = None) -> Context:
        """tileset.rs:178-190: make the (tile, mirrored tile) search set resident on the GPU."""
        ctx = ctx or default_context()
        px = self.pixels_array(tile_size) if tile_size else None
        if len(self) == 0:
            raise EmosaicError(EMO_ERR_ARG, "empty tile set")
        ctx.set_library(self.colors_array(), px)
        return ctx


@dataclass
class RenderResult:
    """rendering.rs:236-243.  `stats` is the pair of maps the reference's RenderStats is fed with."""
    image: np.ndarray
    tile_set: TileSet
    item: np.ndarray
    dist: np.ndarray


def adjust_source_dims(w: int, h: int, downsample: int, dim: int):
    """main.rs:567-587: size the source is resized to before matching."""
    nw, nh = w // downsample, h // downsample
    m = nw % dim
    nw = nw + dim - m if m > dim // 2 else nw - m
    m = nh % dim
    nh = nh + dim - m if m > dim // 2 else nh - m
    return nw, nh


def resize_taps(n_in: int, n_out: int):
    """The tap table of one axis as emo_resize computes it (host-only): (left [n_out], cnt [n_out], weights [n_out, max taps])."""
    lib = _lib.load()
    mt = C.c_uint32(0)
    check(lib.emo_resize_taps(n_in, n_out, None, None, None, 0, C.byref(mt)))
    left = np.zeros(n_out, np.uint32)
    cnt = np.zeros(n_out, np.uint32)
    ws = np.zeros((n_out, mt.value), np.float32)
    check(lib.emo_resize_taps(n_in, n_out, _ptr(left), _ptr(cnt), _ptr(ws), mt.value, None))
    return left, cnt, ws


def resize_source(original_img, downsample: int, dim: int, ctx: Context | None = None) -> np.ndarray:
    """main.rs:567-595: the copy of the source that is matched — ``imageops::resize(original, nw, nh, Lanczos3)`` with
    (nw, nh) from the dimension rule; equal dimensions give a plain copy, as in the crate."""
    original_img = _u8(original_img)
    nw, nh = adjust_source_dims(original_img.shape[1], original_img.shape[0], downsample, dim)
    if nw == 0 or nh == 0:
        raise EmosaicError(EMO_ERR_ARG, f"Invalid source dimensions ({original_img.shape[1]}x{original_img.shape[0]}): nothing "
                           f"left after downsampling by {downsample} to a multiple of {dim}")
    return (ctx or default_context()).resize(original_img, nw, nh)


def _most_common_value(values: np.ndarray) -> int:
    """tiles/utils.rs:262-273.  HashMap + max_by_key leaves the winner among equally frequent values to the hash order;
    the canonical rule is the smallest such value.  Empty input -> 0 (``unwrap_or((0, 0)).0``)."""
    if values.size == 0:
        return 0
    vals, counts = np.unique(values, return_counts=True)
    return int(vals[np.argmax(counts)])


def prepare_view(img, tile_size: int, crop: bool):
    """The view of a decoded photo that ``prepare_tile`` resizes (tiles/utils.rs:93-186): white borders (every channel
    > 240) are trimmed to the most common first / last non-white column and row — the view is [first, last) on both
    axes — then ``crop`` takes the centred largest square.  Host-side index bookkeeping; no pixel is produced here."""
    img = _u8(img)
    h, w = img.shape[:2]
    if w < tile_size or h < tile_size:  # utils.rs:99-106
        raise EmosaicError(EMO_ERR_ARG, f"Image {w}x{h} is smaller than the tile size {tile_size} (DimensionError)")
    nonwhite = ~((img[..., 0] > 240) & (img[..., 1] > 240) & (img[..., 2] > 240))

    def first_last(m):
        n = m.shape[1]
        has = m.any(axis=1)
        return np.where(has, m.argmax(axis=1), n), np.where(has, n - 1 - m[:, ::-1].argmax(axis=1), 0)

    from_left, from_right = first_last(nonwhite)
    from_top, from_bottom = first_last(nonwhite.T)
    c0 = _most_common_value(from_left[from_left != w])
    c1 = _most_common_value(from_right[from_right != 0])
    r0 = _most_common_value(from_top[from_top != h])
    r1 = _most_common_value(from_bottom[from_bottom != 0])
    if not (c0 < c1 and r0 < r1):  # utils.rs:157-158 assert!
        raise EmosaicError(EMO_ERR_ARG, "assertion failed: first_non_white_col < last_non_white_col / first_non_white_row < "
                           "last_non_white_row (no non-white interior)")
    vx, vy, vw, vh = c0, r0, c1 - c0, r1 - r0
    if crop:  # utils.rs:170-182
        size = min(vw, vh)
        vx, vy, vw, vh = vx + (vw - size) // 2, vy + (vh - size) // 2, size, size
    return vx, vy, vw, vh


def rotate(img: np.ndarray, orientation: int) -> np.ndarray:
    """tiles/utils.rs:248-264: undo the EXIF orientation (1..8) of an already resized tile."""
    r90 = lambda a: np.rot90(a, -1)   # imageops::rotate90 = clockwise
    r270 = lambda a: np.rot90(a, 1)
    ops = {2: lambda a: a[:, ::-1], 3: lambda a: a[::-1, ::-1], 4: lambda a: a[::-1],
           5: lambda a: r90(a)[:, ::-1], 6: r90, 7: lambda a: r270(a)[:, ::-1], 8: r270}
    return np.ascontiguousarray(ops.get(orientation, lambda a: a)(img))


def prepare_tile(img, tile_size: int, crop: bool, orientation: int = 1, ctx: Context | None = None) -> np.ndarray:
    """``prepare_tile`` (tiles/utils.rs:63-196) from the decoded photo on: trim + crop view, Lanczos3 resize to
    tile_size x tile_size on the GPU, EXIF rotation.  File read, MD5 and the JPEG cache stay in the caller."""
    view = prepare_view(img, tile_size, crop)
    return rotate((ctx or default_context()).resize(img, tile_size, tile_size, view), orientation)


def tint_alpha(tint_opacity: float) -> int:
    """main.rs:449 ``(255.0 * tint_opacity) as u8`` (saturating, truncating)."""
    v = 255.0 * float(tint_opacity)
    if not (v > 0.0):
        return 0
    return 255 if v >= 255.0 else int(v)


def render_nto1(source_img, tile_set: TileSet, tile_size: int, no_repeat: bool = False, randomize: Optional[float] = None,
                tint_opacity: float = 0.0, ctx: Context | None = None) -> RenderResult:
    """``render_nto1::<N>`` (rendering.rs:124-230) [+ tint block main.rs:447-478 when tint_opacity > 0]."""
    if no_repeat or randomize is not None:
        raise EmosaicError(EMO_ERR_UNSUPPORTED, "no_repeat (greedy, rayon-order dependent) / randomize (thread_rng) inside "
                           "render_nto1 are nondeterministic host algorithms outside the accelerated path "
                           "(rendering.rs:163-209); the deterministic no-repeat renderer is render_nto1_no_repeat")
    source_img = _u8(source_img)
    dim = _isqrt_exact(tile_set.N)
    H, W = source_img.shape[:2]
    if W % dim or H % dim:
        raise EmosaicError(EMO_ERR_ARG, f"Invalid source dimensions ({W}x{H}): Dimensions must be divisible by {dim}")
    if tile_size % dim:
        raise EmosaicError(EMO_ERR_ARG, f"Invalid tile size: Tile size must be divisible by {dim}")
    ctx = tile_set.build_kiddo(ctx, tile_size)
    if tint_opacity > 0.0:
        image, item, dist = ctx.mosaic(source_img, 4, tint_alpha(tint_opacity))
    else:
        image, item, dist = ctx.mosaic(source_img, 3, 0)
    return RenderResult(image, tile_set, item, dist)


def render_nto1_no_repeat(source_img, tile_set: TileSet, tile_size: int, ctx: Context | None = None, page: int = 0) -> RenderResult:
    """``render_nto1_no_repeat::<N>`` (rendering.rs:262-401): every block gets the nearest tile that no nearer block has
    taken; a tile is used once, in either orientation.  The whole assignment is one library call (emo_no_repeat: ranked
    candidate lists from the GPU in pages, the greedy merge of rendering.rs:341-392 on the host inside the library, blocks
    ordered by (distance, reference block number n = bx * vtiles + by), the canonical order of DESIGN.md); a block that runs
    out of candidates stays black (:347-351).  The image is composed on the GPU.  page = candidates per block of the first
    page (0: automatic; small values force the refill path)."""
    source_img = _u8(source_img)
    dim = _isqrt_exact(tile_set.N)
    H, W = source_img.shape[:2]
    if W % dim or H % dim:
        raise EmosaicError(EMO_ERR_ARG, f"Invalid source dimensions ({W}x{H}): Dimensions must be divisible by {dim}")
    if tile_size % dim:
        raise EmosaicError(EMO_ERR_ARG, f"Invalid tile size: Tile size must be divisible by {dim}")
    bh, bw = H // dim, W // dim
    T = len(tile_set)
    if bh * bw > 2 * T:  # rendering.rs:292-298
        raise EmosaicError(EMO_ERR_ARG, f"Insufficient tiles for no-repeat mode: need {bh * bw} tiles but only have {2 * T} available")
    ctx = tile_set.build_kiddo(ctx, tile_size)
    member = ctx.members[0] if isinstance(ctx, Group) else ctx       # the merge is sequential: one GPU serves the lists
    item, dist, _ = member.no_repeat(source_img, page)
    placed = item != 0
    image = ctx.compose(np.where(placed, item, 1).astype(np.int32))
    if not placed.all():                                         # RgbImage::new: unplaced blocks stay black
        mask = np.repeat(np.repeat(~placed, tile_size, 0), tile_size, 1)
        image[mask] = 0
    return RenderResult(image, tile_set, item, dist)


def apply_tint(output, source_img, tint_opacity: float, tile_set: TileSet, tile_size: int, item,
               ctx: Context | None = None) -> np.ndarray:
    """The tint block alone (main.rs:447-478) for an already matched item map: RGBA result."""
    ctx = tile_set.build_kiddo(ctx, tile_size)
    return ctx.compose(item, source_img, 4, tint_alpha(tint_opacity))
