"""CPU oracle for the emosaic hot path — TEST INFRASTRUCTURE ONLY.

ctypes front-end to ``oracle/liboracle.so`` (built from ``emosaic_oracle.c`` by
``oracle/Makefile``) plus a numpy restatement (``oracle_np``) used to cross-check
the C code.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package; nothing under
``emosaic_b200/`` does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

u8p = C.POINTER(C.c_uint8)
i32p = C.POINTER(C.c_int32)
u32p = C.POINTER(C.c_uint32)


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "emosaic_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        _LIB.orc_kd_build.restype = C.c_void_p
        _LIB.orc_cache_serialize.restype = C.c_int64
        _LIB.orc_tint_alpha.restype = C.c_uint8
        _LIB.orc_tint_alpha.argtypes = [C.c_double]
    return _LIB


def _p(a, t=u8p):
    return a.ctypes.data_as(t)


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a


ERRORS = {
    -1: "Rectangle dimensions must be positive",
    -2: "Rectangle extends beyond image width",
    -3: "Rectangle extends beyond image height",
    -4: "bad argument",
    -5: "Closest item should not be zero",
}


class OracleError(AssertionError):
    pass


def _chk(rc):
    if rc != 0:
        raise OracleError(ERRORS.get(rc, f"oracle error {rc}"))


def num_threads() -> int:
    return int(lib().orc_num_threads())


def average_color(img, rect):
    """color.rs:14-42.  img [h,w,3] u8, rect=(left, top, width, height)."""
    img = _u8(img)
    h, w = img.shape[:2]
    out = np.zeros(3, np.uint8)
    left, top, rw, rh = rect
    _chk(lib().orc_average_color(_p(img), C.c_uint32(w), C.c_uint32(h), C.c_uint32(left), C.c_uint32(top),
                                 C.c_uint32(rw), C.c_uint32(rh), _p(out)))
    return out


def analyse(img, N):
    """analysis.rs:5-20.  img [h,w,3] -> [N,3]."""
    img = _u8(img)
    h, w = img.shape[:2]
    out = np.zeros((N, 3), np.uint8)
    _chk(lib().orc_analyse(_p(img), C.c_uint32(w), C.c_uint32(h), C.c_uint32(N), _p(out)))
    return out


def analyse_tiles(tiles, N):
    """main.rs:786-794 over the whole library.  tiles [T,ts,ts,3] -> [T,N,3]."""
    tiles = _u8(tiles)
    T, ts = tiles.shape[0], tiles.shape[1]
    out = np.zeros((T, N, 3), np.uint8)
    _chk(lib().orc_analyse_tiles(_p(tiles), C.c_uint64(T), C.c_uint32(ts), C.c_uint32(N), _p(out)))
    return out


def get_img_colors(src, x, y, step, N):
    """analysis.rs:23-36."""
    src = _u8(src)
    out = np.zeros((N, 3), np.uint8)
    lib().orc_get_img_colors(_p(src), C.c_uint32(src.shape[1]), C.c_uint32(x), C.c_uint32(y), C.c_uint32(step),
                             C.c_uint32(N), _p(out))
    return out


def coords(colors, flipped=False):
    """tile.rs:106-119."""
    colors = _u8(colors).reshape(-1, 3)
    N = colors.shape[0]
    out = np.zeros(3 * N, np.uint32)
    lib().orc_coords(_p(colors), C.c_uint32(N), C.c_int(bool(flipped)), _p(out, u32p))
    return out


def flipped_coords(c):
    """utils.rs:18-43 (returns a flipped copy)."""
    c = np.array(c, dtype=np.uint32).copy()
    lib().orc_flipped_coords_u32(_p(c, u32p), C.c_uint32(c.size))
    return c


def build_candidates(colors):
    """tileset.rs:178-190: candidate vectors [2T,3N] + items [2T] in insertion order."""
    colors = _u8(colors)
    T, N = colors.shape[0], colors.shape[1]
    cand = np.zeros((2 * T, 3 * N), np.uint8)
    items = np.zeros(2 * T, np.int32)
    lib().orc_build_candidates(_p(colors), C.c_uint32(T), C.c_uint32(N), _p(cand), _p(items, i32p))
    return cand, items


def match(colors, src):
    """rendering.rs:158-221 match stage; brute force, canonical tie-break.
    colors [T,N,3], src [H,W,3] -> (item [H/dim,W/dim] i32 signed 1-based, dist u32)."""
    colors, src = _u8(colors), _u8(src)
    T, N = colors.shape[0], colors.shape[1]
    dim = int(round(N ** 0.5))
    H, W = src.shape[:2]
    item = np.zeros((H // dim, W // dim), np.int32)
    dist = np.zeros((H // dim, W // dim), np.uint32)
    _chk(lib().orc_match(_p(colors), C.c_uint32(T), C.c_uint32(N), _p(src), C.c_uint32(W), C.c_uint32(H),
                         _p(item, i32p), _p(dist, u32p)))
    return item, dist


def no_repeat_assign(colors, src):
    """rendering.rs:262-392 up to the placement of tiles (C restatement; oracle_np.no_repeat_assign is the independent numpy
    one): colors [T,N,3], src [H,W,3] -> (item [bh,bw] int32 with 0 = unplaced, dist [bh,bw] uint32)."""
    colors, src = _u8(colors), _u8(src)
    T, N = colors.shape[0], colors.shape[1]
    dim = int(round(N ** 0.5))
    H, W = src.shape[:2]
    if (H // dim) * (W // dim) > 2 * T:
        raise OracleError(f"Insufficient tiles for no-repeat mode: need {(H // dim) * (W // dim)} tiles but only have {2 * T} available")
    item = np.zeros((H // dim, W // dim), np.int32)
    dist = np.zeros((H // dim, W // dim), np.uint32)
    _chk(lib().orc_no_repeat(_p(colors), C.c_uint32(T), C.c_uint32(N), _p(src), C.c_uint32(W), C.c_uint32(H),
                             _p(item, i32p), _p(dist, u32p)))
    return item, dist


class KdTree:
    """Bucketed (640) L1 KD-tree over the mirrored candidate set — the reference's algorithm class."""

    def __init__(self, colors, bucket=640):
        colors = _u8(colors)
        self.N = colors.shape[1]
        cand, _ = build_candidates(colors)
        self._cand = cand
        self._h = C.c_void_p(lib().orc_kd_build(_p(cand), C.c_uint32(cand.shape[0]), C.c_uint32(cand.shape[1]),
                                                C.c_uint32(bucket)))

    def match(self, src):
        src = _u8(src)
        dim = int(round(self.N ** 0.5))
        H, W = src.shape[:2]
        item = np.zeros((H // dim, W // dim), np.int32)
        dist = np.zeros((H // dim, W // dim), np.uint32)
        _chk(lib().orc_match_kd(self._h, _p(src), C.c_uint32(W), C.c_uint32(H), C.c_uint32(self.N),
                                _p(item, i32p), _p(dist, u32p)))
        return item, dist

    def __del__(self):
        try:
            lib().orc_kd_free(self._h)
        except Exception:
            pass


def render(tile_px, item, out=None):
    """rendering.rs:51-101 + tileset.rs:131-161.  tile_px [T,ts,ts,3], item [bh,bw] -> [bh*ts,bw*ts,3]."""
    tile_px = _u8(tile_px)
    item = np.ascontiguousarray(item, dtype=np.int32)
    T, ts = tile_px.shape[0], tile_px.shape[1]
    bh, bw = item.shape
    if out is None:
        out = np.zeros((bh * ts, bw * ts, 3), np.uint8)
    _chk(lib().orc_render(_p(tile_px), C.c_uint32(T), C.c_uint32(ts), _p(item, i32p), C.c_uint32(bw), C.c_uint32(bh),
                          _p(out)))
    return out


def tint_alpha(tint_opacity: float) -> int:
    """main.rs:449."""
    return int(lib().orc_tint_alpha(float(tint_opacity)))


def tint(mosaic, src, A):
    """main.rs:447-478.  mosaic [OH,OW,3], src [H,W,3], alpha byte A -> RGBA [OH,OW,4]."""
    mosaic, src = _u8(mosaic), _u8(src)
    OH, OW = mosaic.shape[:2]
    H, W = src.shape[:2]
    out = np.zeros((OH, OW, 4), np.uint8)
    lib().orc_tint(_p(mosaic), C.c_uint32(OW), C.c_uint32(OH), _p(src), C.c_uint32(W), C.c_uint32(H), C.c_uint8(A),
                   _p(out))
    return out


def blend_lut(A):
    """[256 bg, 256 fg] table of the f32 src-over blend and the output alpha byte."""
    lut = np.zeros((256, 256), np.uint8)
    a = C.c_uint8(0)
    lib().orc_blend_lut(C.c_uint8(A), _p(lut), C.byref(a))
    return lut, int(a.value)


def adjust_dims(w, h, downsample, dim):
    """main.rs:567-587."""
    nw, nh = C.c_uint32(0), C.c_uint32(0)
    lib().orc_adjust_dims(C.c_uint32(w), C.c_uint32(h), C.c_uint32(downsample), C.c_uint32(dim), C.byref(nw), C.byref(nh))
    return int(nw.value), int(nh.value)


def cache_serialize(colors, idx, dates, paths) -> bytes:
    """tileset.rs:28-49 / tile.rs:38-52 under bincode 1.3.3 defaults."""
    colors = _u8(colors)
    T, N = colors.shape[0], colors.shape[1]
    idx = np.ascontiguousarray(idx, dtype=np.uint16)
    d_arr = (C.c_char_p * max(T, 1))(*[None if d is None else d.encode() for d in dates])
    p_arr = (C.c_char_p * max(T, 1))(*[p.encode() for p in paths])
    need = lib().orc_cache_serialize(_p(colors), C.c_uint32(T), C.c_uint32(N), idx.ctypes.data_as(C.POINTER(C.c_uint16)),
                                     d_arr, p_arr, None, C.c_uint64(0))
    buf = (C.c_uint8 * need)()
    n = lib().orc_cache_serialize(_p(colors), C.c_uint32(T), C.c_uint32(N), idx.ctypes.data_as(C.POINTER(C.c_uint16)),
                                  d_arr, p_arr, buf, C.c_uint64(need))
    assert n == need
    return bytes(buf)


def resize_lanczos3(img, nw, nh, view=None):
    """image 0.25.2 imageops::resize(view, nw, nh, Lanczos3) — main.rs:595, tiles/utils.rs:188-189.
    img [H,W,3]; view = (x0, y0, cw, ch) or None for the whole image -> [nh,nw,3]."""
    img = _u8(img)
    H, W = img.shape[:2]
    x0, y0, cw, ch = view if view is not None else (0, 0, W, H)
    out = np.zeros((nh, nw, 3), np.uint8)
    rc = lib().orc_resize_lanczos3(_p(img), C.c_uint32(W), C.c_uint32(H), C.c_uint32(x0), C.c_uint32(y0), C.c_uint32(cw),
                                   C.c_uint32(ch), C.c_uint32(nw), C.c_uint32(nh), _p(out))
    if rc:
        raise ValueError(ERRORS.get(rc, str(rc)))
    return out


def resize_axis(n_in, n_out):
    """Taps of one axis: (left [out], cnt [out], weights [out, max cnt] f32, zero beyond cnt)."""
    pitch = lib().orc_resize_axis(C.c_uint32(n_in), C.c_uint32(n_out), None, None, None, C.c_uint32(0))
    left = np.zeros(n_out, np.uint32)
    cnt = np.zeros(n_out, np.uint32)
    ws = np.zeros((n_out, pitch), np.float32)
    lib().orc_resize_axis(C.c_uint32(n_in), C.c_uint32(n_out), _p(left, u32p), _p(cnt, u32p),
                          ws.ctypes.data_as(C.POINTER(C.c_float)), C.c_uint32(pitch))
    return left, cnt, ws


def prepare_view(img, tile_size, crop):
    """tiles/utils.rs:93-186: the (x0, y0, w, h) view prepare_tile resizes (white-border trim by the mode of the
    per-row / per-column first and last non-white positions, then the centred square when crop)."""
    img = _u8(img)
    H, W = img.shape[:2]
    v = (C.c_uint32 * 4)()
    rc = lib().orc_prepare_view(_p(img), C.c_uint32(W), C.c_uint32(H), C.c_uint32(tile_size), C.c_int(int(bool(crop))), v)
    if rc:
        raise ValueError("prepare_tile: image smaller than the tile size, or no non-white interior (utils.rs:99-106, :157-158)")
    return tuple(int(x) for x in v)
