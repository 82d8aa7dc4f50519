"""``.emosaic_*`` analysis cache — host-side, byte-compatible with the reference.

Layout (bincode 1.3.3 defaults: little-endian, fixed-width ints, u64 lengths), from
``impl Serialize for TileSet`` (tiles/tileset.rs:28-49) and ``impl Serialize for Tile``
(tiles/tile.rs:38-52):

    u64 T | T x { u64 3N | 3N colour bytes | u16 idx | u8 tag [| u64 len | utf-8 date] }
          | u64 T | T x { u64 len | utf-8 path }

Loader semantics follow main.rs:617-654: the stored idx is ignored, entries whose file is missing
or whose extension is not allowed are dropped, survivors are renumbered 1..n.
"""
from __future__ import annotations

import os
import struct
from typing import Iterable, List, Optional, Tuple

import numpy as np


def cache_file_name(N: int, crop: bool) -> str:
    """main.rs:597-601."""
    return f".emosaic_{N}to1{'_cropped' if crop else ''}"


def serialize_tile_set(colors: np.ndarray, paths: List[str], dates: Optional[List[Optional[str]]] = None,
                       idx: Optional[Iterable[int]] = None) -> bytes:
    colors = np.ascontiguousarray(colors, dtype=np.uint8)
    T, N = colors.shape[0], colors.shape[1]
    if dates is None:
        dates = [None] * T
    # main.rs:791 `(idx + 1) as u16` wraps silently for T > 65535
    idx = [(i + 1) & 0xFFFF for i in range(T)] if idx is None else [int(i) & 0xFFFF for i in idx]
    out = bytearray()
    out += struct.pack("<Q", T)
    for t in range(T):
        out += struct.pack("<Q", 3 * N)
        out += colors[t].tobytes()
        out += struct.pack("<H", idx[t])
        d = dates[t]
        if d is None:
            out += b"\x00"
        else:
            b = d.encode("utf-8")
            out += b"\x01" + struct.pack("<Q", len(b)) + b
    out += struct.pack("<Q", T)
    for p in paths:
        b = os.fsencode(p) if not isinstance(p, bytes) else p
        out += struct.pack("<Q", len(b)) + b
    return bytes(out)


def deserialize_tile_set(buf: bytes, N: int, extensions: Optional[Iterable[str]] = None, check_exists: bool = False
                         ) -> Tuple[np.ndarray, List[str], List[Optional[str]]]:
    """Returns (colors [n,N,3], paths, dates) after the reference's filter + renumber step.
    Raises ValueError on malformed input (the reference's `.ok()` then falls back to re-analysis)."""
    mv = memoryview(buf)
    pos = 0

    def take(n):
        nonlocal pos
        if pos + n > len(mv):
            raise ValueError("truncated cache file")
        b = mv[pos:pos + n]
        pos += n
        return b

    (T,) = struct.unpack("<Q", take(8))
    # every tile costs at least 8 + 3N + 3 bytes here and 8 more in the path list: a corrupt count fails before any allocation
    if T > (len(mv) - pos) // (3 * N + 19):
        raise ValueError("truncated cache file")
    colors = np.zeros((T, N, 3), np.uint8)
    dates: List[Optional[str]] = []
    for t in range(T):
        (ln,) = struct.unpack("<Q", take(8))
        if ln != 3 * N:
            raise ValueError(f"tile {t}: {ln} colour bytes, expected {3 * N}")  # try_into().unwrap() tileset.rs:66
        colors[t] = np.frombuffer(take(ln), np.uint8).reshape(N, 3)
        take(2)  # stored idx: ignored by the loader (renumbered, main.rs:643-652)
        tag = take(1)[0]
        if tag == 0:
            dates.append(None)
        elif tag == 1:
            (dl,) = struct.unpack("<Q", take(8))
            dates.append(bytes(take(dl)).decode("utf-8"))
        else:
            raise ValueError("bad Option tag")
    (T2,) = struct.unpack("<Q", take(8))
    if T2 != T:
        raise ValueError("tiles / paths length mismatch")
    paths = []
    for _ in range(T):
        (pl,) = struct.unpack("<Q", take(8))
        paths.append(os.fsdecode(bytes(take(pl))))
    keep = list(range(T))
    if extensions is not None or check_exists:
        exts = set(extensions) if extensions is not None else None
        keep = []
        for t, p in enumerate(paths):
            ext = os.path.splitext(p)[1][1:]
            if not ext:
                continue
            if exts is not None and ext not in exts:
                continue
            if check_exists and not os.path.exists(p):
                continue
            keep.append(t)
    return colors[keep], [paths[t] for t in keep], [dates[t] for t in keep]
