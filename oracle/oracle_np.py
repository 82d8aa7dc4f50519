"""Independent numpy restatement of the same reference functions (TEST INFRASTRUCTURE).

Written separately from emosaic_oracle.c so the two can be cross-checked against each
other and against the reference's own known-answer vectors.  Citations as in the C file.
"""
from __future__ import annotations

import numpy as np


def analyse(img: np.ndarray, N: int) -> np.ndarray:
    """analysis.rs:5-20 + color.rs:14-42 (u64 sums, truncating division)."""
    h, w = img.shape[:2]
    dim = int(np.sqrt(N))
    cw, ch = int(np.floor(w / np.sqrt(N))), int(np.floor(h / np.sqrt(N)))
    assert cw > 0 and ch > 0, "Rectangle dimensions must be positive"
    out = np.zeros((N, 3), np.uint8)
    for i in range(N):
        top, left = i // dim, i % dim
        cell = img[top * ch:(top + 1) * ch, left * cw:(left + 1) * cw].astype(np.uint64)
        out[i] = (cell.sum(axis=(0, 1)) // np.uint64(cw * ch)).astype(np.uint8)
    return out


def analyse_tiles(tiles: np.ndarray, N: int) -> np.ndarray:
    T, ts = tiles.shape[:2]
    dim = int(np.sqrt(N))
    c = ts // dim
    t = tiles[:, :dim * c, :dim * c].reshape(T, dim, c, dim, c, 3).astype(np.uint64)
    return (t.sum(axis=(2, 4)) // np.uint64(c * c)).astype(np.uint8).reshape(T, N, 3)


def queries(src: np.ndarray, N: int) -> np.ndarray:
    """analysis.rs:23-36 for every block: [bh,bw,N*3], cells row-major inside the block."""
    dim = int(np.sqrt(N))
    H, W = src.shape[:2]
    q = src.reshape(H // dim, dim, W // dim, dim, 3).transpose(0, 2, 1, 3, 4)
    return np.ascontiguousarray(q).reshape(H // dim, W // dim, N * 3)


def mirror(vecs: np.ndarray, N: int) -> np.ndarray:
    """utils.rs:18-43 on [..., 3N] vectors."""
    dim = int(np.sqrt(N))
    v = vecs.reshape(vecs.shape[:-1] + (dim, dim, 3))
    return np.ascontiguousarray(v[..., ::-1, :]).reshape(vecs.shape)


def match(colors: np.ndarray, src: np.ndarray, chunk: int = 4096):
    """tileset.rs:178-190 + rendering.rs:187-195 with the canonical tie-break."""
    T, N = colors.shape[:2]
    base = colors.reshape(T, 3 * N)
    cand = np.empty((2 * T, 3 * N), np.int32)
    cand[0::2] = base
    cand[1::2] = mirror(base, N)
    q = queries(src, N).astype(np.int32)
    bh, bw = q.shape[:2]
    qf = q.reshape(-1, 3 * N)
    item = np.zeros(qf.shape[0], np.int32)
    dist = np.zeros(qf.shape[0], np.uint32)
    for s in range(0, qf.shape[0], chunk):
        d = np.abs(qf[s:s + chunk, None, :] - cand[None, :, :]).sum(axis=2)
        r = d.argmin(axis=1)  # first minimum == lowest insertion rank
        dist[s:s + chunk] = d[np.arange(d.shape[0]), r]
        item[s:s + chunk] = np.where(r % 2 == 0, r // 2 + 1, -(r // 2 + 1))
    return item.reshape(bh, bw), dist.reshape(bh, bw)


def render(tile_px: np.ndarray, item: np.ndarray) -> np.ndarray:
    """rendering.rs:51-101."""
    T, ts = tile_px.shape[:2]
    bh, bw = item.shape
    a = np.abs(item) - 1
    px = tile_px[a]  # [bh,bw,ts,ts,3]
    px = np.where((item < 0)[:, :, None, None, None], px[:, :, :, ::-1, :], px)
    return np.ascontiguousarray(px.transpose(0, 2, 1, 3, 4)).reshape(bh * ts, bw * ts, 3)


def blend(bg: np.ndarray, fg: np.ndarray, A: int):
    """image 0.25.2 Rgba<u8>::blend with bg alpha 255 and fg alpha A. Returns (rgb, alpha_byte)."""
    f = np.float32
    if A == 0:
        return bg.copy(), 255
    if A == 255:
        return fg.copy(), 255
    m = f(255.0)
    bga, fga = f(255.0) / m, f(A) / m
    af = f(f(bga + fga) - f(bga * fga))
    b = bg.astype(np.float32) / m
    g = fg.astype(np.float32) / m
    o = ((g * fga).astype(np.float32) + ((b * bga).astype(np.float32) * f(f(1.0) - fga)).astype(np.float32)).astype(np.float32)
    o = (o / af).astype(np.float32)
    return np.trunc((m * o).astype(np.float32)).astype(np.uint8), int(np.trunc(m * af))


def tint(mosaic: np.ndarray, src: np.ndarray, A: int) -> np.ndarray:
    """main.rs:447-478."""
    OH, OW = mosaic.shape[:2]
    H, W = src.shape[:2]
    f = np.float32
    sx = np.minimum(np.floor((np.arange(OW, dtype=np.float32) + f(0.5)) * (f(W) / f(OW))).astype(np.int64), W - 1)
    sy = np.minimum(np.floor((np.arange(OH, dtype=np.float32) + f(0.5)) * (f(H) / f(OH))).astype(np.int64), H - 1)
    fg = src[sy][:, sx]
    rgb, a = blend(mosaic, fg, A)
    out = np.empty((OH, OW, 4), np.uint8)
    out[..., :3] = rgb
    out[..., 3] = a
    return out


def l1_voronoi_table(colors: np.ndarray, bits: int = 8):
    """Restatement (numpy, test infrastructure) of the construction in emosaic_b200/csrc/index.cu on a cube with
    `bits` bits per channel: for every colour of the cube the (L1 distance, tile) of the tile a scan with strict `<`
    would keep — minimum distance, ties to the smallest tile index.  Exact city-block distance transform: seed the
    library colours with key = dist << 22 | tile, then one forward/backward min-plus sweep per axis.
    colors [T,1,3] with values < 2**bits.  Returns (tile [S,S,S] int64 0-based indexed [b][g][r], dist [S,S,S])."""
    S = 1 << bits
    INC = 1 << 22
    EMPTY = np.int64(1) << 40
    lut = np.full((S, S, S), EMPTY, dtype=np.int64)          # [b][g][r]
    c = colors.reshape(-1, 3).astype(np.int64)
    for t in range(c.shape[0] - 1, -1, -1):                   # descending, so the smallest tile index is written last
        lut[c[t, 2], c[t, 1], c[t, 0]] = t
    for axis in (2, 1, 0):                                    # r (contiguous), g, b
        a = np.moveaxis(lut, axis, 0)
        for i in range(1, S):
            a[i] = np.minimum(a[i], a[i - 1] + INC)
        for i in range(S - 2, -1, -1):
            a[i] = np.minimum(a[i], a[i + 1] + INC)
    return lut & (INC - 1), lut >> 22


# ---- no-repeat renderer (rendering.rs:262-401) ---------------------------------------------------------------
def sorted_candidates(colors: np.ndarray, src: np.ndarray):
    """nearest_n::<Manhattan>(coords, 100000) for every block (rendering.rs:307-321): ALL candidates of the search set
    (tileset.rs:178-190: +idx then -idx per tile), nearest first.  Canonical order inside equal distances: insertion
    rank (2t = tile t, 2t+1 = its mirror) — kiddo's own order among equal distances is unpinned.
    Returns (item [Q, 2T] int32 signed 1-based, dist [Q, 2T] int64), blocks in row-major order (by * bw + bx)."""
    T, N = colors.shape[:2]
    base = colors.reshape(T, 3 * N)
    cand = np.empty((2 * T, 3 * N), np.int64)
    cand[0::2] = base
    cand[1::2] = mirror(base, N)
    q = queries(src, N).astype(np.int64).reshape(-1, 3 * N)
    d = np.abs(q[:, None, :] - cand[None, :, :]).sum(axis=2)
    order = np.argsort(d, axis=1, kind="stable")                      # stable: ties stay in rank order
    rank_item = np.where(np.arange(2 * T) % 2 == 0, np.arange(2 * T) // 2 + 1, -(np.arange(2 * T) // 2 + 1)).astype(np.int32)
    return rank_item[order], np.take_along_axis(d, order, axis=1)


def no_repeat_assign(colors: np.ndarray, src: np.ndarray, report_ties: bool = False):
    """render_nto1_no_repeat (rendering.rs:262-401) up to the placement of tiles: every block gets the nearest tile
    nobody nearer has taken; a tile is used once, in either orientation (`used.insert(item); used.insert(-item)`,
    :357-358).  The reference keeps the blocks in a vector sorted by the distance of their current best candidate
    (:323-326), pops the smallest, and re-inserts a block whose candidate was taken under its next candidate's
    distance (:380-391) — a merge of the per-block sorted lists in globally increasing distance.  Order among equal
    distances (its unstable sort and `insert(ix + 1)`) is unpinned; canonical here: (distance, n) with the reference's
    block number n = bx * vtiles + by (:300-301).  Blocks that run out of candidates stay unplaced (item 0; :347-351).
    Returns (item [bh,bw] int32, dist [bh,bw] uint32) (+ whether two blocks ever competed at equal distance, i.e.
    whether the canonical order was needed, if report_ties)."""
    import heapq

    T, N = colors.shape[:2]
    dim = int(np.sqrt(N))
    bh, bw = src.shape[0] // dim, src.shape[1] // dim
    if bh * bw > 2 * T:
        raise AssertionError(f"Insufficient tiles for no-repeat mode: need {bh * bw} tiles but only have {2 * T} available")
    items, dists = sorted_candidates(colors, src)
    item = np.zeros(bh * bw, np.int32)
    dist = np.zeros(bh * bw, np.uint32)
    ptr = np.zeros(bh * bw, np.int64)
    heap = []
    for by in range(bh):
        for bx in range(bw):
            q = by * bw + bx
            heap.append((int(dists[q, 0]), bx * bh + by, q))
    heapq.heapify(heap)
    used = set()
    ties = False
    while heap:
        d, n, q = heapq.heappop(heap)
        ties = ties or (bool(heap) and heap[0][0] == d)
        it = int(items[q, ptr[q]])
        if abs(it) not in used:
            used.add(abs(it))
            item[q], dist[q] = it, d
            continue
        ptr[q] += 1
        if ptr[q] < items.shape[1]:
            heapq.heappush(heap, (int(dists[q, ptr[q]]), n, q))
    if report_ties:
        return item.reshape(bh, bw), dist.reshape(bh, bw), ties
    return item.reshape(bh, bw), dist.reshape(bh, bw)


def no_repeat_assign_literal(colors: np.ndarray, src: np.ndarray):
    """The same function written the way the reference writes it — a vector of (n, remaining candidates) kept sorted by the
    distance of each block's best remaining candidate, popped from the end, re-inserted by binary search — to check
    that the heap formulation above is the same algorithm.  Only defined for inputs without equal distances between
    different (block, candidate) pairs at decision points (the reference's tie order is implementation-defined);
    small inputs only (pure Python)."""
    import bisect

    T, N = colors.shape[:2]
    dim = int(np.sqrt(N))
    bh, bw = src.shape[0] // dim, src.shape[1] // dim
    items, dists = sorted_candidates(colors, src)
    matches = []
    for n in range(bw * bh):                                         # :317-321, n -> (x = n / vtiles, y = n % vtiles)
        bx, by = n // bh, n % bh
        q = by * bw + bx
        near = [(int(dists[q, j]), int(items[q, j])) for j in range(items.shape[1])]
        near.reverse()                                               # :312 nearest.reverse(): best candidate last
        matches.append((n, near))
    matches.sort(key=lambda m: -m[1][-1][0])                         # :323-326 descending, popped from the end
    item = np.zeros((bh, bw), np.int32)
    dist = np.zeros((bh, bw), np.uint32)
    used = set()
    while matches:
        n, near = matches.pop()
        if not near:
            continue
        d, it = near.pop()
        bx, by = n // bh, n % bh
        if it not in used:
            used.add(it)
            used.add(-it)
            item[by, bx], dist[by, bx] = it, d
        else:
            if not near:
                continue
            keys = [-m[1][-1][0] for m in matches]                   # ascending view of the descending vector
            matches.insert(bisect.bisect_right(keys, -near[-1][0]), (n, near))
    return item, dist


# ---- image 0.25.2 imageops::resize(.., Lanczos3): main.rs:595, tiles/utils.rs:188-189 ------------------------
def _sinf(x: np.float32) -> np.float32:
    """f32::sin = the platform libm's sinf (numpy's own float32 sin is a different implementation)."""
    import ctypes
    import ctypes.util
    global _LIBM
    try:
        _LIBM
    except NameError:
        _LIBM = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
        _LIBM.sinf.restype = ctypes.c_float
        _LIBM.sinf.argtypes = [ctypes.c_float]
    return np.float32(_LIBM.sinf(float(x)))


def _lanczos3(x: np.float32) -> np.float32:
    f = np.float32

    def sinc(t):
        a = f(t * f(np.pi))
        return f(1.0) if t == 0 else f(_sinf(a) / a)

    return f(sinc(x) * sinc(f(x / f(3.0)))) if abs(x) < 3.0 else f(0.0)


def resize_axis(n_in: int, n_out: int):
    """sample.rs horizontal_sample / vertical_sample tap set-up: per output index (left, normalised f32 weights)."""
    f = np.float32
    ratio = f(f(n_in) / f(n_out))
    sratio = f(1.0) if ratio < 1 else ratio
    support = f(f(3.0) * sratio)
    taps = []
    for o in range(n_out):
        c = f(f(f(o) + f(0.5)) * ratio)
        left = min(max(int(np.floor(f(c - support))), 0), n_in - 1)
        right = min(max(int(np.ceil(f(c + support))), left + 1), n_in)
        c = f(c - f(0.5))
        ws = [_lanczos3(f(f(f(i) - c) / sratio)) for i in range(left, right)]
        s = f(0.0)
        for w in ws:
            s = f(s + w)
        taps.append((left, np.array([f(w / s) for w in ws], np.float32)))
    return taps


def resize_lanczos3(img: np.ndarray, nw: int, nh: int) -> np.ndarray:
    """resize(): vertical_sample into f32, then horizontal_sample with clamp and round (half away from zero)."""
    h, w = img.shape[:2]
    if (nw, nh) == (w, h):
        return img.copy()
    tmp = np.zeros((nh, w, 3), np.float32)
    for oy, (left, ws) in enumerate(resize_axis(h, nh)):
        t = np.zeros((w, 3), np.float32)
        for i, wt in enumerate(ws):
            t = t + img[left + i].astype(np.float32) * wt  # two separately rounded f32 operations
        tmp[oy] = t
    out = np.zeros((nh, nw, 3), np.uint8)
    for ox, (left, ws) in enumerate(resize_axis(w, nw)):
        t = np.zeros((nh, 3), np.float32)
        for i, wt in enumerate(ws):
            t = t + tmp[:, left + i] * wt
        t = np.clip(t, np.float32(0), np.float32(255))
        out[:, ox] = np.where(t - np.floor(t) >= 0.5, np.floor(t) + 1, np.floor(t)).astype(np.uint8)  # t >= 0: round half up
    return out


def prepare_view(img: np.ndarray, tile_size: int, crop: bool):
    """tiles/utils.rs:93-186 with the smallest value winning a tie of most_common_value."""
    h, w = img.shape[:2]
    if w < tile_size or h < tile_size:
        raise ValueError("DimensionError")
    nonwhite = ~((img[..., 0] > 240) & (img[..., 1] > 240) & (img[..., 2] > 240))

    def first_last(m):  # m [lines, n]: per line first non-white (n if none) and last non-white (0 if none)
        n = m.shape[1]
        anyv = m.any(axis=1)
        first = np.where(anyv, m.argmax(axis=1), n)
        last = np.where(anyv, n - 1 - m[:, ::-1].argmax(axis=1), 0)
        return first, last

    def mode(v, skip):
        v = v[v != skip]
        if v.size == 0:
            return 0
        vals, counts = np.unique(v, return_counts=True)
        return int(vals[counts.argmax()])  # np.unique sorts: the first maximum is the smallest value

    fl, fr = first_last(nonwhite)
    ft, fb = first_last(nonwhite.T)
    c0, c1, r0, r1 = mode(fl, w), mode(fr, 0), mode(ft, h), mode(fb, 0)
    assert c0 < c1 and r0 < r1
    vw, vh, vx, vy = c1 - c0, r1 - r0, c0, r0
    if crop:
        size = min(vw, vh)
        vx += (vw - size) // 2
        vy += (vh - size) // 2
        vw = vh = size
    return vx, vy, vw, vh
