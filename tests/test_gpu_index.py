"""The 1to1 search index (colour-cube table, index.cu) against the oracle and against the scan kernel.

The index is the GPU stand-in for the KD-tree of tileset.rs:178-190; its answers must be the scan's answers
bit for bit (minimum L1 distance, ties to the smallest tile index)."""
import numpy as np
import pytest

import emosaic_b200 as emo
import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture()
def ictx(ctx):
    yield ctx
    ctx.set_match_mode("auto")


def both(ctx, src):
    ctx.set_match_mode("scan")
    a = ctx.match(src)
    ctx.set_match_mode("index")
    b = ctx.match(src)
    return a, b


@pytest.mark.parametrize("T,H,W", [(1, 7, 9), (2, 16, 16), (7, 33, 35), (300, 100, 100), (5000, 64, 257), (70_000, 40, 40)])
def test_index_matches_oracle(ictx, T, H, W):
    rng = np.random.default_rng(T)
    colors = rng.integers(0, 256, (T, 1, 3), dtype=np.uint8)
    src = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    src[0, 0] = 0
    src[-1, -1] = 255
    ictx.set_library(colors)
    (si, sd), (ii, id_) = both(ictx, src)
    ri, rd = oracle.match(colors, src) if T * H * W <= 3e9 else oracle.KdTree(colors).match(src)
    assert (ii == ri).all() and (id_ == rd).all()
    assert (si == ri).all() and (sd == rd).all()


def test_index_every_colour_of_the_cube(ictx):
    """All 2^24 queries at once: the index output IS the table, checked against the scan kernel cell by cell and
    against the oracle on a sample."""
    rng = np.random.default_rng(5)
    colors = rng.integers(0, 256, (3000, 1, 3), dtype=np.uint8)
    colors[:8, 0] = [[0, 0, 0], [255, 255, 255], [255, 0, 0], [0, 255, 0], [0, 0, 255], [0, 0, 0], [255, 255, 255], [7, 7, 7]]
    v = np.arange(1 << 24, dtype=np.uint32)
    src = np.stack([v & 255, (v >> 8) & 255, v >> 16], -1).astype(np.uint8).reshape(4096, 4096, 3)
    ictx.set_library(colors)
    (si, sd), (ii, id_) = both(ictx, src)
    assert (si == ii).all() and (sd == id_).all()
    rows = rng.choice(4096, 6, replace=False)
    for r in rows:
        ri, rd = oracle.match(colors, src[r:r + 1])
        assert (ii[r:r + 1] == ri).all() and (id_[r:r + 1] == rd).all()
    assert ii[0, 0] == 1 and ii[-1, -1] == 2  # duplicates (tiles 6 and 7) lose to the smaller index


def test_index_ties(ictx):
    # quantised library: many equidistant tiles at non-zero distance, duplicates of every colour
    rng = np.random.default_rng(11)
    colors = (rng.integers(0, 8, (5000, 1, 3)) * 32).astype(np.uint8)
    src = (rng.integers(0, 8, (96, 96, 3)) * 32 + 16).astype(np.uint8)  # exactly half way between library colours
    ictx.set_library(colors)
    (si, sd), (ii, id_) = both(ictx, src)
    ri, rd = oracle.match(colors, src)
    assert (ii == ri).all() and (id_ == rd).all() and (si == ri).all()
    # one axis only: the library lives on the grey diagonal
    g = rng.integers(0, 256, (400, 1, 1)).astype(np.uint8)
    colors = np.repeat(g, 3, axis=2)
    src = rng.integers(0, 256, (50, 60, 3), dtype=np.uint8)
    ictx.set_library(colors)
    (si, sd), (ii, id_) = both(ictx, src)
    ri, rd = oracle.match(colors, src)
    assert (ii == ri).all() and (id_ == rd).all() and (si == ri).all()


def test_index_follows_the_library(ictx):
    rng = np.random.default_rng(3)
    src = rng.integers(0, 256, (31, 45, 3), dtype=np.uint8)
    a = rng.integers(0, 256, (500, 1, 3), dtype=np.uint8)
    b = rng.integers(0, 256, (900, 1, 3), dtype=np.uint8)
    ictx.set_match_mode("index")
    ictx.set_library(a)
    ia, _ = ictx.match(src)
    ictx.set_library(b)       # drops the index of library a
    ib, db = ictx.match(src)
    ri, rd = oracle.match(b, src)
    assert (ib == ri).all() and (db == rd).all()
    assert (ia == oracle.match(a, src)[0]).all()
    ictx.build_index()        # explicit rebuild gives the same table
    ib2, db2 = ictx.match(src)
    assert (ib2 == ib).all() and (db2 == db).all()


def test_index_auto_rule_and_errors(ictx):
    rng = np.random.default_rng(4)
    colors = rng.integers(0, 256, (2000, 1, 3), dtype=np.uint8)
    src = rng.integers(0, 256, (64, 64, 3), dtype=np.uint8)
    ictx.set_match_mode("auto")
    ictx.set_library(colors)
    n0 = ictx.launch_count()
    ictx.match(src)                       # 4096 x 2000 pairs: scanned, no index build
    assert ictx.launch_count() - n0 <= 3  # the scan kernel (+ key init / finalize when the candidates are split)
    big = rng.integers(0, 256, (2100, 2100, 3), dtype=np.uint8)   # 4.4 M x 2000 >= 2^31: index built, then used
    n0 = ictx.launch_count()
    bi, bd = ictx.match(big)
    assert ictx.launch_count() - n0 == 6  # seed + 3 sweeps + compaction to the 32 MiB form + lookup
    n0 = ictx.launch_count()
    i2, d2 = ictx.match(src)              # the index exists now: lookup
    assert ictx.launch_count() - n0 == 1
    ri, rd = oracle.match(colors, src)
    assert (i2 == ri).all() and (d2 == rd).all()
    sel = big[:3]
    ri, rd = oracle.match(colors, sel)
    assert (bi[:3] == ri).all() and (bd[:3] == rd).all()
    # 4to1 has no colour-cube index
    c4 = rng.integers(0, 256, (50, 4, 3), dtype=np.uint8)
    ictx.set_library(c4)
    with pytest.raises(emo.EmosaicError, match="N == 1"):
        ictx.build_index()
    ictx.set_match_mode("index")          # no index for this library: the scan answers
    q = rng.integers(0, 256, (8, 8, 3), dtype=np.uint8)
    it, ds = ictx.match(q)
    ri, rd = oracle.match(c4, q)
    assert (it == ri).all() and (ds == rd).all()
    with pytest.raises(emo.EmosaicError, match="unknown mode"):
        ictx.set_match_mode(7)
    c2 = emo.Context(0)
    with pytest.raises(emo.EmosaicError, match="no library"):
        c2.build_index()
    c2.close()


def test_index_unaligned_and_ragged(ictx):
    """Odd source address, item/dist at 4-byte (not 16-byte) alignment, Q not a multiple of 4."""
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(9)
    colors = rng.integers(0, 256, (777, 1, 3), dtype=np.uint8)
    ictx.set_library(colors)
    ictx.set_match_mode("index")
    for (H, W, so, oo) in [(5, 7, 1, 4), (9, 13, 2, 0), (3, 341, 3, 8), (4, 64, 0, 12), (1, 1, 1, 4), (2, 3, 0, 0)]:
        src_h = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        raw = torch.zeros(H * W * 3 + 16, dtype=torch.uint8, device=dev)
        raw[so:so + H * W * 3] = torch.from_numpy(src_h.reshape(-1)).to(dev)
        item = torch.full((H * W + 16,), -77, dtype=torch.int32, device=dev)
        dist = torch.full((H * W + 16,), 4242, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        ictx.match_dev(raw.data_ptr() + so, W, H, item.data_ptr() + oo, dist.data_ptr() + oo)
        ictx.sync()
        o = oo // 4
        ri, rd = oracle.match(colors, src_h)
        ih, dh = item.cpu().numpy(), dist.cpu().numpy()
        assert (ih[o:o + H * W] == ri.reshape(-1)).all() and (dh[o:o + H * W] == rd.reshape(-1)).all()
        assert (ih[:o] == -77).all() and (ih[o + H * W:] == -77).all()      # nothing written outside [0, Q)
        assert (dh[:o] == 4242).all() and (dh[o + H * W:] == 4242).all()


def test_config4_index_equals_scan_full_size(ictx):
    """C4 at full size (100 000 tiles, 4096 x 4096 blocks): the table lookup and the 1.7e12-pair scan agree on every block."""
    rng = np.random.default_rng(1234)
    colors = rng.integers(0, 256, (100_000, 1, 3), dtype=np.uint8)
    src = np.random.default_rng(5678).integers(0, 256, (4096, 4096, 3), dtype=np.uint8)
    ictx.set_library(colors)
    (si, sd), (ii, id_) = both(ictx, src)
    assert (si == ii).all() and (sd == id_).all()
    # clustered library (what averaging random tiles gives: colours within a few units of mid-grey)
    colors = np.clip(rng.normal(127, 4, (100_000, 1, 3)), 0, 255).astype(np.uint8)
    ictx.set_library(colors)
    (si, sd), (ii, id_) = both(ictx, src[:1024])
    assert (si == ii).all() and (sd == id_).all()


@pytest.mark.parametrize("ts,bh,bw", [(8, 5, 128), (8, 33, 64), (16, 7, 96), (16, 3, 32), (8, 4, 100), (32, 3, 16), (6, 5, 64)])
def test_mosaic_dev(ictx, ts, bh, bw):
    """emo_mosaic_dev (device-resident match + compose in one call) with the index, with tint and with the scan; canaries
    around all three outputs."""
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(ts * 1000 + bw)
    T = 500
    tiles_h = rng.integers(0, 256, (T, ts, ts, 3), dtype=np.uint8)
    colors = ictx.analyse_tiles(tiles_h, 1)
    src_h = rng.integers(0, 256, (bh, bw, 3), dtype=np.uint8)
    ictx.set_library(colors, tiles_h)
    ictx.set_match_mode("index")
    ictx.build_index()
    Q, OB = bh * bw, bh * ts * bw * ts * 3
    src = torch.from_numpy(src_h.reshape(-1)).to(dev)
    item = torch.full((Q + 32,), -77, dtype=torch.int32, device=dev)
    dist = torch.full((Q + 32,), 4242, dtype=torch.int32, device=dev)
    out = torch.full((OB + 64,), 0xA5, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    n0 = ictx.launch_count()
    ictx.mosaic_dev(src.data_ptr(), bw, bh, 3, 0, item.data_ptr() + 64, dist.data_ptr() + 64, out.data_ptr() + 32)
    ictx.sync()
    assert ictx.launch_count() - n0 == 2
    ri, rd = oracle.match(colors, src_h)
    ih, dh, oh = item.cpu().numpy(), dist.cpu().numpy(), out.cpu().numpy()
    assert (ih[16:16 + Q] == ri.reshape(-1)).all() and (dh[16:16 + Q] == rd.reshape(-1)).all()
    assert (oh[32:32 + OB].reshape(bh * ts, bw * ts, 3) == oracle.render(tiles_h, ri)).all()
    assert (ih[:16] == -77).all() and (ih[16 + Q:] == -77).all() and (dh[:16] == 4242).all() and (dh[16 + Q:] == 4242).all()
    assert (oh[:32] == 0xA5).all() and (oh[32 + OB:] == 0xA5).all()
    # tint (RGBA) and the scan mode go through the two-kernel path of the same call
    out4 = torch.zeros(bh * ts * bw * ts * 4, dtype=torch.uint8, device=dev)
    ictx.mosaic_dev(src.data_ptr(), bw, bh, 4, 127, item.data_ptr(), dist.data_ptr(), out4.data_ptr())
    ictx.sync()
    assert (out4.cpu().numpy().reshape(bh * ts, bw * ts, 4) == oracle.tint(oracle.render(tiles_h, ri), src_h, 127)).all()
    ictx.set_match_mode("scan")
    out.fill_(0)
    ictx.mosaic_dev(src.data_ptr(), bw, bh, 3, 0, item.data_ptr(), dist.data_ptr(), out.data_ptr())
    ictx.sync()
    assert (item.cpu().numpy()[:Q] == ri.reshape(-1)).all()
    assert (out.cpu().numpy()[:OB].reshape(bh * ts, bw * ts, 3) == oracle.render(tiles_h, ri)).all()


def test_mosaic_dev_repeated_calls_and_state_changes(ictx):
    """emo_mosaic_dev called over and over on the same buffers (the steady state of a frame loop, and of bench.py): every call
    must see the new buffer contents, and library, match mode, tint and geometry may change between calls.  Every call is checked
    against the oracle."""
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(909)
    T, ts, bh, bw = 700, 8, 24, 96
    Q, OB = bh * bw, bh * ts * bw * ts * 3

    def library():
        tiles = rng.integers(0, 256, (T, ts, ts, 3), dtype=np.uint8)
        colors = ictx.analyse_tiles(tiles, 1)
        ictx.set_library(colors, tiles)
        return tiles, colors

    tiles_h, colors = library()
    ictx.set_match_mode("index")
    ictx.build_index()
    src = torch.zeros(Q * 3, dtype=torch.uint8, device=dev)
    item = torch.zeros(Q, dtype=torch.int32, device=dev)
    dist = torch.zeros(Q, dtype=torch.int32, device=dev)
    out = torch.zeros(OB, dtype=torch.uint8, device=dev)
    out4 = torch.zeros(bh * ts * bw * ts * 4, dtype=torch.uint8, device=dev)

    def step(oc=3, alpha=0, rows=bh):
        src_h = rng.integers(0, 256, (bh, bw, 3), dtype=np.uint8)
        src.copy_(torch.from_numpy(src_h.reshape(-1)).to(dev))
        item.fill_(0); dist.fill_(0); out.fill_(0); out4.fill_(0)
        torch.cuda.synchronize()
        n0 = ictx.launch_count()
        ictx.mosaic_dev(src.data_ptr(), bw, rows, oc, alpha, item.data_ptr(), dist.data_ptr(), (out if oc == 3 else out4).data_ptr())
        ictx.sync()
        ri, rd = oracle.match(colors, src_h[:rows])
        q = rows * bw
        assert (item.cpu().numpy()[:q] == ri.reshape(-1)).all() and (dist.cpu().numpy()[:q] == rd.reshape(-1)).all()
        want = oracle.render(tiles_h, ri)
        if oc == 3:
            assert (out.cpu().numpy()[:q * ts * ts * 3].reshape(rows * ts, bw * ts, 3) == want).all()
        else:
            assert (out4.cpu().numpy()[:q * ts * ts * 4].reshape(rows * ts, bw * ts, 4) == oracle.tint(want, src_h[:rows], alpha)).all()
        return ictx.launch_count() - n0

    assert [step() for _ in range(5)] == [2] * 5          # new pixels every time
    tiles_h, colors = library()                           # same shape, new tiles: the index is rebuilt
    for _ in range(4):
        step()
    for _ in range(3):
        step(rows=bh - 5)                                  # another geometry ...
    for _ in range(3):
        step()                                             # ... and back
    for _ in range(3):
        step(oc=4, alpha=127)                              # tint tables are built on the first of these
    for _ in range(3):
        step(oc=4, alpha=60)
    ictx.set_match_mode("scan")
    for _ in range(4):
        step()
    ictx.set_match_mode("index")
    for _ in range(3):
        step()


def test_mosaic_dev_4to1_and_errors(ictx):
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(21)
    tiles_h = rng.integers(0, 256, (120, 16, 16, 3), dtype=np.uint8)
    colors = ictx.analyse_tiles(tiles_h, 2)
    src_h = rng.integers(0, 256, (12, 64, 3), dtype=np.uint8)
    ictx.set_library(colors, tiles_h)
    src = torch.from_numpy(src_h.reshape(-1)).to(dev)
    item = torch.zeros(6 * 32, dtype=torch.int32, device=dev)
    dist = torch.zeros(6 * 32, dtype=torch.int32, device=dev)
    out = torch.zeros(6 * 16 * 32 * 16 * 3, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    ictx.mosaic_dev(src.data_ptr(), 64, 12, 3, 0, item.data_ptr(), dist.data_ptr(), out.data_ptr())
    ictx.sync()
    ri, rd = oracle.match(colors, src_h)
    assert (item.cpu().numpy().reshape(6, 32) == ri).all() and (dist.cpu().numpy().reshape(6, 32) == rd).all()
    assert (out.cpu().numpy().reshape(96, 512, 3) == oracle.render(tiles_h, ri)).all()
    with pytest.raises(emo.EmosaicError, match="item/dist is NULL"):
        ictx.mosaic_dev(src.data_ptr(), 64, 12, 3, 0, 0, dist.data_ptr(), out.data_ptr())
    with pytest.raises(emo.EmosaicError, match="divisible by 2"):
        ictx.mosaic_dev(src.data_ptr(), 63, 12, 3, 0, item.data_ptr(), dist.data_ptr(), out.data_ptr())
    ictx.set_library(colors)   # no tile pixels
    with pytest.raises(emo.EmosaicError, match="no tile pixels"):
        ictx.mosaic_dev(src.data_ptr(), 64, 12, 3, 0, item.data_ptr(), dist.data_ptr(), out.data_ptr())


def test_index_at_the_tile_limit(ictx):
    """The key holds 22 bits of tile index: T = 2^22 is the largest indexed library, T = 2^22 + 1 is scanned."""
    rng = np.random.default_rng(22)
    T = 1 << 22
    colors = rng.integers(0, 256, (T + 1, 1, 3), dtype=np.uint8)
    colors[T - 1, 0] = [1, 2, 250]        # make the last indexable tile the unique answer for one query
    same = (colors[:T - 1, 0] == colors[T - 1, 0]).all(1)
    colors[:T - 1][same] = [[9, 9, 9]]
    src = rng.integers(0, 256, (24, 24, 3), dtype=np.uint8)
    src[0, 0] = [1, 2, 250]
    ri, rd = oracle.match(colors[:T], src)
    ictx.set_match_mode("index")
    ictx.set_library(colors[:T])
    ii, id_ = ictx.match(src)
    assert (ii == ri).all() and (id_ == rd).all()
    assert ii[0, 0] == T and id_[0, 0] == 0
    ictx.set_library(colors)              # one tile more: no index, the scan answers (also in "index" mode)
    with pytest.raises(emo.EmosaicError, match="T <= 2\\^22"):
        ictx.build_index()
    si, sd = ictx.match(src)
    ri2, rd2 = oracle.match(colors, src)
    assert (si == ri2).all() and (sd == rd2).all()


def test_index_random_sweep(ictx):
    """40 random (library, source) pairs — uniform, clustered, few-colour and duplicated libraries, ragged source sizes:
    index == scan on every block, and == the oracle on every third case."""
    rng = np.random.default_rng(2024)
    for case in range(40):
        T = int(rng.choice([1, 2, 3, 17, 255, 256, 257, 1000, 4999]))
        kind = case % 4
        if kind == 0:
            colors = rng.integers(0, 256, (T, 1, 3), dtype=np.uint8)
        elif kind == 1:
            colors = np.clip(rng.normal(127, 6, (T, 1, 3)), 0, 255).astype(np.uint8)
        elif kind == 2:
            colors = (rng.integers(0, 3, (T, 1, 3)) * 127).astype(np.uint8)          # at most 27 distinct colours
        else:
            base = rng.integers(0, 256, (max(1, T // 3), 1, 3), dtype=np.uint8)
            colors = base[rng.integers(0, base.shape[0], T)]                         # every colour ~3 times
        H, W = int(rng.integers(1, 200)), int(rng.integers(1, 300))
        src = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        ictx.set_library(colors)
        (si, sd), (ii, id_) = both(ictx, src)
        assert (si == ii).all() and (sd == id_).all(), f"case {case}: T={T} kind={kind} {H}x{W}"
        if case % 3 == 0:
            ri, rd = oracle.match(colors, src)
            assert (ii == ri).all() and (id_ == rd).all(), f"case {case}"


# ---- the compact form of the index (u16 winner slot per colour; index.cu) ------------------------------------------
def _palette_library(T, P, seed):
    """T tiles whose colours come from a palette of P distinct colours (duplicates: the smallest index must win)."""
    rng = np.random.default_rng(seed)
    pal = rng.choice(1 << 24, P, replace=False).astype(np.uint32)
    pick = pal[rng.integers(0, P, T)]
    return np.stack([pick & 255, (pick >> 8) & 255, pick >> 16], -1).astype(np.uint8).reshape(T, 1, 3)


@pytest.mark.parametrize("T,P,H,W,why", [
    (300, 0, 96, 128, "slot = tile, small launch"),
    (300, 0, 1024, 4096, "slot = tile, 4 M pixels"),
    (40_000, 0, 64, 4096, "slot = tile, 160 KB of colours"),
    (60_000, 0, 128, 1024, "slot = tile, 240 KB of colours"),
    (65_536, 0, 64, 512, "largest library whose tile index fits a slot"),
    (100_000, 40_000, 256, 4096, "more tiles than slots: compacted to the distinct winners"),
    (70_000, 0, 64, 512, "more than 65 536 distinct colours: falls back to the 64 MiB table"),
])
def test_compact_index_equals_wide_index_and_scan(ictx, T, P, H, W, why):
    colors = _palette_library(T, P, T) if P else np.random.default_rng(T).integers(0, 256, (T, 1, 3), dtype=np.uint8)
    src = np.random.default_rng(T + 1).integers(0, 256, (H, W, 3), dtype=np.uint8)
    src[0, :4] = [[0, 0, 0], [255, 255, 255], colors[-1, 0], colors[0, 0]]
    ictx.set_library(colors)
    res = {}
    for mode in ("index_wide", "index_compact", "auto"):
        ictx.set_match_mode(mode)
        res[mode] = ictx.match(src)
    wi, wd = res["index_wide"]
    for mode in ("index_compact", "auto"):
        assert (res[mode][0] == wi).all() and (res[mode][1] == wd).all(), f"{mode} differs from the 64 MiB table: {why}"
    rows = slice(0, min(H, 24))
    ri, rd = oracle.KdTree(colors).match(src[rows])
    assert (wi[rows] == ri).all() and (wd[rows] == rd).all()
    # dist is the L1 distance to the chosen tile's colour everywhere
    ch = colors[wi - 1, 0].astype(np.int64)
    assert (np.abs(ch - src.astype(np.int64)).sum(-1) == wd).all()


def test_compact_index_ragged_and_unaligned(ictx):
    """Pixel counts that are not multiples of 4 (the tail goes through the scalar path) and device buffers at odd
    addresses (the compact kernel needs aligned words: those launches take the 64 MiB table) give the same maps."""
    import ctypes as C
    colors = np.random.default_rng(3).integers(0, 256, (5000, 1, 3), dtype=np.uint8)
    ictx.set_library(colors)
    for H, W in ((1, 1), (1, 3), (3, 5), (7, 11), (31, 33)):
        src = np.random.default_rng(H * W).integers(0, 256, (H, W, 3), dtype=np.uint8)
        ictx.set_match_mode("index_compact")
        ci, cd = ictx.match(src)
        ictx.set_match_mode("scan")
        si, sd = ictx.match(src)
        assert (ci == si).all() and (cd == sd).all()
    # unaligned device pointers through the *_dev call
    H, W = 9, 13
    src = np.random.default_rng(99).integers(0, 256, (H, W, 3), dtype=np.uint8)
    ictx.set_match_mode("scan")
    si, sd = ictx.match(src)
    ictx.set_match_mode("index_compact")
    Q = H * W
    d_src = ictx.dev_alloc(Q * 3 + 64); d_item = ictx.dev_alloc(Q * 4 + 64); d_dist = ictx.dev_alloc(Q * 4 + 64)
    try:
        for so, io in ((1, 4), (0, 4), (3, 0), (0, 0)):
            ictx.h2d(d_src + so, src)
            ictx.match_dev(d_src + so, W, H, d_item + io, d_dist + io)
            gi = np.zeros((H, W), np.int32); gd = np.zeros((H, W), np.uint32)
            ictx.d2h(gi, d_item + io); ictx.d2h(gd, d_dist + io); ictx.sync()
            assert (gi == si).all() and (gd == sd).all()
    finally:
        for p in (d_src, d_item, d_dist):
            ictx.dev_free(p)
