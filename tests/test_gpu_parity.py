"""Parity of the CUDA path against the CPU oracle, through the C ABI.  Needs a B200 (-m gpu).

Bit-exact everywhere: the path is integer/byte arithmetic, and the f32 tint blend is reproduced
exactly (tolerance 0)."""
import hashlib
import os

import numpy as np
import pytest

import emosaic_b200 as emo
import oracle
from oracle import oracle_np as onp

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ---- analysis ------------------------------------------------------------------------------------
def test_reference_kat_average(ctx):
    # color.rs:49-64 through the kernel: a 2x2 tile analysed with N=1
    img = np.array([[[100, 150, 200], [200, 100, 50]], [[50, 200, 100], [150, 50, 150]]], np.uint8)
    assert emo.analyse(img, 1, ctx).tolist() == [[125, 125, 125]]
    # analysis.rs:44-55
    red = np.zeros((2, 2, 3), np.uint8)
    red[..., 0] = 255
    assert emo.analyse(red, 4, ctx).tolist() == [[255, 0, 0]] * 4


@pytest.mark.parametrize("ts,dim,T", [(64, 1, 301), (64, 2, 301), (32, 1, 1000), (32, 2, 1000), (16, 1, 77), (16, 2, 77),
                                      (8, 1, 50), (8, 2, 50), (12, 3, 40), (16, 4, 33), (10, 3, 21), (7, 2, 19), (5, 5, 9),
                                      (1, 1, 5), (128, 2, 6)])
def test_analyse_parity(ctx, ts, dim, T):
    rng = np.random.default_rng(ts * 100 + dim)
    tiles = rng.integers(0, 256, (T, ts, ts, 3), dtype=np.uint8)
    tiles[0] = 255  # saturated sums
    tiles[1] = 0
    got = ctx.analyse_tiles(tiles, dim)
    assert (got == oracle.analyse_tiles(tiles, dim * dim)).all()


@pytest.mark.parametrize("ts,T", [(64, 4099), (32, 5000), (16, 100), (6, 10)])
def test_analyse_fused_parity(ctx, ts, T):
    rng = np.random.default_rng(ts)
    tiles = rng.integers(0, 256, (T, ts, ts, 3), dtype=np.uint8)
    o1, o4 = ctx.analyse_tiles_fused(tiles)
    assert (o1 == oracle.analyse_tiles(tiles, 1)).all()
    assert (o4 == oracle.analyse_tiles(tiles, 4)).all()


def test_analyse_errors(ctx):
    tiles = np.zeros((2, 4, 4, 3), np.uint8)
    with pytest.raises(emo.EmosaicError, match="Rectangle dimensions must be positive"):
        ctx.analyse_tiles(tiles, 5)  # floor(4/5) == 0 -> color.rs:18 panic
    assert ctx.analyse_tiles(np.zeros((0, 8, 8, 3), np.uint8), 2).shape == (0, 4, 3)  # empty library
    with pytest.raises(emo.EmosaicError, match="divisible by 2"):
        ctx.analyse_tiles_fused(np.zeros((2, 5, 5, 3), np.uint8))


def test_analyse_full_size_properties(ctx):
    """C3 geometry at a size the box handles in seconds: 200k tiles of 64x64 generated on the host in
    slabs; property: the 1to1 mean is bounded by the 4to1 quadrant means and both equal the oracle on a
    sampled subset."""
    T = 200_000
    rng = np.random.default_rng(1234)
    tiles = rng.integers(0, 256, (T, 64, 64, 3), dtype=np.uint8)
    o1, o4 = ctx.analyse_tiles_fused(tiles)
    sel = rng.choice(T, 4096, replace=False)
    assert (o1[sel] == oracle.analyse_tiles(tiles[sel], 1)).all()
    assert (o4[sel] == oracle.analyse_tiles(tiles[sel], 4)).all()
    q = o4.astype(np.int64)
    lo, hi = q.min(1), q.max(1)
    assert ((o1[:, 0] >= lo) & (o1[:, 0] <= hi)).all()
    # exact identity: floor(total/4096) with total = sum of quadrant sums; check via uint64 numpy on all tiles
    tot = tiles.reshape(T, -1, 3).sum(1, dtype=np.uint64) // np.uint64(4096)
    assert (o1[:, 0] == tot.astype(np.uint8)).all()


def test_config3_full_size_device_api(ctx):
    """C3 at full size through the device-pointer ABI (the way bench.py drives it): 1 000 000 tiles of 64x64
    (12.29 GB) generated in HBM with torch, fused 1to1 + 4to1 analysis, oracle parity on 3 000 sampled tiles and
    the exact identity mean1 = floor(sum of the four quadrant sums / 4096) checked against separate dim=1 / dim=2 runs."""
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda", 0)
    T = 1_000_000
    g = torch.Generator(device=dev)
    g.manual_seed(1234)
    tiles = torch.randint(0, 256, (T, 64, 64, 3), dtype=torch.uint8, device=dev, generator=g)
    tiles[0] = 255
    tiles[1] = 0
    o1 = torch.empty((T, 1, 3), dtype=torch.uint8, device=dev)
    o4 = torch.empty((T, 4, 3), dtype=torch.uint8, device=dev)
    s1 = torch.empty_like(o1)
    s4 = torch.empty_like(o4)
    torch.cuda.synchronize()
    ctx.analyse_fused_dev(tiles.data_ptr(), T, 64, o1.data_ptr(), o4.data_ptr())
    ctx.analyse_dev(tiles.data_ptr(), T, 64, 1, s1.data_ptr())
    ctx.analyse_dev(tiles.data_ptr(), T, 64, 2, s4.data_ptr())
    ctx.sync()
    assert bool((o1 == s1).all()) and bool((o4 == s4).all())
    sel = torch.cat([torch.arange(0, 1000, device=dev), torch.randint(0, T, (2000,), device=dev, generator=g)])
    th = tiles[sel].cpu().numpy()
    assert (o1[sel].cpu().numpy() == oracle.analyse_tiles(th, 1)).all()
    assert (o4[sel].cpu().numpy() == oracle.analyse_tiles(th, 4)).all()
    assert o1[0].tolist() == [[255, 255, 255]] and o4[1].tolist() == [[0, 0, 0]] * 4
    del tiles


# ---- match ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,T,H,W", [(1, 300, 100, 100), (1, 4096, 100, 100), (1, 1, 3, 5), (1, 17, 1, 1), (4, 500, 64, 96),
                                     (4, 10000, 128, 128), (9, 200, 30, 42), (16, 100, 16, 24), (1, 33000, 64, 64),
                                     (4, 20, 2, 2), (1, 2500, 700, 900)])
def test_match_parity(ctx, N, T, H, W):
    rng = np.random.default_rng(N * 1000 + T)
    colors = rng.integers(0, 256, (T, N, 3), dtype=np.uint8)
    src = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    ctx.set_library(colors)
    item, dist = ctx.match(src)
    if T * H * W <= 3e9:
        ri, rd = oracle.match(colors, src)
    else:
        ri, rd = oracle.KdTree(colors).match(src)
    assert (dist == rd).all()
    assert (item == ri).all()


@pytest.mark.parametrize("N,T,bh,bw", [(25, 300, 9, 11), (36, 150, 7, 5), (64, 500, 12, 9), (256, 200, 5, 6), (1024, 70, 3, 4),
                                       (4096, 40, 2, 2), (16384, 12, 1, 2), (25, 1, 1, 1), (49, 129, 65, 3)])
def test_match_wide_modes(ctx, N, T, bh, bw):
    """--mode 5 ... 128 (main.rs:403-413): vectors of up to 49 152 bytes through match_wide_kernel."""
    dim = int(N ** 0.5)
    rng = np.random.default_rng(N + T)
    colors = rng.integers(0, 256, (T, N, 3), dtype=np.uint8)
    colors[T // 2] = colors[0]  # duplicate -> tie, smallest idx must win
    src = rng.integers(0, 256, (bh * dim, bw * dim, 3), dtype=np.uint8)
    # plant exact and mirrored matches
    src[:dim, :dim] = colors[T - 1].reshape(dim, dim, 3)
    if bw > 1:
        src[:dim, dim:2 * dim] = colors[0].reshape(dim, dim, 3)[:, ::-1]
    ctx.set_library(colors)
    item, dist = ctx.match(src)
    ri, rd = oracle.match(colors, src)
    assert (dist == rd).all() and (item == ri).all()
    assert dist[0, 0] == 0


def test_wide_mode_end_to_end(ctx):
    """mode 8 (N = 64), ts 16: analyse -> match -> compose + tint against the oracle."""
    rng = np.random.default_rng(8)
    tiles = rng.integers(0, 256, (300, 16, 16, 3), dtype=np.uint8)
    src = rng.integers(0, 256, (5 * 8, 7 * 8, 3), dtype=np.uint8)
    colors = ctx.analyse_tiles(tiles, 8)
    assert (colors == oracle.analyse_tiles(tiles, 64)).all()
    ctx.set_library(colors, tiles)
    out, item, dist = ctx.mosaic(src, 4, 100)
    ri, rd = oracle.match(colors, src)
    assert (item == ri).all() and (dist == rd).all()
    assert (out == oracle.tint(oracle.render(tiles, ri), src, 100)).all()


def test_match_ties_and_duplicates(ctx):
    # heavy exact ties: 8-colour palette, many duplicate tiles; the canonical winner is the smallest idx
    g = np.load(os.path.join(GOLD, "ties_palette.npz"))
    ctx.set_library(g["colors"])
    item, dist = ctx.match(g["src"])
    assert (item == g["item"]).all() and (dist == g["dist"]).all() and (dist == 0).all()
    # all tiles identical: always +1
    colors = np.full((1000, 4, 3), 77, np.uint8)
    ctx.set_library(colors)
    item, dist = ctx.match(np.random.default_rng(0).integers(0, 256, (32, 32, 3), dtype=np.uint8))
    assert (item == 1).all()
    # quantised colours (multiples of 32): ties between distinct tiles at non-zero distance
    rng = np.random.default_rng(11)
    colors = (rng.integers(0, 8, (5000, 1, 3)) * 32).astype(np.uint8)
    src = (rng.integers(0, 8, (96, 96, 3)) * 32 + 5).astype(np.uint8)
    ctx.set_library(colors)
    item, dist = ctx.match(src)
    ri, rd = oracle.match(colors, src)
    assert (item == ri).all() and (dist == rd).all()


def test_match_mirror_sign(ctx):
    asym = np.array([[[1, 1, 1], [200, 200, 200], [1, 1, 1], [200, 200, 200]]], np.uint8)
    q = np.array([[[200, 200, 200], [1, 1, 1]], [[200, 200, 200], [1, 1, 1]]], np.uint8)
    ctx.set_library(asym)
    item, dist = ctx.match(q)
    assert item.tolist() == [[-1]] and dist.tolist() == [[0]]
    item, dist = ctx.match(q[:, ::-1].copy())
    assert item.tolist() == [[1]] and dist.tolist() == [[0]]


def test_match_errors(ctx):
    ctx.set_library(np.zeros((3, 4, 3), np.uint8))
    with pytest.raises(emo.EmosaicError, match="Dimensions must be divisible by 2"):
        ctx.match(np.zeros((5, 4, 3), np.uint8))
    with pytest.raises(emo.EmosaicError, match="not a square"):
        ctx.set_library(np.zeros((3, 5, 3), np.uint8))
    c2 = emo.Context(0)
    with pytest.raises(emo.EmosaicError, match="no library"):
        c2.match(np.zeros((4, 4, 3), np.uint8))
    c2.close()


# ---- the reference's end-to-end universe test (mod.rs:83-161) through the CUDA path ------------------
@pytest.mark.parametrize("N", [1, 4, 9])
def test_universe_roundtrip(ctx, N):
    from test_oracle_kat import universe
    dim = int(N ** 0.5)
    uni = universe(N)
    ts = emo.TileSet(N=N)
    colors = ctx.analyse_tiles(uni, dim)
    for i, img in enumerate(uni):
        ts.push_tile_with_image(f"{i}.png", colors[i], img)
    tall = uni.reshape(-1, dim, 3)
    res = emo.render_nto1(tall, ts, dim, ctx=ctx)
    assert (res.image == tall).all() and (res.dist == 0).all()
    for a in range(0, len(uni) - 1, 2):
        img = np.concatenate([uni[a], uni[a + 1]], 0)
        res = emo.render_nto1(img, ts, dim, ctx=ctx)
        assert (res.image == img).all()
        if a > 40:
            break
    # mod.rs:59-68 output dimensions
    src = np.zeros((2 * dim, 5 * dim, 3), np.uint8)
    assert emo.render_nto1(src, ts, dim, ctx=ctx).image.shape == (2 * dim, 5 * dim, 3)


# ---- compose / tint ------------------------------------------------------------------------------
@pytest.mark.parametrize("N,ts,T,bh,bw", [(1, 8, 100, 37, 64), (1, 16, 64, 10, 10), (1, 32, 16, 5, 7), (4, 16, 50, 9, 8),
                                          (1, 8, 20, 3, 5), (1, 4, 20, 6, 8), (1, 5, 9, 4, 3), (9, 12, 30, 5, 6), (1, 64, 4, 2, 3),
                                          (4, 6, 10, 3, 3), (1, 1, 5, 7, 9), (1, 16, 64, 5, 64), (4, 16, 40, 4, 32), (1, 8, 50, 3, 128), (1, 8, 50, 2, 192)])
def test_compose_parity(ctx, N, ts, T, bh, bw):
    rng = np.random.default_rng(ts * 7 + bw)
    tiles = rng.integers(0, 256, (T, ts, ts, 3), dtype=np.uint8)
    colors = oracle.analyse_tiles(tiles, N)
    item = rng.integers(1, T + 1, (bh, bw)).astype(np.int32)
    item[rng.random((bh, bw)) < 0.4] *= -1  # mirrored placements
    ctx.set_library(colors, tiles)
    out = ctx.compose(item)
    assert (out == oracle.render(tiles, item)).all()


@pytest.mark.parametrize("A", [0, 1, 64, 85, 127, 128, 153, 170, 200, 254, 255])
@pytest.mark.parametrize("N,ts", [(1, 8), (1, 32), (4, 16), (1, 6), (4, 8)])
def test_tint_parity(ctx, A, N, ts):
    rng = np.random.default_rng(A * 31 + ts)
    dim = int(N ** 0.5)
    T, bh, bw = 40, 6, 9
    tiles = rng.integers(0, 256, (T, ts, ts, 3), dtype=np.uint8)
    # make exceptions likely: tile pixels equal to the source colour
    src = rng.integers(0, 256, (bh * dim, bw * dim, 3), dtype=np.uint8)
    tiles[:, ::2, ::2] = src[0, 0]
    tiles[:, 1::2, ::2, 0] = 255
    tiles[:, ::2, 1::2, 1] = 0
    colors = oracle.analyse_tiles(tiles, N)
    item = rng.integers(1, T + 1, (bh, bw)).astype(np.int32)
    item[rng.random((bh, bw)) < 0.3] *= -1
    ctx.set_library(colors, tiles)
    got = ctx.compose(item, src, 4, A)
    want = oracle.tint(oracle.render(tiles, item), src, A)
    assert (got == want).all()


def test_tint_all_pairs_every_alpha(ctx):
    """Every (bg, fg) pair for every alpha through the fast kernel geometry: ts=16 tiles whose
    16 rows x 16 cols enumerate bg, source pixels enumerate fg."""
    bg = np.arange(256, dtype=np.uint8).reshape(16, 16)
    tiles = np.stack([bg, bg[::-1], bg.T], -1)[None]  # [1,16,16,3], each channel a permutation of 0..255
    src = np.zeros((16, 16, 3), np.uint8)
    fgv = np.arange(256, dtype=np.uint8).reshape(16, 16)
    src[..., 0], src[..., 1], src[..., 2] = fgv, fgv[::-1], fgv.T
    ctx.set_library(np.zeros((1, 1, 3), np.uint8), tiles)
    item = np.ones((16, 16), np.int32)
    base = oracle.render(tiles, item)
    for A in range(256):
        got = ctx.compose(item, src, 4, A)
        want = onp.tint(base, src, A)
        assert (got == want).all(), A


def test_compose_errors(ctx):
    tiles = np.zeros((3, 8, 8, 3), np.uint8)
    ctx.set_library(np.zeros((3, 1, 3), np.uint8), tiles)
    for bad in (0, 4, -4):
        item = np.ones((2, 2), np.int32)
        item[1, 1] = bad
        with pytest.raises(emo.EmosaicError, match="item map"):
            ctx.compose(item)
    ctx.set_library(np.zeros((3, 1, 3), np.uint8))  # no pixels resident
    with pytest.raises(emo.EmosaicError, match="no tile pixels"):
        ctx.compose(np.ones((2, 2), np.int32))
    with pytest.raises(emo.EmosaicError, match="Tile size must be divisible by 2"):
        ctx.set_library(np.zeros((3, 4, 3), np.uint8), np.zeros((3, 7, 7, 3), np.uint8))


# ---- golden fixtures -----------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["c1_1to1_t300", "c1_1to1_t300_smooth", "c2_4to1_small", "m3_9to1_small",
                                  "c5_tint_1to1", "tint_4to1_a200"])
def test_golden(ctx, name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    N = int(g["N"])
    dim = int(N ** 0.5)
    colors = ctx.analyse_tiles(g["tiles"], dim)
    assert (colors == g["colors"]).all()
    ctx.set_library(colors, g["tiles"])
    item, dist = ctx.match(g["src"])
    assert (item == g["item"]).all() and (dist == g["dist"]).all()
    assert sha(ctx.compose(item)) == str(g["out_sha256"])
    out, it2, d2 = ctx.mosaic(g["src"], 3, 0)
    assert sha(out) == str(g["out_sha256"]) and (it2 == item).all() and (d2 == dist).all()
    if "A" in g.files:
        assert sha(ctx.compose(item, g["src"], 4, int(g["A"]))) == str(g["tint_sha256"])
        out4, _, _ = ctx.mosaic(g["src"], 4, int(g["A"]))
        assert sha(out4) == str(g["tint_sha256"])


# ---- BASELINE configs at (near) full size: oracle where it finishes in seconds, properties beyond ------
def test_config2_4to1_full(ctx):
    """C2: 4to1, 10k tiles, 1024x1024 source, ts 16.  KD-tree oracle (exact, canonical ties) on the full map."""
    rng = np.random.default_rng(1234)
    tiles = rng.integers(0, 256, (10000, 16, 16, 3), dtype=np.uint8)
    src = np.random.default_rng(5678).integers(0, 256, (1024, 1024, 3), dtype=np.uint8)
    colors = ctx.analyse_tiles(tiles, 2)
    assert (colors == oracle.analyse_tiles(tiles, 4)).all()
    ctx.set_library(colors, tiles)
    out, item, dist = ctx.mosaic(src, 3, 0)
    ri, rd = oracle.KdTree(colors).match(src)
    assert (dist == rd).all() and (item == ri).all()
    assert out.shape == (8192, 8192, 3)
    assert (out == oracle.render(tiles, item)).all()


def test_config4_stripe_and_properties(ctx):
    """C4 geometry (100k tiles, ts 8, 4096-wide source): oracle parity on a 64-row stripe, plus size-independent
    properties on a 512-row stripe: dist equals the L1 distance to the chosen tile, every placed tile is copied
    verbatim, and matching the library's own colours returns distance 0 with the smallest duplicate index."""
    rng = np.random.default_rng(1234)
    T = 100_000
    tiles = rng.integers(0, 256, (T, 8, 8, 3), dtype=np.uint8)
    src = np.random.default_rng(5678).integers(0, 256, (512, 4096, 3), dtype=np.uint8)
    colors = ctx.analyse_tiles(tiles, 1)
    ctx.set_library(colors, tiles)
    out, item, dist = ctx.mosaic(src, 3, 0)
    kd = oracle.KdTree(colors)
    ri, rd = kd.match(src[:64])
    assert (item[:64] == ri).all() and (dist[:64] == rd).all()
    chosen = colors[np.abs(item) - 1, 0].astype(np.int64)
    assert (np.abs(chosen - src.astype(np.int64)).sum(-1) == dist).all()
    assert (item > 0).all()  # 1to1: the mirror coincides and never wins
    rows = rng.choice(512, 8, replace=False)
    for r in rows:
        assert (out[r * 8:(r + 1) * 8] == oracle.render(tiles, item[r:r + 1])).all()
    # the full C4 match (4096 x 4096 queries x 100 000 tiles): L1 property on all 16.7 M blocks, stripe equality
    full = np.random.default_rng(5678).integers(0, 256, (4096, 4096, 3), dtype=np.uint8)
    fi, fd = ctx.match(full)
    assert (fi[:512] == item).all() and (fd[:512] == dist).all()  # a stripe is just a smaller image
    assert (fi > 0).all()
    ch = colors[fi - 1, 0].astype(np.int16)
    assert (np.abs(ch - full.astype(np.int16)).sum(-1, dtype=np.int64) == fd).all()
    ri3, rd3 = kd.match(full[4000:4032])
    assert (fi[4000:4032] == ri3).all() and (fd[4000:4032] == rd3).all()
    # idempotence: the library's own colours as the source
    self_src = colors[:4096 * 8, 0].reshape(8, 4096, 3)
    it, ds = ctx.match(self_src)
    assert (ds == 0).all()
    ri2, _ = kd.match(self_src)
    assert (it == ri2).all()


def test_config4_full_map(ctx):
    """C4, the bench's headline workload, whole map: 4096 x 4096 source x 100 000 tiles.  item and dist of all 16.7 M blocks
    against the KD-tree oracle (exact L1 nearest_one with the canonical tie-break; 1-2 minutes of host time), through the index
    (both table forms) and through the brute-force scan."""
    T = 100_000
    tiles = np.random.default_rng(1234).integers(0, 256, (T, 8, 8, 3), dtype=np.uint8)
    full = np.random.default_rng(5678).integers(0, 256, (4096, 4096, 3), dtype=np.uint8)
    colors = ctx.analyse_tiles(tiles, 1)
    assert (colors == oracle.analyse_tiles(tiles, 1)).all()
    ctx.set_library(colors, tiles)
    ri, rd = oracle.KdTree(colors).match(full)
    try:
        for mode in ("auto", "index_wide", "index_compact", "scan"):
            ctx.set_match_mode(mode)
            gi, gd = ctx.match(full)
            assert (gi == ri).all(), f"{mode}: {(gi != ri).sum()} items differ from the oracle"
            assert (gd == rd).all(), f"{mode}: {(gd != rd).sum()} distances differ from the oracle"
    finally:
        ctx.set_match_mode("auto")


def test_config5_tint_large(ctx):
    """C5 geometry (ts 32, A=127, RGBA out) on a 128x256 source: oracle parity on the whole 4096x8192x4 image."""
    rng = np.random.default_rng(1234)
    tiles = rng.integers(0, 256, (4096, 32, 32, 3), dtype=np.uint8)
    src = np.random.default_rng(5678).integers(0, 256, (128, 256, 3), dtype=np.uint8)
    colors = ctx.analyse_tiles(tiles, 1)
    ctx.set_library(colors, tiles)
    out, item, dist = ctx.mosaic(src, 4, emo.tint_alpha(0.5))
    ri, rd = oracle.match(colors, src)
    assert (item == ri).all() and (dist == rd).all()
    assert (out == oracle.tint(oracle.render(tiles, item), src, 127)).all()


def test_config5_full_size(ctx):
    """C5 at full size: 4096 tiles of 32x32, 1024x1024 source, A = 127 -> 32768 x 32768 x 4 (4.29 GB, offsets beyond
    2^32).  Maps equal the KD-tree oracle everywhere; composited + tinted rows equal the oracle on sampled block rows."""
    rng = np.random.default_rng(1234)
    tiles = rng.integers(0, 256, (4096, 32, 32, 3), dtype=np.uint8)
    src = np.random.default_rng(5678).integers(0, 256, (1024, 1024, 3), dtype=np.uint8)
    colors = ctx.analyse_tiles(tiles, 1)
    ctx.set_library(colors, tiles)
    out, item, dist = ctx.mosaic(src, 4, 127)
    assert out.shape == (32768, 32768, 4)
    ri, rd = oracle.KdTree(colors).match(src)
    assert (item == ri).all() and (dist == rd).all()
    for r in (0, 1, 511, 777, 1022, 1023):
        want = oracle.tint(oracle.render(tiles, ri[r:r + 1]), src[r:r + 1], 127)
        assert (out[r * 32:(r + 1) * 32] == want).all(), r
    assert (out[..., 3] == 255).all()
    del out


def test_mosaic_chunked_pipeline_matches_single_calls(ctx):
    """emo_mosaic pipelines block-row chunks; result must equal match + compose done in one piece."""
    rng = np.random.default_rng(2)
    tiles = rng.integers(0, 256, (2000, 32, 32, 3), dtype=np.uint8)
    src = rng.integers(0, 256, (300, 700, 3), dtype=np.uint8)  # 645 MB RGB out -> several chunks
    colors = ctx.analyse_tiles(tiles, 1)
    ctx.set_library(colors, tiles)
    out, item, dist = ctx.mosaic(src, 3, 0)
    i2, d2 = ctx.match(src)
    assert (item == i2).all() and (dist == d2).all()
    o2 = ctx.compose(i2)
    assert sha(out) == sha(o2)


# ---- out-of-bounds write audit (compute-sanitizer is closed on this pool): canaries around every output -------
def _guarded(torch, nbytes, dev, pad=4096):
    buf = torch.full((pad + nbytes + pad,), 0xA5, dtype=torch.uint8, device=dev)
    return buf, buf.data_ptr() + pad, pad


def _intact(torch, buf, nbytes, pad):
    return bool((buf[:pad] == 0xA5).all()) and bool((buf[pad + nbytes:] == 0xA5).all())


@pytest.mark.parametrize("N,ts,T,bh,bw", [(1, 8, 300, 5, 64), (1, 8, 300, 3, 7), (1, 16, 200, 4, 32), (4, 16, 100, 3, 5), (1, 32, 50, 2, 9),
                                          (9, 12, 60, 2, 3), (1, 64, 20, 1, 2), (25, 10, 40, 2, 3)])
def test_no_out_of_bounds_writes(ctx, N, ts, T, bh, bw):
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda", 0)
    dim = int(N ** 0.5)
    rng = np.random.default_rng(N * 100 + ts)
    tiles_h = rng.integers(0, 256, (T, ts, ts, 3), dtype=np.uint8)
    src_h = rng.integers(0, 256, (bh * dim, bw * dim, 3), dtype=np.uint8)
    tiles = torch.from_numpy(tiles_h).to(dev)
    src = torch.from_numpy(src_h).to(dev)
    Q = bh * bw
    cbuf, cptr, pad = _guarded(torch, T * N * 3, dev)
    ibuf, iptr, _ = _guarded(torch, Q * 4, dev)
    dbuf, dptr, _ = _guarded(torch, Q * 4, dev)
    obuf, optr, _ = _guarded(torch, Q * ts * ts * 3, dev)
    tbuf, tptr, _ = _guarded(torch, Q * ts * ts * 4, dev)
    torch.cuda.synchronize()
    ctx.analyse_dev(tiles.data_ptr(), T, ts, dim, cptr)
    ctx.set_library_dev(cptr, tiles.data_ptr(), T, N, ts)
    ctx.match_dev(src.data_ptr(), bw * dim, bh * dim, iptr, dptr)
    ctx.compose_dev(iptr, 0, bw * dim, bh * dim, 3, 0, optr)
    ctx.compose_dev(iptr, src.data_ptr(), bw * dim, bh * dim, 4, 127, tptr)
    ctx.sync()
    for buf, n in ((cbuf, T * N * 3), (ibuf, Q * 4), (dbuf, Q * 4), (obuf, Q * ts * ts * 3), (tbuf, Q * ts * ts * 4)):
        assert _intact(torch, buf, n, pad)
    colors = cbuf[pad:pad + T * N * 3].cpu().numpy().reshape(T, N, 3)
    assert (colors == oracle.analyse_tiles(tiles_h, N)).all()
    ri, rd = oracle.match(colors, src_h)
    item = ibuf[pad:pad + Q * 4].cpu().numpy().view(np.int32).reshape(bh, bw)
    assert (item == ri).all()
    out = obuf[pad:pad + Q * ts * ts * 3].cpu().numpy().reshape(bh * ts, bw * ts, 3)
    assert (out == oracle.render(tiles_h, ri)).all()
    tint = tbuf[pad:pad + Q * ts * ts * 4].cpu().numpy().reshape(bh * ts, bw * ts, 4)
    assert (tint == oracle.tint(out, src_h, 127)).all()


def test_random_geometry_sweep(ctx):
    """80 seeded random geometries (mode, tile size, library size, source shape, alpha) through the whole path."""
    rng = np.random.default_rng(20260101)
    for case in range(80):
        dim = int(rng.choice([1, 1, 1, 2, 2, 3, 4, 5]))
        N = dim * dim
        ts = dim * int(rng.integers(1, 9)) if dim > 1 else int(rng.choice([1, 2, 3, 4, 5, 8, 8, 12, 16, 16, 32]))
        T = int(rng.choice([1, 2, 7, 64, 129, 300, 1000, 2049]))
        bh, bw = int(rng.integers(1, 12)), int(rng.choice([1, 3, 8, 32, 33, 64, 65]))
        A = int(rng.choice([0, 1, 60, 85, 127, 128, 200, 254, 255]))
        tiles = rng.integers(0, 256, (T, ts, ts, 3), dtype=np.uint8)
        if case % 3 == 0:  # few distinct colours: many exact ties
            tiles = (tiles // 128 * 128).astype(np.uint8)
        src = rng.integers(0, 256, (bh * dim, bw * dim, 3), dtype=np.uint8)
        if case % 3 == 0:
            src = (src // 128 * 128).astype(np.uint8)
        tag = f"case {case}: N={N} ts={ts} T={T} blocks={bh}x{bw} A={A}"
        colors = ctx.analyse_tiles(tiles, dim)
        assert (colors == oracle.analyse_tiles(tiles, N)).all(), tag
        ctx.set_library(colors, tiles)
        out, item, dist = ctx.mosaic(src, 4, A)
        ri, rd = oracle.match(colors, src)
        assert (item == ri).all() and (dist == rd).all(), tag
        rgb = oracle.render(tiles, ri)
        assert (out == oracle.tint(rgb, src, A)).all(), tag
        assert (ctx.compose(item) == rgb).all(), tag


def test_unaligned_device_pointers(ctx):
    """Device pointers with odd alignment take the generic kernels (or are rejected where a 4-byte store is needed)."""
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(77)
    T, ts, bh, bw = 90, 16, 3, 32
    tiles_h = rng.integers(0, 256, (T, ts, ts, 3), dtype=np.uint8)
    src_h = rng.integers(0, 256, (bh, bw, 3), dtype=np.uint8)
    raw = torch.zeros(T * ts * ts * 3 + 64, dtype=torch.uint8, device=dev)
    raw[1:1 + T * ts * ts * 3] = torch.from_numpy(tiles_h.reshape(-1)).to(dev)     # tiles at an odd address
    colors = torch.empty(T * 3 + 8, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    ctx.analyse_dev(raw.data_ptr() + 1, T, ts, 1, colors.data_ptr() + 1)
    ctx.sync()
    ch = colors[1:1 + T * 3].cpu().numpy().reshape(T, 1, 3)
    assert (ch == oracle.analyse_tiles(tiles_h, 1)).all()
    ctx.set_library(ch, tiles_h)
    item_h, _ = oracle.match(ch, src_h)
    item = torch.from_numpy(item_h.reshape(-1)).to(dev)
    src = torch.from_numpy(src_h.reshape(-1)).to(dev)
    out = torch.zeros(bh * ts * bw * ts * 4 + 64, dtype=torch.uint8, device=dev)
    ctx.compose_dev(item.data_ptr(), 0, bw, bh, 3, 0, out.data_ptr() + 3)            # RGB out at an odd address
    ctx.sync()
    got = out[3:3 + bh * ts * bw * ts * 3].cpu().numpy().reshape(bh * ts, bw * ts, 3)
    assert (got == oracle.render(tiles_h, item_h)).all()
    ctx.compose_dev(item.data_ptr(), src.data_ptr(), bw, bh, 4, 127, out.data_ptr() + 4)   # 4-byte aligned RGBA: generic path
    ctx.sync()
    got4 = out[4:4 + bh * ts * bw * ts * 4].cpu().numpy().reshape(bh * ts, bw * ts, 4)
    assert (got4 == oracle.tint(oracle.render(tiles_h, item_h), src_h, 127)).all()
    with pytest.raises(emo.EmosaicError, match="4-byte aligned"):
        ctx.compose_dev(item.data_ptr(), src.data_ptr(), bw, bh, 4, 127, out.data_ptr() + 2)
    with pytest.raises(emo.EmosaicError, match="4-byte aligned"):
        ctx.match_dev(src.data_ptr(), bw, bh, out.data_ptr() + 1, out.data_ptr() + 8)


@pytest.mark.parametrize("A", [0, 64, 127, 200, 255])
def test_tint_overlay_of_another_size(ctx, A):
    """main.rs:447-478 overlays the ORIGINAL image (any size) nearest-resized to the output: emo_compose_overlay."""
    rng = np.random.default_rng(A + 5)
    for N, ts, bh, bw, oh, ow in ((1, 8, 9, 13, 37, 50), (4, 16, 5, 6, 21, 25), (1, 5, 7, 4, 3, 2), (9, 12, 3, 4, 100, 90)):
        dim = int(N ** 0.5)
        T = 30
        tiles = rng.integers(0, 256, (T, ts, ts, 3), dtype=np.uint8)
        colors = oracle.analyse_tiles(tiles, N)
        item = rng.integers(1, T + 1, (bh, bw)).astype(np.int32)
        item[rng.random((bh, bw)) < 0.3] *= -1
        overlay = rng.integers(0, 256, (oh, ow, 3), dtype=np.uint8)
        ctx.set_library(colors, tiles)
        got = ctx.compose_overlay(item, overlay, A)
        want = oracle.tint(oracle.render(tiles, item), overlay, A)
        assert (got == want).all(), (N, ts, bh, bw, oh, ow)


def test_stats_reduction_on_the_gpu(ctx):
    """emo_stats (the reductions of RenderStats, stats.rs:87-139 / :169-175) against numpy on maps with mirrored ids and unplaced
    blocks; the summary built from it equals the host-only one; an id beyond the library is rejected."""
    import io
    from emosaic_b200 import stats
    rng = np.random.default_rng(8)
    T = 777
    ctx.set_library(rng.integers(0, 256, (T, 4, 3), dtype=np.uint8))
    item = (rng.integers(1, T + 1, (300, 411)) * rng.choice([-1, 1], (300, 411))).astype(np.int32)
    item[rng.random(item.shape) < 0.1] = 0
    item[:, :7] = 5                                   # a heavily used tile
    dist = rng.integers(0, 3060, item.shape).astype(np.uint32)
    sums, usage = ctx.stats(item, dist)
    placed = item != 0
    assert sums == {"placed": int(placed.sum()), "total_distance": int(dist[placed].astype(np.uint64).sum()),
                    "max_distance": int(dist[placed].max())}
    want = np.bincount(np.abs(item[placed]) - 1, minlength=T)
    assert (usage == want).all()
    a = stats.summarise(item, dist, None, file=io.StringIO(), ctx=ctx)
    b = stats.summarise(item, dist, None, file=io.StringIO())
    assert a == b
    empty = np.zeros((3, 3), np.int32)
    assert ctx.stats(empty, empty.astype(np.uint32))[0] == {"placed": 0, "total_distance": 0, "max_distance": 0}
    bad = item.copy(); bad[0, 0] = T + 1
    with pytest.raises(emo.EmosaicError):
        ctx.stats(bad, dist)


def test_reserve_presizes_the_staging_buffers(ctx):
    """emo_reserve: after it, the host-pointer calls for images up to that size do not allocate (device memory in use does not
    change across the first emo_mosaic), and results are what they were."""
    import torch
    rng = np.random.default_rng(31)
    tiles = rng.integers(0, 256, (500, 8, 8, 3), dtype=np.uint8)
    colors = oracle.analyse_tiles(tiles, 4)
    c = emo.Context(0)
    try:
        c.set_library(colors, tiles)
        with pytest.raises(emo.EmosaicError):
            c.reserve(0, 10, 3)
        c.reserve(640, 480, 4)
        c.sync()
        free0 = torch.cuda.mem_get_info(0)[0]
        src = rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)
        out, item, dist = c.mosaic(src, 4, 127)
        out3, _, _ = c.mosaic(src[:100, :200], 3, 0)
        assert torch.cuda.mem_get_info(0)[0] == free0, "a staged call allocated after emo_reserve"
        ri, rd = oracle.match(colors, src)
        assert (item == ri).all() and (dist == rd).all()
        assert (out == oracle.tint(oracle.render(tiles, ri), src, 127)).all()
        # 1to1: the search index's tables are part of the reservation
        c1 = oracle.analyse_tiles(tiles, 1)
        c.set_library(c1, tiles)
        c.set_match_mode("index")
        c.reserve(256, 256, 3)
        c.sync()
        free1 = torch.cuda.mem_get_info(0)[0]
        o1, i1, d1 = c.mosaic(src[:256, :256], 3, 0)
        assert torch.cuda.mem_get_info(0)[0] == free1, "the first 1to1 match allocated after emo_reserve"
        r1, rd1 = oracle.match(c1, src[:256, :256])
        assert (i1 == r1).all() and (d1 == rd1).all()
    finally:
        c.close()
    fresh = emo.Context(0)
    try:
        with pytest.raises(emo.EmosaicError):
            fresh.reserve(64, 64, 3)          # no library yet
    finally:
        fresh.close()


@pytest.mark.parametrize("N,T,H,W", [(4, 4000, 2 * 449, 2 * 512), (4, 4000, 2 * 460, 2 * 512), (1, 20000, 300, 2048 + 40), (9, 2500, 3 * 150, 3 * 768)])
def test_scan_with_a_ragged_last_wave(ctx, N, T, H, W):
    """Query-tile counts just above a whole number of resident CTA waves: the remainder tiles are dealt out in parts (split CTAs +
    whole-range CTAs in one grid, merged through the (distance, window) keys).  Oracle parity on the whole map."""
    rng = np.random.default_rng(N * T)
    colors = (rng.integers(0, 256, (T, N, 3)) // 8 * 8).astype(np.uint8)      # coarse colours: ties across the part boundaries
    src = (rng.integers(0, 256, (H, W, 3)) // 4 * 4).astype(np.uint8)
    ctx.set_library(colors)
    ctx.set_match_mode("scan")
    try:
        item, dist = ctx.match(src)
    finally:
        ctx.set_match_mode("auto")
    ri, rd = oracle.KdTree(colors).match(src)
    assert (dist == rd).all() and (item == ri).all()
