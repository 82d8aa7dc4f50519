"""Ranked candidate lists (emo_topk) and the no-repeat renderer on the GPU against the oracle (-m gpu).

Reference: render_nto1_no_repeat, rendering.rs:262-401; pinned by mod.rs:118-145 (universe round trips)."""
import numpy as np
import pytest

import emosaic_b200 as emo
import oracle
from oracle import oracle_np as onp
from test_no_repeat_oracle import universe

pytestmark = pytest.mark.gpu


def expected_lists(colors, src):
    items, dists = onp.sorted_candidates(colors, src)
    if colors.shape[1] == 1:      # N = 1: the mirrored twins are omitted by the library
        keep = items > 0
        T = colors.shape[0]
        items = items[keep].reshape(-1, T)
        dists = dists[keep].reshape(-1, T)
    return items, dists


@pytest.mark.parametrize("N,T,bh,bw", [(1, 300, 6, 7), (4, 200, 5, 4), (9, 77, 3, 5), (16, 50, 2, 3), (1, 1, 2, 2), (4, 1, 1, 3),
                                       (1, 3000, 3, 3), (4, 1500, 2, 4)])
def test_topk_pages(ctx, N, T, bh, bw):
    dim = int(N ** 0.5)
    rng = np.random.default_rng(N * 100 + T)
    colors = rng.integers(0, 256, (T, N, 3), dtype=np.uint8)
    src = rng.integers(0, 256, (bh * dim, bw * dim, 3), dtype=np.uint8)
    ctx.set_library(colors)
    ei, ed = expected_lists(colors, src)
    L = ei.shape[1]
    for first, k in [(0, 1), (0, 64), (0, 1024), (5, 33), (L - 1, 4), (L, 8), (L + 10, 2), (max(0, L - 700), 1000), (17, 128)]:
        gi, gd = ctx.topk(src, first, k)
        n = max(0, min(k, L - first))
        assert (gi[:, :n] == ei[:, first:first + n]).all(), (first, k)
        assert (gd[:, :n] == ed[:, first:first + n]).all(), (first, k)
        assert (gi[:, n:] == 0).all() and (gd[:, n:] == 0xFFFFFFFF).all()


def test_topk_heavy_ties(ctx):
    # palette libraries: thousands of candidates at the same distance -> the order inside a distance is the insertion rank
    rng = np.random.default_rng(5)
    for N in (1, 4):
        dim = int(N ** 0.5)
        colors = (rng.integers(0, 2, (900, N, 3)) * 200).astype(np.uint8)
        src = (rng.integers(0, 2, (4 * dim, 5 * dim, 3)) * 200 + 7).astype(np.uint8)
        ctx.set_library(colors)
        ei, ed = expected_lists(colors, src)
        for first, k in [(0, 256), (100, 1024), (513, 300)]:
            gi, gd = ctx.topk(src, first, k)
            n = min(k, ei.shape[1] - first)
            assert (gi[:, :n] == ei[:, first:first + n]).all() and (gd[:, :n] == ed[:, first:first + n]).all()
    colors = np.full((500, 4, 3), 9, np.uint8)          # every candidate identical: pure rank order +1,-1,+2,-2,...
    ctx.set_library(colors)
    gi, gd = ctx.topk(np.zeros((2, 2, 3), np.uint8), 0, 1000)
    want = np.stack([np.arange(1, 501), -np.arange(1, 501)], 1).reshape(-1)
    assert (gi[0] == want).all() and (gd[0] == 9 * 12).all()


@pytest.mark.parametrize("N,T", [(1, 500), (4, 300), (9, 64)])
def test_topk_excluding_retired_tiles(ctx, N, T):
    """exclude = the tiles already placed: their candidates (both orientations) vanish from every list, which is what the
    reference's refill sees after tree.remove (rendering.rs:366-386)."""
    dim = int(N ** 0.5)
    rng = np.random.default_rng(N + T)
    colors = rng.integers(0, 256, (T, N, 3), dtype=np.uint8)
    src = rng.integers(0, 256, (3 * dim, 4 * dim, 3), dtype=np.uint8)
    ctx.set_library(colors)
    ei, ed = expected_lists(colors, src)
    for frac in (0.0, 0.3, 0.9, 1.0):
        retired = (rng.random(T) < frac).astype(np.uint8)
        if frac == 1.0:
            retired[:] = 1
        for first, k in [(0, 50), (7, 200)]:
            gi, gd = ctx.topk(src, first, k, exclude=retired)
            for q in range(ei.shape[0]):
                keep = retired[np.abs(ei[q]) - 1] == 0
                wi, wd = ei[q][keep][first:first + k], ed[q][keep][first:first + k]
                assert (gi[q, :wi.size] == wi).all() and (gd[q, :wi.size] == wd).all(), (frac, first, k, q)
                assert (gi[q, wi.size:] == 0).all() and (gd[q, wi.size:] == 0xFFFFFFFF).all()
    with pytest.raises(emo.EmosaicError, match="one byte per tile"):
        ctx.topk(src, 0, 4, exclude=np.zeros(T + 1, np.uint8))


def test_topk_first_entry_is_the_match(ctx):
    rng = np.random.default_rng(8)
    colors = rng.integers(0, 256, (5000, 4, 3), dtype=np.uint8)
    src = rng.integers(0, 256, (64, 64, 3), dtype=np.uint8)
    ctx.set_library(colors)
    gi, gd = ctx.topk(src, 0, 4)
    mi, md = ctx.match(src)
    assert (gi[:, 0].reshape(32, 32) == mi).all() and (gd[:, 0].reshape(32, 32) == md).all()
    assert (gd[:, 1:] >= gd[:, :-1]).all()


def test_topk_errors(ctx):
    ctx.set_library(np.zeros((3, 4, 3), np.uint8))
    with pytest.raises(emo.EmosaicError, match="outside \\[1,1024\\]"):
        ctx.topk(np.zeros((2, 2, 3), np.uint8), 0, 2000)
    with pytest.raises(emo.EmosaicError, match="divisible by 2"):
        ctx.topk(np.zeros((3, 2, 3), np.uint8), 0, 4)
    ctx.set_library(np.zeros((3, 25, 3), np.uint8))
    with pytest.raises(emo.EmosaicError, match="mode 1..4"):
        ctx.topk(np.zeros((5, 5, 3), np.uint8), 0, 4)


@pytest.mark.parametrize("N,ts,T,bh,bw,page", [(1, 8, 400, 12, 15, 64), (4, 16, 150, 8, 9, 64), (9, 12, 60, 4, 5, 64), (4, 8, 90, 9, 10, 2),
                                              (1, 8, 60, 8, 8, 4), (4, 8, 13, 5, 5, 3), (1, 16, 40, 7, 8, 1)])
def test_render_no_repeat_parity(ctx, N, ts, T, bh, bw, page):
    """Oracle parity of the whole renderer, including forced list refills (tiny pages), contention (few tiles, many blocks)
    and libraries too small to fill the image (T < blocks <= 2T: the rest stays black)."""
    dim = int(N ** 0.5)
    rng = np.random.default_rng(N * 1000 + T)
    tiles = rng.integers(0, 256, (T, ts, ts, 3), dtype=np.uint8)
    colors = oracle.analyse_tiles(tiles, N)
    src = rng.integers(0, 256, (bh * dim, bw * dim, 3), dtype=np.uint8)
    tset = emo.TileSet.from_arrays(colors, tiles)
    res = emo.render_nto1_no_repeat(src, tset, ts, ctx, page=page)
    ri, rd = onp.no_repeat_assign(colors, src)
    assert (res.item == ri).all() and (res.dist == rd).all()
    want = oracle.render(tiles, np.where(ri == 0, 1, ri))
    want[np.repeat(np.repeat(ri == 0, ts, 0), ts, 1)] = 0
    assert (res.image == want).all()
    placed = np.abs(ri[ri != 0])
    assert len(set(placed.tolist())) == placed.size


def test_render_no_repeat_contention(ctx):
    # clustered library + smooth source: the same few tiles are everybody's favourites, lists are consumed deep
    rng = np.random.default_rng(77)
    T, ts = 700, 8
    tiles = np.clip(rng.normal(127, 6, (T, ts, ts, 3)), 0, 255).astype(np.uint8)
    colors = oracle.analyse_tiles(tiles, 1)
    src = np.clip(rng.normal(127, 3, (24, 28, 3)), 0, 255).astype(np.uint8)      # 672 blocks for 700 tiles
    res = emo.render_nto1_no_repeat(src, emo.TileSet.from_arrays(colors, tiles), ts, ctx, page=16)
    ri, rd = onp.no_repeat_assign(colors, src)
    assert (res.item == ri).all() and (res.dist == rd).all()
    assert (res.image == oracle.render(tiles, ri)).all()


@pytest.mark.parametrize("N,T,bh,bw,page", [(4, 5000, 60, 60, 0), (4, 5000, 60, 60, 8), (1, 3000, 50, 60, 0), (1, 3000, 50, 60, 2),
                                            (9, 2000, 30, 40, 16)])
def test_no_repeat_assignment_large(ctx, N, T, bh, bw, page):
    """emo_no_repeat against the C oracle at sizes where pages run dry in bulk (batched refills) and with the automatic deep
    first page; the refill path must have been exercised when the page is small."""
    dim = int(N ** 0.5)
    rng = np.random.default_rng(N + T)
    colors = rng.integers(0, 256, (T, N, 3), dtype=np.uint8)
    src = rng.integers(0, 256, (bh * dim, bw * dim, 3), dtype=np.uint8)
    ctx.set_library(colors)
    item, dist, cnt = ctx.no_repeat(src, page)
    ri, rd = oracle.no_repeat_assign(colors, src)
    assert (item == ri).all() and (dist == rd).all()
    assert cnt["placed"] == (ri != 0).sum()
    if page and page <= 8:
        assert cnt["refill_launches"] > 0 and cnt["blocks_refilled"] >= cnt["refill_launches"]


def test_no_repeat_120x120_blocks_from_20k_tiles(ctx):
    """The size DESIGN.md quotes (a 120 x 120-block render from 20 000 tiles, 4to1): parity with the C oracle, every tile at most
    once, and the whole assignment (lists + merge) well under a second (it took 5 s while the merge lived in Python)."""
    import time
    rng = np.random.default_rng(120)
    T = 20_000
    colors = rng.integers(0, 256, (T, 4, 3), dtype=np.uint8)
    src = rng.integers(0, 256, (240, 240, 3), dtype=np.uint8)
    ctx.set_library(colors)
    ctx.no_repeat(src[:8, :8])                                # warm-up (buffers)
    t0 = time.perf_counter()
    item, dist, cnt = ctx.no_repeat(src)
    dt = time.perf_counter() - t0
    print(f"no_repeat 120x120 blocks / 20k tiles: {dt * 1e3:.1f} ms, {cnt}")
    assert dt < 1.0, f"{dt:.2f} s"
    used = np.abs(item[item != 0])
    assert used.size == 14400 and len(set(used.tolist())) == used.size
    ri, rd = oracle.no_repeat_assign(colors, src)             # 2.3 GB of candidate order on the host
    assert (item == ri).all() and (dist == rd).all()


@pytest.mark.parametrize("N", [1, 4, 9])
def test_universe_no_repeat_gpu(ctx, N):
    """mod.rs:118-145 through the CUDA path."""
    dim = int(N ** 0.5)
    uni = universe(N)
    colors = ctx.analyse_tiles(uni, dim)
    tset = emo.TileSet.from_arrays(colors, uni)
    for img in uni[:48]:
        res = emo.render_nto1_no_repeat(img, tset, dim, ctx)
        assert (res.image == img).all() and (res.dist == 0).all()
    for a in range(0, min(len(uni) - 1, 48), 2):
        img = np.concatenate([uni[a], uni[a + 1]], 0)
        res = emo.render_nto1_no_repeat(img, tset, dim, ctx)
        assert (res.image == img).all()


def test_render_no_repeat_errors(ctx):
    rng = np.random.default_rng(1)
    tiles = rng.integers(0, 256, (12, 8, 8, 3), dtype=np.uint8)
    tset = emo.TileSet.from_arrays(oracle.analyse_tiles(tiles, 4), tiles)
    with pytest.raises(emo.EmosaicError, match="Insufficient tiles for no-repeat mode: need 25 tiles but only have 24"):
        emo.render_nto1_no_repeat(rng.integers(0, 256, (10, 10, 3), dtype=np.uint8), tset, 8, ctx)
    with pytest.raises(emo.EmosaicError, match="divisible by 2"):
        emo.render_nto1_no_repeat(rng.integers(0, 256, (5, 4, 3), dtype=np.uint8), tset, 8, ctx)
