"""Pins the CPU oracle (C and numpy restatements) to every known-answer vector the reference's own
unit tests hold for the hot path.  CPU only."""
import itertools

import numpy as np
import pytest

import oracle
from oracle import oracle_np as onp


# --- src/mosaic/color.rs:49-72 -----------------------------------------------------------------
def test_average_color_basic():
    img = np.array([[[100, 150, 200], [200, 100, 50]], [[50, 200, 100], [150, 50, 150]]], np.uint8)
    assert oracle.average_color(img, (0, 0, 2, 2)).tolist() == [125, 125, 125]


def test_average_color_single_pixel():
    img = np.tile(np.array([42, 84, 126], np.uint8), (3, 3, 1))
    assert oracle.average_color(img, (1, 1, 1, 1)).tolist() == [42, 84, 126]


# --- color.rs:75-99 (should_panic messages) ------------------------------------------------------
@pytest.mark.parametrize("shape,rect,msg", [
    ((10, 10), (0, 0, 0, 5), "Rectangle dimensions must be positive"),
    ((10, 10), (0, 0, 5, 0), "Rectangle dimensions must be positive"),
    ((5, 5), (3, 0, 5, 2), "Rectangle extends beyond image width"),
    ((5, 5), (0, 3, 2, 5), "Rectangle extends beyond image height"),
])
def test_average_color_panics(shape, rect, msg):
    img = np.zeros(shape + (3,), np.uint8)
    with pytest.raises(oracle.OracleError, match=msg):
        oracle.average_color(img, rect)


# --- analysis.rs:44-55 ---------------------------------------------------------------------------
def test_analyse_single_color():
    img = np.zeros((2, 2, 3), np.uint8)
    img[..., 0] = 255
    for fn in (oracle.analyse, onp.analyse):
        c = fn(img, 4)
        assert c.tolist() == [[255, 0, 0]] * 4


# --- analysis.rs:58-71 ---------------------------------------------------------------------------
def test_get_img_colors():
    img = np.zeros((4, 4, 3), np.uint8)
    for y in range(4):
        for x in range(4):
            img[y, x] = (x * 64, y * 64, 128)
    c = oracle.get_img_colors(img, 0, 0, 2, 4)
    assert c.tolist() == [[0, 0, 128], [64, 0, 128], [0, 64, 128], [64, 64, 128]]
    assert onp.queries(img, 4)[0, 0].reshape(4, 3).tolist() == c.tolist()


# --- tiles/tile.rs:127-140 -----------------------------------------------------------------------
def test_tile_coords():
    assert oracle.coords([[1, 2, 3]]).tolist() == [1, 2, 3]
    assert oracle.coords([[1, 2, 3], [4, 5, 6], [7, 8, 9], [10, 11, 12]]).tolist() == list(range(1, 13))


# --- tiles/utils.rs:302-308 ----------------------------------------------------------------------
def test_flipped_coords():
    c = list(range(1, 13))
    f = oracle.flipped_coords(c)
    assert f.tolist() == [4, 5, 6, 1, 2, 3, 10, 11, 12, 7, 8, 9]
    assert oracle.flipped_coords(f).tolist() == c
    assert onp.mirror(np.array(c)[None], 4)[0].tolist() == f.tolist()
    assert oracle.coords(np.array(c).reshape(4, 3), flipped=True).tolist() == f.tolist()


# --- mod.rs:48-68 (output dimensions) --------------------------------------------------------------
def test_render_shapes():
    src = np.zeros((2, 5, 3), np.uint8)
    colors = np.zeros((1, 1, 3), np.uint8)
    tiles = np.zeros((1, 8, 8, 3), np.uint8)
    item, _ = oracle.match(colors, src)
    out = oracle.render(tiles, item)
    assert out.shape == (2 * 8, 5 * 8, 3)


# --- mod.rs:83-161: the N = 1 / 4 / 9 black&white universe ------------------------------------------
def universe(N):
    dim = int(N ** 0.5)
    imgs = []
    for index in range(2 ** N - 1):  # all-white excluded, mod.rs:91-92
        bits = [(index & (1 << i)) != 0 for i in range(N)][::-1]
        img = np.zeros((dim, dim, 3), np.uint8)
        for y in range(dim):
            for x in range(dim):
                img[y, x] = 255 if bits[y * dim + x] else 0
        imgs.append(img)
    return np.stack(imgs)


@pytest.mark.parametrize("N", [1, 4, 9])
def test_universe_roundtrip(N):
    dim = int(N ** 0.5)
    uni = universe(N)
    colors = oracle.analyse_tiles(uni, N)
    assert (colors == onp.analyse_tiles(uni, N)).all()
    # every universe member rendered with tile_size = dim reproduces itself (mod.rs:112-127)
    tall = uni.reshape(-1, dim, 3)  # all members stacked vertically: one render call covers them all
    item, dist = oracle.match(colors, tall)
    assert (dist == 0).all()
    assert (oracle.render(uni, item) == tall).all()
    i2, d2 = onp.match(colors, tall)
    assert (i2 == item).all() and (d2 == dist).all()
    # 1x2 vertical stacks of consecutive members (mod.rs:130-145)
    for a in range(0, len(uni) - 1, 2):
        img = np.concatenate([uni[a], uni[a + 1]], 0)
        it, ds = oracle.match(colors, img)
        assert (oracle.render(uni, it) == img).all() and (ds == 0).all()


# --- canonical tie-break = kiddo single-leaf scan order ---------------------------------
def test_tie_break_insertion_order():
    # tiles 1 and 3 identical, tile 2 further away: winner must be +1 (smallest idx, unflipped first)
    colors = np.array([[[10, 10, 10]], [[200, 0, 0]], [[10, 10, 10]]], np.uint8)
    src = np.array([[[12, 10, 10]]], np.uint8)
    item, dist = oracle.match(colors, src)
    assert item.tolist() == [[1]] and dist.tolist() == [[2]]
    # 4to1: a left/right symmetric tile ties with its own mirror -> unflipped wins
    sym = np.array([[[5, 5, 5], [5, 5, 5], [9, 9, 9], [9, 9, 9]]], np.uint8)
    q = np.array([[[5, 5, 5], [5, 5, 5]], [[9, 9, 9], [9, 9, 9]]], np.uint8)
    item, dist = oracle.match(sym, q)
    assert item.tolist() == [[1]] and dist.tolist() == [[0]]
    # an asymmetric tile whose mirror is the exact match -> negative id
    asym = np.array([[[1, 1, 1], [200, 200, 200], [1, 1, 1], [200, 200, 200]]], np.uint8)
    q = np.array([[[200, 200, 200], [1, 1, 1]], [[200, 200, 200], [1, 1, 1]]], np.uint8)
    item, dist = oracle.match(asym, q)
    assert item.tolist() == [[-1]] and dist.tolist() == [[0]]


def test_kdtree_equals_bruteforce():
    rng = np.random.default_rng(7)
    for N, T in ((1, 3000), (4, 1500), (9, 400)):
        dim = int(N ** 0.5)
        colors = rng.integers(0, 256, (T, N, 3), dtype=np.uint8)
        colors[T // 2:T // 2 + 50] = colors[:50]  # exact duplicates -> ties across leaves
        src = rng.integers(0, 256, (12 * dim, 20 * dim, 3), dtype=np.uint8)
        src[:2 * dim, :4 * dim] = onp_render_like(colors[:8], dim)
        a = oracle.match(colors, src)
        b = oracle.KdTree(colors, bucket=64).match(src)
        c = oracle.KdTree(colors).match(src)
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all()
        assert (a[0] == c[0]).all() and (a[1] == c[1]).all()


def onp_render_like(cols, dim):
    # lay 8 candidate vectors out as exact-match source blocks (2 block rows x 4 block cols)
    v = cols.reshape(2, 4, dim, dim, 3).transpose(0, 2, 1, 3, 4)
    return v.reshape(2 * dim, 4 * dim, 3)


# --- tint: derived known answers (oracle-derived, not reference-pinned) ------------------
def test_blend_known_answers():
    lut, a = oracle.blend_lut(127)
    assert a == 255
    for (bg, fg), want in {(0, 255): 127, (255, 0): 127, (100, 200): 149, (10, 20): 14, (255, 255): 255, (128, 128): 128}.items():
        assert lut[bg, fg] == want
    for A, ab in ((1, 255), (64, 255), (127, 255), (254, 255), (128, 254), (200, 254)):
        assert oracle.blend_lut(A)[1] == ab
    for A in (0, 1, 127, 128, 255):
        bg, fg = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
        rgb, ab = onp.blend(bg, fg, A)
        l2, a2 = oracle.blend_lut(A)
        assert (rgb == l2).all() and ab == a2


def test_tint_alpha():
    assert oracle.tint_alpha(0.5) == 127
    assert oracle.tint_alpha(1.0) == 255
    assert oracle.tint_alpha(0.0) == 0
    assert oracle.tint_alpha(0.999) == 254


# --- golden fixtures are self-consistent with the oracle of this checkout ---------------------------
@pytest.mark.parametrize("name", ["c1_1to1_t300", "c1_1to1_t300_smooth", "c2_4to1_small", "m3_9to1_small",
                                  "c5_tint_1to1", "tint_4to1_a200", "ties_palette"])
def test_golden_matches_oracle(name):
    import hashlib
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    N = int(g["N"])
    assert (oracle.analyse_tiles(g["tiles"], N) == g["colors"]).all()
    item, dist = oracle.match(g["colors"], g["src"])
    assert (item == g["item"]).all() and (dist == g["dist"]).all()
    out = oracle.render(g["tiles"], item)
    assert hashlib.sha256(out.tobytes()).hexdigest() == str(g["out_sha256"])
    assert (onp.render(g["tiles"], item) == out).all()
    if "A" in g.files:
        t = oracle.tint(out, g["src"], int(g["A"]))
        assert hashlib.sha256(t.tobytes()).hexdigest() == str(g["tint_sha256"])
        assert (onp.tint(out, g["src"], int(g["A"])) == t).all()
