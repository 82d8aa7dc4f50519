"""CPU checks of the Lanczos3 resize / prepare_tile restatements (image 0.25.2 imageops::resize at main.rs:595 and
tiles/utils.rs:188-189; trim view utils.rs:93-186): the C oracle against the independent numpy restatement, hand-derived
known answers, the committed golden vectors, and the product's host-side bookkeeping (prepare_view, rotate,
most_common_value, resize_source's dimension rule) against the oracle.  No reference test pins a resize; the only
reference vectors on this stage are utils.rs:284-289 (most_common_value) and :291-299 (prepare_tile output is ts x ts)."""
import os

import numpy as np
import pytest

import oracle
from oracle import oracle_np as onp
from emosaic_b200 import api

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

GEOMS = [(17, 23, 8, 8), (9, 7, 20, 31), (33, 35, 32, 34), (64, 64, 64, 64), (5, 5, 1, 1), (1, 1, 7, 3), (40, 30, 39, 30),
         (100, 80, 7, 5), (2, 300, 2, 11), (31, 2, 5, 2)]


@pytest.mark.parametrize("h,w,nh,nw", GEOMS)
def test_c_oracle_equals_numpy_restatement(h, w, nh, nw):
    img = np.random.default_rng(h * 1000 + w).integers(0, 256, (h, w, 3), dtype=np.uint8)
    assert (oracle.resize_lanczos3(img, nw, nh) == onp.resize_lanczos3(img, nw, nh)).all()


def test_axis_taps():
    # downscale 8 -> 2: ratio 4, support 12 covers the whole axis from both centres
    left, cnt, ws = oracle.resize_axis(8, 2)
    assert left.tolist() == [0, 0] and cnt.tolist() == [8, 8]
    assert np.allclose(ws.sum(axis=1), 1.0, atol=1e-6)
    assert np.allclose(ws[0], ws[1][::-1], atol=1e-7)          # mirror symmetry of the two output positions
    # upscale 4 -> 8: support stays 3 source pixels either side of the centre
    left, cnt, ws = oracle.resize_axis(4, 8)
    assert cnt.max() <= 7 and (left + cnt <= 4).all()
    # the numpy restatement builds the same taps
    for (l, w), l2, c2, w2 in zip(onp.resize_axis(8, 2), *oracle.resize_axis(8, 2)):
        assert l == l2 and len(w) == c2 and (w == w2[:c2]).all()
    # identity ratio: the centre tap carries (almost) all the weight, sinc zeros elsewhere
    left, cnt, ws = oracle.resize_axis(9, 9)
    assert all(abs(ws[o, o - left[o]] - 1.0) < 1e-6 for o in range(9))


def test_known_answers():
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (12, 10, 3), dtype=np.uint8)
    assert (oracle.resize_lanczos3(img, 10, 12) == img).all()                      # same dimensions: copy
    flat = np.full((30, 41, 3), (7, 200, 255), np.uint8)
    assert (oracle.resize_lanczos3(flat, 13, 9) == flat[:9, :13]).all()            # normalised weights keep a flat image
    assert (oracle.resize_lanczos3(flat, 90, 70) == (7, 200, 255)).all()
    one = oracle.resize_lanczos3(img, 1, 1)                                        # 1 x 1: windowed mean of the image
    assert np.abs(one.astype(int).reshape(3) - img.reshape(-1, 3).mean(0)).max() <= 12
    # mirror symmetry: resizing the mirrored image mirrors the result (taps are symmetric up to f32 summation order,
    # so allow the last bit)
    a = oracle.resize_lanczos3(img, 4, 5).astype(int)
    b = oracle.resize_lanczos3(img[:, ::-1], 4, 5)[:, ::-1].astype(int)
    assert np.abs(a - b).max() <= 1
    # a step edge overshoots (negative lobes) and is clamped to [0, 255]
    edge = np.zeros((8, 64, 3), np.uint8)
    edge[:, 32:] = 255
    up = oracle.resize_lanczos3(edge, 256, 8)
    assert up.min() == 0 and up.max() == 255 and (up[:, :100] == 0).all() and (up[:, 160:] == 255).all()
    # a view is the same as resizing the cropped copy
    big = rng.integers(0, 256, (40, 50, 3), dtype=np.uint8)
    assert (oracle.resize_lanczos3(big, 9, 8, view=(3, 5, 30, 20)) == oracle.resize_lanczos3(big[5:25, 3:33], 9, 8)).all()
    with pytest.raises(ValueError):
        oracle.resize_lanczos3(big, 9, 8, view=(30, 5, 30, 20))


def test_golden_vectors():
    g = np.load(os.path.join(GOLD, "resize_lanczos3.npz"))
    for k in range(int(g["cases"])):
        img, view, want = g[f"img{k}"], tuple(int(v) for v in g[f"view{k}"]), g[f"out{k}"]
        assert (oracle.resize_lanczos3(img, want.shape[1], want.shape[0], view) == want).all()
        x0, y0, cw, ch = view
        assert (onp.resize_lanczos3(img[y0:y0 + ch, x0:x0 + cw], want.shape[1], want.shape[0]) == want).all()


def _framed(rng, h, w, top, bottom, left, right, noise=False):
    img = np.full((h, w, 3), 255, np.uint8)
    img[top:h - bottom, left:w - right] = rng.integers(0, 230, (h - top - bottom, w - left - right, 3))
    if noise:  # a few ragged rows / columns: the MODE decides, not the extreme
        img[top + 3, :left + 4] = 255
        img[top:top + 2, left + 7] = 255
    return img


def test_prepare_view_formulations_agree():
    rng = np.random.default_rng(11)
    for k in range(30):
        h, w = int(rng.integers(20, 70)), int(rng.integers(20, 70))
        t, b, l, r = (int(v) for v in rng.integers(0, 6, 4))
        img = _framed(rng, h, w, t, b, l, r, noise=k % 2 == 1)
        for crop in (False, True):
            want = oracle.prepare_view(img, 8, crop)
            assert want == onp.prepare_view(img, 8, crop) == api.prepare_view(img, 8, crop)
    img = _framed(rng, 40, 60, 4, 6, 5, 3)
    # [first, last): the last non-white column / row is left out (utils.rs:160-161)
    assert oracle.prepare_view(img, 8, False) == (5, 4, 60 - 3 - 1 - 5, 40 - 6 - 1 - 4)
    x, y, s, s2 = oracle.prepare_view(img, 8, True)
    assert s == s2 == 29 and (x, y) == (5 + (51 - 29) // 2, 4)


def test_prepare_view_errors():
    white = np.full((20, 20, 3), 255, np.uint8)
    with pytest.raises(ValueError):
        oracle.prepare_view(white, 8, False)                      # utils.rs:157-158 asserts (mode of nothing is 0)
    with pytest.raises(api.EmosaicError, match="assertion failed"):
        api.prepare_view(white, 8, False)
    small = np.zeros((7, 30, 3), np.uint8)
    with pytest.raises(ValueError):
        oracle.prepare_view(small, 8, False)                      # utils.rs:99-106
    with pytest.raises(api.EmosaicError, match="smaller than the tile size"):
        api.prepare_view(small, 8, False)


def test_most_common_value():
    assert api._most_common_value(np.array([1, 2, 2, 3, 3, 3, 4])) == 3           # utils.rs:284-289
    assert api._most_common_value(np.array([], np.int64)) == 0                     # unwrap_or((0, 0)).0
    assert api._most_common_value(np.array([9, 4, 9, 4])) == 4                     # canonical tie: smallest value


def test_rotate_orientations():
    a = np.arange(2 * 3 * 3, dtype=np.uint8).reshape(2, 3, 3)
    assert (api.rotate(a, 1) == a).all() and (api.rotate(a, 9) == a).all()
    assert (api.rotate(a, 2) == a[:, ::-1]).all() and (api.rotate(a, 3) == a[::-1, ::-1]).all() and (api.rotate(a, 4) == a[::-1]).all()
    assert api.rotate(a, 6).shape == (3, 2, 3) and (api.rotate(a, 6)[0, 1] == a[0, 0]).all()   # clockwise: top-left -> top-right
    assert (api.rotate(a, 8)[2, 0] == a[0, 0]).all()                                               # counter-clockwise
    assert (api.rotate(a, 5) == a.transpose(1, 0, 2)).all()
    assert (api.rotate(api.rotate(a, 6), 8) == a).all() and (api.rotate(api.rotate(a, 7), 7) == a).all()


def test_source_dimension_rule_matches_oracle():
    for w, h, ds, dim in [(101, 103, 1, 2), (1023, 77, 2, 3), (4096, 4096, 1, 1), (50, 50, 7, 4), (5, 5, 1, 4)]:
        assert api.adjust_source_dims(w, h, ds, dim) == oracle.adjust_dims(w, h, ds, dim)


def test_oracle_tracks_exact_lanczos3():
    """The f32 two-pass restatement against the same filter evaluated in float64 with exact normalisation: the results may
    differ only where f32 rounding moves a value across a .5 boundary, i.e. by at most one level (checks the algorithm —
    window, weights, normalisation, clamp, rounding mode — independently of the f32 operation order)."""
    rng = np.random.default_rng(42)

    def exact_axis(n_in, n_out):
        ratio = n_in / n_out
        sratio = max(ratio, 1.0)
        taps = []
        for o in range(n_out):
            c = (o + 0.5) * ratio
            left = min(max(int(np.floor(c - 3 * sratio)), 0), n_in - 1)
            right = min(max(int(np.ceil(c + 3 * sratio)), left + 1), n_in)
            x = (np.arange(left, right) - (c - 0.5)) / sratio
            w = np.where(np.abs(x) < 3, np.sinc(x) * np.sinc(x / 3), 0.0)
            taps.append((left, w / w.sum()))
        return taps

    worst = 0
    for h, w, nh, nw in [(40, 52, 12, 12), (33, 29, 47, 61), (64, 64, 63, 62), (90, 70, 9, 7)]:
        yy, xx = np.mgrid[0:h, 0:w]
        img = np.clip(np.stack([np.sin(xx / 5.0) * 120 + 128, np.cos(yy / 4.0) * 120 + 128, (xx * 7 + yy * 3) % 256], -1)
                      + rng.integers(-10, 11, (h, w, 3)), 0, 255).astype(np.uint8)
        tmp = np.stack([sum(img[l + i].astype(np.float64) * wt for i, wt in enumerate(ws)) for l, ws in exact_axis(h, nh)])
        ex = np.stack([sum(tmp[:, l + i] * wt for i, wt in enumerate(ws)) for l, ws in exact_axis(w, nw)], axis=1)
        ex = np.floor(np.clip(ex, 0, 255) + 0.5)
        got = oracle.resize_lanczos3(img, nw, nh).astype(np.float64)
        diff = np.abs(got - ex)
        worst = max(worst, int(diff.max()))
        assert diff.max() <= 1 and (diff > 0).mean() < 0.02
    assert worst <= 1


@pytest.mark.parametrize("n_in,n_out", [(4097, 4096), (4098, 4096), (8192, 4096), (3000, 64), (2048, 64), (256, 8), (37, 36), (45, 44), (7, 31),
                                        (1, 5), (300, 1), (1024, 4096), (64, 64), (100, 99), (1000, 333)])
def test_library_tap_tables_equal_the_oracle(n_in, n_out):
    """The host side of libemosaic_cuda.so (resize_axis in resize.cu, exported as emo_resize_taps; no GPU involved) against the
    oracle's restatement of sample.rs: same windows, bit-identical f32 weights."""
    left, cnt, ws = api.resize_taps(n_in, n_out)
    ol, oc, ow = oracle.resize_axis(n_in, n_out)
    assert (left == ol).all() and (cnt == oc).all() and ws.shape == ow.shape
    assert (ws.view(np.uint32) == ow.view(np.uint32)).all()


def test_tap_windows_of_neighbouring_outputs_are_monotone():
    """resize_vertical2_kernel (resize.cu) gives two neighbouring output rows to one thread and walks the union of their windows
    once; that needs left[o] <= left[o + 1] and left[o] + cnt[o] <= left[o + 1] + cnt[o + 1] (the launcher checks it per geometry
    and falls back to one row per thread otherwise).  The windows are floor / ceil of monotone functions of o clamped to the
    axis, so this must hold for every geometry: shrink, stretch, near-identity, one-pixel axes."""
    rng = np.random.default_rng(12)
    geoms = [(4097, 4096), (8192, 4096), (3000, 64), (2048, 64), (256, 8), (7, 31), (1, 5), (300, 1), (1024, 4096), (64, 64), (65535, 3)]
    geoms += [(int(a), int(b)) for a, b in zip(rng.integers(1, 5000, 60), rng.integers(1, 3000, 60))]
    for n_in, n_out in geoms:
        left, cnt, _ = api.resize_taps(n_in, n_out)
        end = left.astype(np.int64) + cnt
        assert (np.diff(left.astype(np.int64)) >= 0).all() and (np.diff(end) >= 0).all(), (n_in, n_out)
        assert (cnt >= 1).all() and (end <= n_in).all()
