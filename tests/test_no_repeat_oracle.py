"""The no-repeat renderer's oracle (oracle_np.no_repeat_assign; reference render_nto1_no_repeat, rendering.rs:262-401).

Pinned by the reference's own test: mod.rs:118-145 runs render_nto1_no_repeat on the black-and-white universe — every
member renders to itself, 1x2 stacks of distinct members render exactly.  The heap formulation is checked against a
literal restatement of the reference's sorted-vector / pop / binary-insert loop on inputs without competing equal
distances (where the reference's order is implementation-defined)."""
import itertools

import numpy as np
import pytest

import oracle
from oracle import oracle_np as onp


def universe(N):
    """mod.rs:83-106: all black/white dim x dim images except the all-white one."""
    dim = int(N ** 0.5)
    out = []
    for bits in itertools.product([0, 255], repeat=N):
        if all(b == 255 for b in bits):
            continue
        out.append(np.repeat(np.array(bits, np.uint8).reshape(dim, dim, 1), 3, axis=2))
    return np.stack(out)


@pytest.mark.parametrize("N", [1, 4, 9])
def test_universe_no_repeat(N):
    dim = int(N ** 0.5)
    uni = universe(N)
    colors = oracle.analyse_tiles(uni, N)
    limit = 64
    for img in uni[:limit]:
        item, dist = onp.no_repeat_assign(colors, img)
        assert (dist == 0).all()
        assert (oracle.render(uni, item) == img).all()
    for a in range(0, min(len(uni) - 1, limit), 2):   # 1x2 stacks of consecutive members (mod.rs:131-145)
        img = np.concatenate([uni[a], uni[a + 1]], 0)
        item, dist = onp.no_repeat_assign(colors, img)
        assert (dist == 0).all() and abs(int(item[0, 0])) != abs(int(item[1, 0]))
        assert (oracle.render(uni, item) == img).all()


def test_heap_equals_the_literal_loop():
    rng = np.random.default_rng(0)
    checked = 0
    for trial in range(90):
        N = [1, 4, 9][trial % 3]
        dim = int(N ** 0.5)
        T, bh, bw = int(rng.integers(6, 40)), int(rng.integers(1, 5)), int(rng.integers(1, 6))
        if bh * bw > 2 * T:
            continue
        colors = rng.integers(0, 256, (T, N, 3), dtype=np.uint8)
        src = rng.integers(0, 256, (bh * dim, bw * dim, 3), dtype=np.uint8)
        i1, d1, ties = onp.no_repeat_assign(colors, src, True)
        if ties:
            continue
        i2, d2 = onp.no_repeat_assign_literal(colors, src)
        assert (i1 == i2).all() and (d1 == d2).all()
        checked += 1
    assert checked >= 40


def test_properties_and_errors():
    rng = np.random.default_rng(3)
    colors = rng.integers(0, 256, (30, 4, 3), dtype=np.uint8)
    src = rng.integers(0, 256, (10, 10, 3), dtype=np.uint8)       # 25 blocks <= 30 tiles: everything placed, no tile twice
    item, dist = onp.no_repeat_assign(colors, src)
    assert (item != 0).all() and len(set(np.abs(item).reshape(-1).tolist())) == 25
    q = onp.queries(src, 4).reshape(-1, 12).astype(np.int64)
    for blk, it in enumerate(item.reshape(-1)):                    # dist is the L1 distance to the placed orientation
        v = colors[abs(it) - 1].reshape(12).astype(np.int64)
        if it < 0:
            v = onp.mirror(v[None], 4)[0]
        assert np.abs(q[blk] - v).sum() == dist.reshape(-1)[blk]
    first_item, _ = oracle.match(colors, src)                      # the globally nearest pair is always granted
    best = np.unravel_index(np.argmin(dist), dist.shape)
    assert item[best] == first_item[best]
    # 15 tiles for 25 blocks: allowed (25 <= 2 * 15), 10 blocks stay unplaced — a tile serves once in either orientation
    item, dist = onp.no_repeat_assign(colors[:15], src)
    assert (item == 0).sum() == 10 and len(set(np.abs(item[item != 0]).tolist())) == 15
    with pytest.raises(AssertionError, match="Insufficient tiles for no-repeat mode: need 25 tiles but only have 24"):
        onp.no_repeat_assign(colors[:12], src)


@pytest.mark.parametrize("N,T,bh,bw,quant", [(1, 400, 12, 15, 1), (4, 150, 8, 9, 1), (9, 60, 4, 5, 1), (4, 90, 9, 10, 64), (1, 30, 6, 10, 1),
                                             (4, 40, 8, 10, 32), (16, 25, 3, 4, 1), (1, 2000, 40, 45, 8)])
def test_c_oracle_equals_numpy_oracle(N, T, bh, bw, quant):
    """The C restatement (oracle.no_repeat_assign: counting-sorted lists + heap merge, sized for 120 x 120 blocks) and the
    independent numpy one agree, including quantised colours (competing equal distances) and libraries too small to fill the
    image (T < blocks <= 2T: the rest stays unplaced)."""
    dim = int(N ** 0.5)
    rng = np.random.default_rng(N * 100 + T)
    colors = (rng.integers(0, 256, (T, N, 3)) // quant * quant).astype(np.uint8)
    src = (rng.integers(0, 256, (bh * dim, bw * dim, 3)) // quant * quant).astype(np.uint8)
    ci, cd = oracle.no_repeat_assign(colors, src)
    ni, nd = onp.no_repeat_assign(colors, src)
    assert (ci == ni).all() and (cd == nd).all()
    with pytest.raises(AssertionError, match="Insufficient tiles"):
        oracle.no_repeat_assign(colors[:1], src)
