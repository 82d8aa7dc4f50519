"""The claim behind the 1to1 search index (emosaic_b200/csrc/index.cu), checked on the CPU: a separable forward/backward
min-plus sweep over keys `dist << 22 | tile` yields, for every colour of the cube, exactly the (distance, tile) a brute-force
scan with the reference's strict-`<` leaf scan keeps (kiddo nearest_one::<Manhattan>, rendering.rs:187-195; tie-break of
DESIGN.md §2).  Runs on a reduced cube (5 bits per channel) so that brute force over all cells takes milliseconds."""
import numpy as np
import pytest

import oracle
from oracle import oracle_np as onp


def brute(colors, S):
    r, g, b = np.meshgrid(np.arange(S), np.arange(S), np.arange(S), indexing="ij")
    cells = np.stack([r, g, b], -1).reshape(-1, 3).astype(np.int64)          # index = r*S*S + g*S + b
    c = colors.reshape(-1, 3).astype(np.int64)
    d = np.abs(cells[:, None, :] - c[None, :, :]).sum(-1)                     # [cells, T]
    best = d.argmin(1)                                                         # first minimum = smallest tile index
    return best.reshape(S, S, S), d.min(1).reshape(S, S, S)                    # [r][g][b]


@pytest.mark.parametrize("T,seed", [(1, 0), (2, 1), (5, 2), (40, 3), (300, 4), (2000, 5)])
def test_table_equals_bruteforce_on_a_small_cube(T, seed):
    bits, S = 5, 32
    rng = np.random.default_rng(seed)
    colors = rng.integers(0, S, (T, 1, 3), dtype=np.uint8)
    if T >= 5:
        colors[3] = colors[0]          # duplicate colours: the smaller index must win
        colors[4] = colors[1]
    tile, dist = onp.l1_voronoi_table(colors, bits)                           # [b][g][r]
    bt, bd = brute(colors, S)                                                  # [r][g][b]
    assert (np.transpose(dist, (2, 1, 0)) == bd).all()
    assert (np.transpose(tile, (2, 1, 0)) == bt).all()


def test_table_ties_half_way():
    # two tiles at equal distance from every cell of the plane between them: smallest index everywhere on the plane
    colors = np.array([[[20, 10, 10]], [[10, 10, 10]], [[10, 20, 10]]], np.uint8)
    tile, dist = onp.l1_voronoi_table(colors, 5)
    bt, bd = brute(colors, 32)
    assert (np.transpose(tile, (2, 1, 0)) == bt).all() and (np.transpose(dist, (2, 1, 0)) == bd).all()
    assert tile[10, 10, 15] == 0 and dist[10, 10, 15] == 5                    # (r=15,g=10,b=10): tiles 0 and 1 tie, 0 wins


def test_table_agrees_with_the_c_oracle_at_8_bits():
    """Full 8-bit cube against the C oracle's brute-force match on a sample of colours."""
    rng = np.random.default_rng(9)
    colors = rng.integers(0, 256, (700, 1, 3), dtype=np.uint8)
    colors[10] = colors[2]
    tile, dist = onp.l1_voronoi_table(colors, 8)
    src = rng.integers(0, 256, (64, 64, 3), dtype=np.uint8)
    src[0, :8] = colors[:8, 0]
    ri, rd = oracle.match(colors, src)
    assert (tile[src[..., 2], src[..., 1], src[..., 0]] + 1 == ri).all()
    assert (dist[src[..., 2], src[..., 1], src[..., 0]] == rd).all()
