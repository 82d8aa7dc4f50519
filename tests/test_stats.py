"""The reference's own RenderStats tests (src/mosaic/stats.rs:219-311) restated against emosaic_b200.stats, plus the
vectorised map path the renderers use (one entry per block of the GPU's item / dist maps).  CPU only."""
import io

import numpy as np
import pytest

from emosaic_b200 import stats
from emosaic_b200.stats import RenderStats


def test_render_stats_new_and_default():            # stats.rs:223-233
    assert RenderStats().tile_count() == 0


def test_push_tile():                               # stats.rs:235-245
    s = RenderStats()
    s.push_tile(10, 20, 1, 100)
    assert s.tile_count() == 1
    s.push_tile(30, 40, 1, 200)
    assert s.tile_count() == 2
    s.push_tile(30, 40, 1, 5)                       # HashMap::insert on the same key replaces (stats.rs:63)
    assert s.tile_count() == 2 and s.tiles[(30, 40)] == (1, 5)


def test_summarise_empty():                         # stats.rs:247-254: prints the empty message, no panic
    f = io.StringIO()
    assert RenderStats().summarise([], file=f) == {}
    assert f.getvalue() == "No tiles recorded in statistics\n"


def test_summarise_with_tiles():                    # stats.rs:256-278
    s = RenderStats()
    paths = ["test1.jpg", "test2.jpg"]
    s.push_tile(0, 0, 1, 10)
    s.push_tile(10, 10, 2, 20)
    s.push_tile(20, 20, 1, 15)                      # tile 1 again
    assert s.tile_count() == 3
    f = io.StringIO()
    out = s.summarise(paths, file=f)
    assert out["total"] == 3 and out["unique"] == 2
    assert out["average_distance"] == pytest.approx(15.0)
    assert out["top"] == [("test1.jpg", 2), ("test2.jpg", 1)]
    assert out["worst"] == [("test2.jpg", 20), ("test1.jpg", 15), ("test1.jpg", 10)]
    txt = f.getvalue()
    assert "Mosaic Statistics:" in txt and "  Total tiles placed: 3" in txt and "  Unique images used: 2" in txt
    assert "  Average color distance: 15.000" in txt            # {:.3}
    assert "  1. test1.jpg (2 times)" in txt and "  1. test2.jpg (distance: 20)" in txt


def test_render_empty_panic():                      # stats.rs:280-285
    with pytest.raises(ValueError, match="Cannot render visualization: no tiles recorded"):
        RenderStats().render(16)


def test_render_zero_tile_size_panic():             # stats.rs:287-295
    s = RenderStats()
    s.push_tile(0, 0, 1, 100)
    with pytest.raises(ValueError, match="Tile size must be greater than 0"):
        s.render(0)


def test_render_basic():                            # stats.rs:297-317
    s = RenderStats()
    s.push_tile(0, 0, 1, 50)
    s.push_tile(16, 16, 1, 150)
    img = s.render(16)
    assert img.shape == (2, 2, 3)
    assert img[0, 0, 0] < img[1, 1, 0]              # lower distance = darker
    assert img[0, 0, 0] == int(50 / 150 * 255.0) and img[1, 1, 0] == 255
    assert (img[0, 0] == img[0, 0, 0]).all()        # grey


def test_flipped_tiles_count_as_their_tile():
    """push_tile keeps tile.idx (unsigned) and the flipped flag separately (stats.rs:57-62): +k and -k are one image."""
    out = stats.summarise(np.array([3, -3, 2]), np.array([1, 2, 3]), ["a", "b", "c"], file=io.StringIO())
    assert out["unique"] == 2 and out["top"][0] == ("c", 2)


@pytest.mark.parametrize("dim,ts", [(1, 1), (2, 2), (1, 8), (2, 16), (3, 12)])
def test_map_path_equals_entry_path(dim, ts):
    """stats.render / stats.summarise over whole maps == RenderStats.from_maps(...) entry by entry (the reference's loop),
    for source-coordinate keys (render_nto1, step = dim) and output-coordinate keys (no-repeat, step = tile size)."""
    rng = np.random.default_rng(dim * 100 + ts)
    item = rng.integers(1, 40, (9, 7)).astype(np.int32) * rng.choice([-1, 1], (9, 7))
    dist = rng.integers(0, 3000, (9, 7)).astype(np.uint32)
    for step in (dim, ts):
        want = RenderStats.from_maps(item, dist, step).render(ts)
        got = stats.render(dist, step, ts)
        assert got.shape == want.shape and (got == want).all()
    a = RenderStats.from_maps(item, dist, dim).summarise(None, file=io.StringIO())
    b = stats.summarise(item, dist, None, file=io.StringIO())
    assert a["total"] == b["total"] and a["unique"] == b["unique"] and a["average_distance"] == b["average_distance"]
    assert sorted(c for _, c in a["top"]) == sorted(c for _, c in b["top"])
    assert [d for _, d in a["worst"]] == [d for _, d in b["worst"]]


def test_unplaced_blocks_have_no_entry():
    """A no-repeat block that ran out of tiles (item 0) is not recorded (rendering.rs:347-365): it does not count, does not
    set the normalisation maximum and does not extend the image."""
    item = np.array([[1, 2, 0], [3, 0, 0]], np.int32)
    dist = np.array([[10, 20, 999], [30, 999, 999]], np.uint32)
    placed = item != 0
    img = stats.render(dist, 4, 4, placed)
    assert img.shape == (2, 2, 3)                   # the all-unplaced last column is outside the bounds
    assert img[1, 0, 0] == 255 and img[0, 1, 0] == int(20 / 30 * 255.0) and img[1, 1, 0] == 0
    assert (img == RenderStats.from_maps(item, dist, 4).render(4)).all()
    out = stats.summarise(item[placed], dist[placed], None, file=io.StringIO())
    assert out["total"] == 3 and out["average_distance"] == 20.0
    with pytest.raises(ValueError, match="no tiles recorded"):
        stats.render(dist, 4, 4, np.zeros_like(placed))
