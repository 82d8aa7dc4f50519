"""Host-side logic mirrored from the reference (no GPU): cache format, dimension rules, coords, sharding."""
import numpy as np
import pytest

import emosaic_b200 as emo
import oracle
from emosaic_b200 import sharding


def test_flipped_coords_matches_reference_vector():
    # tiles/utils.rs:302-308
    c = list(range(1, 13))
    f = emo.flipped_coords(c)
    assert f.tolist() == [4, 5, 6, 1, 2, 3, 10, 11, 12, 7, 8, 9]
    assert emo.flipped_coords(f).tolist() == c
    for n in (3, 12, 27, 48, 75):
        v = np.arange(n)
        assert (emo.flipped_coords(v) == oracle.flipped_coords(v)).all()


def test_tile_coords():
    # tiles/tile.rs:127-140
    assert emo.Tile.from_colors([[1, 2, 3]]).coords().tolist() == [1, 2, 3]
    t = emo.Tile.from_colors([[1, 2, 3], [4, 5, 6], [7, 8, 9], [10, 11, 12]])
    assert t.coords().tolist() == list(range(1, 13))
    t.flipped = True
    assert t.coords().tolist() == [4, 5, 6, 1, 2, 3, 10, 11, 12, 7, 8, 9]


def test_get_img_colors():
    # analysis.rs:58-71
    img = np.zeros((4, 4, 3), np.uint8)
    for y in range(4):
        for x in range(4):
            img[y, x] = (x * 64, y * 64, 128)
    assert emo.get_img_colors(0, 0, 2, img, 4).tolist() == [[0, 0, 128], [64, 0, 128], [0, 64, 128], [64, 64, 128]]


def test_tile_set_container():
    # mod.rs:26-46
    ts = emo.TileSet(N=1)
    assert len(ts) == 0
    ts.push_tile("a.jpg", [[1, 2, 3]])
    assert len(ts) == 1
    assert ts.get_tile(1).idx == 1 and not ts.get_tile(1).flipped
    assert ts.get_tile(-1).flipped
    assert ts.get_tile(2) is None and ts.get_tile(0) is None


@pytest.mark.parametrize("w,h,ds,dim", [(100, 100, 1, 1), (101, 103, 1, 2), (1023, 77, 2, 3), (640, 481, 1, 4), (17, 19, 1, 8)])
def test_adjust_dims(w, h, ds, dim):
    # main.rs:567-587
    assert emo.adjust_source_dims(w, h, ds, dim) == oracle.adjust_dims(w, h, ds, dim)
    nw, nh = emo.adjust_source_dims(w, h, ds, dim)
    assert nw % dim == 0 and nh % dim == 0


def test_tint_alpha():
    for t in (0.0, 0.001, 0.25, 0.5, 0.75, 0.999, 1.0, 2.0):
        assert emo.tint_alpha(t) == oracle.tint_alpha(t)


def test_cache_file_name():
    # main.rs:597-601
    assert emo.cache_file_name(1, False) == ".emosaic_1to1"
    assert emo.cache_file_name(4, True) == ".emosaic_4to1_cropped"


def test_cache_bytes_match_oracle_and_roundtrip(tmp_path):
    rng = np.random.default_rng(3)
    for N in (1, 4):
        T = 37
        colors = rng.integers(0, 256, (T, N, 3), dtype=np.uint8)
        paths = [str(tmp_path / f"dir{t % 3}" / f"img_{t}.{'jpg' if t % 5 else 'png'}") for t in range(T)]
        dates = [None if t % 4 == 0 else f"2021:0{1 + t % 9}:1{t % 10}" for t in range(T)]
        blob = emo.serialize_tile_set(colors, paths, dates)
        assert blob == oracle.cache_serialize(colors, np.arange(1, T + 1), dates, paths)
        # hand-checked prefix: u64 T, u64 3N, colours, u16 idx=1, tag
        assert blob[:8] == T.to_bytes(8, "little") and blob[8:16] == (3 * N).to_bytes(8, "little")
        assert blob[16:16 + 3 * N] == colors[0].tobytes()
        assert blob[16 + 3 * N:18 + 3 * N] == (1).to_bytes(2, "little") and blob[18 + 3 * N] == 0
        c2, p2, d2 = emo.deserialize_tile_set(blob, N)
        assert (c2 == colors).all() and p2 == paths and d2 == dates
        # main.rs:624-654: drop missing files / wrong extensions, keep order (renumbering is positional)
        c3, p3, d3 = emo.deserialize_tile_set(blob, N, extensions={"jpg"})
        keep = [t for t in range(T) if t % 5]
        assert (c3 == colors[keep]).all() and p3 == [paths[t] for t in keep]
        for t in (1, 2, 7):
            (tmp_path / f"dir{t % 3}").mkdir(exist_ok=True)
            open(paths[t], "wb").close()
        c4, p4, _ = emo.deserialize_tile_set(blob, N, extensions={"jpg", "png"}, check_exists=True)
        assert p4 == [paths[t] for t in (1, 2, 7)] and (c4 == colors[[1, 2, 7]]).all()
    with pytest.raises(ValueError):
        emo.deserialize_tile_set(blob[:-3], 4)
    with pytest.raises(ValueError):
        emo.deserialize_tile_set(blob, 1)  # wrong N: try_into().unwrap() in the reference


def test_cache_idx_wraps_like_reference():
    # main.rs:791 `(idx + 1) as u16`
    T = 65537
    colors = np.zeros((T, 1, 3), np.uint8)
    blob = emo.serialize_tile_set(colors, ["p"] * T)
    rec = 8 + 3 + 2 + 1
    off = 8 + (65535) * rec + 8 + 3
    assert blob[off:off + 2] == (0).to_bytes(2, "little")


@pytest.mark.parametrize("n,world", [(4096, 1), (4096, 8), (4095, 8), (7, 8), (100, 3)])
def test_stripes_partition(n, world):
    st = sharding.all_stripes(n, world)
    assert st[0][0] == 0 and st[-1][1] == n
    for (a, b), (c, d) in zip(st, st[1:]):
        assert b == c and 0 <= (b - a) - (d - c) <= 1


def test_render_nto1_rejects_like_reference():
    ts = emo.TileSet(N=4)
    ts.push_tile_with_image("x", np.zeros((4, 3)), np.zeros((8, 8, 3)))
    with pytest.raises(emo.EmosaicError, match="Dimensions must be divisible by 2"):
        emo.render_nto1(np.zeros((5, 4, 3), np.uint8), ts, 8)
    with pytest.raises(emo.EmosaicError, match="Tile size must be divisible by 2"):
        emo.render_nto1(np.zeros((4, 4, 3), np.uint8), ts, 7)
    with pytest.raises(emo.EmosaicError):
        emo.render_nto1(np.zeros((4, 4, 3), np.uint8), ts, 8, no_repeat=True)


def test_stats_summary_and_image():
    # stats.rs:87-195 from the (item, dist) maps
    import io
    from emosaic_b200 import stats
    item = np.array([[1, -2, 1], [3, 1, 2]])
    dist = np.array([[5, 9, 0], [7, 7, 1]], dtype=np.uint32)
    buf = io.StringIO()
    s = stats.summarise(item, dist, ["a", "b", "c"], file=buf)
    assert s["total"] == 6 and s["unique"] == 3 and abs(s["average_distance"] - 29 / 6) < 1e-12
    assert s["top"][0] == ("a", 3) and s["worst"][0] == ("b", 9)
    assert "Average color distance: 4.833" in buf.getvalue()
    img = stats.render(dist, 1, 1)
    assert img[:, :, 0].tolist() == [[141, 255, 0], [198, 198, 28]]  # (d / max_d * 255) as u8
    assert stats.render(np.zeros((2, 2), np.uint32), 1, 1).max() == 0
    assert stats.render(dist, 2, 4).shape == (1, 2, 3)  # source coordinates / tile_size, stats.rs:176-178
    with pytest.raises(ValueError):
        stats.render(np.zeros((0, 0), np.uint32), 1, 1)


def test_cli_parser_accepts_both_spellings():
    from emosaic_b200 import cli
    p = cli.build_parser()
    a = p.parse_args(["-s", "8", "img.png", "mosaic", "tiles", "-m", "4to1", "-t", "0.5", "-f"])
    assert cli.MODES[a.mode] == 2 and a.force and a.tint_opacity == 0.5 and a.tile_size == 8
    a = p.parse_args(["img.png", "mosaic", "tiles", "--mode", "128"])
    assert cli.MODES[a.mode] == 128 and a.tile_size == 16 and a.output_path == "./output.jpg"


def test_cache_golden_bytes():
    """A committed .emosaic_4to1 blob: the Python serialiser reproduces it byte for byte and the loader reads it back."""
    import os
    here = os.path.join(os.path.dirname(__file__), "golden")
    blob = open(os.path.join(here, "cache_4to1.bin"), "rb").read()
    meta = np.load(os.path.join(here, "cache_4to1_meta.npz"))
    paths = [str(p) for p in meta["paths"]]
    dates = [str(d) or None for d in meta["dates"]]
    assert emo.serialize_tile_set(meta["colors"], paths, dates) == blob
    c, p, d = emo.deserialize_tile_set(blob, 4)
    assert (c == meta["colors"]).all() and p == paths and d == dates
    # hand-decoded header: 5 tiles, 12 colour bytes each
    assert blob[:8] == (5).to_bytes(8, "little") and blob[8:16] == (12).to_bytes(8, "little")
    assert len(blob) == 8 + 5 * (8 + 12 + 2 + 1) + 2 * (8 + 10) + 8 + sum(8 + len(x.encode()) for x in paths)


def test_cache_corrupt_lengths_fail_fast():
    """A cache file with absurd length fields (tile count, date length, path length) is rejected with ValueError before any
    allocation or out-of-range read — the reference's bincode returns Err and the tile set is re-analysed (main.rs:617-623)."""
    import struct
    from emosaic_b200 import cache
    colors = np.arange(6, dtype=np.uint8).reshape(2, 1, 3)
    blob = bytearray(cache.serialize_tile_set(colors, ["/tmp/a.jpg", "/tmp/b.jpg"], ["2020:01:01", None]))
    ok = cache.deserialize_tile_set(bytes(blob), 1)
    assert ok[1] == ["/tmp/a.jpg", "/tmp/b.jpg"] and ok[2] == ["2020:01:01", None]
    huge = struct.pack("<Q", 0xFFFFFFFFFFFFFFFF)
    for off in (0,                      # tile count
                8 + 8 + 3 + 2 + 1,      # length of the first date string
                len(blob) - 10 - 8):    # length of the last path
        bad = bytearray(blob)
        bad[off:off + 8] = huge
        with pytest.raises(ValueError):
            cache.deserialize_tile_set(bytes(bad), 1)
    with pytest.raises(ValueError):
        cache.deserialize_tile_set(bytes(blob[:-3]), 1)
    with pytest.raises(ValueError):
        cache.deserialize_tile_set(bytes(blob), 4)      # wrong vector length for this mode
