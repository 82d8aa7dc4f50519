"""The reference-shaped command line end to end on the GPU: cache build (-f), cache reuse, 1to1 / 4to1 / random,
tint.  The expected tiles (trim view, crop, Lanczos3 resize) and images come from the CPU oracle."""
import os

import numpy as np
import pytest

import oracle
from emosaic_b200 import cache, cli

pytestmark = pytest.mark.gpu
PIL = pytest.importorskip("PIL.Image")


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    d = tmp_path_factory.mktemp("cli")
    tiles = d / "tiles"
    (tiles / "sub").mkdir(parents=True)
    rng = np.random.default_rng(0)
    for i in range(60):
        base = rng.integers(0, 256, 3)
        img = np.clip(base + rng.integers(-30, 31, (40, 52, 3)), 0, 255).astype(np.uint8)
        img[:, :20] = np.clip(img[:, :20].astype(int) - 60, 0, 255)  # left/right asymmetry -> mirrored matches in 4to1
        img = np.minimum(img, 225)                                    # nothing "white" (> 240 after JPEG ringing) inside ...
        if i % 4 == 0:
            img[:3 + i % 5] = img[-4:] = img[:, :5] = img[:, -2 - i % 3:] = 255   # ... but a white frame on some (utils.rs:93-167)
        p = tiles / ("sub" if i % 3 == 0 else "") / f"t{i:03d}.{'jpg' if i % 2 else 'jpeg'}"
        PIL.fromarray(img).save(p, quality=95)
    (tiles / "notes.txt").write_text("not an image")
    PIL.fromarray(rng.integers(0, 256, (8, 8, 3), dtype=np.uint8)).save(tiles / "ignored.png")
    src = rng.integers(0, 256, (36, 48, 3), dtype=np.uint8)
    PIL.fromarray(src).save(d / "src.png")
    return d, src


def oracle_tile(path, ts, crop):
    """prepare_tile (tiles/utils.rs:63-196) on the CPU oracle: decode, trim view / crop, Lanczos3 resize (no EXIF here)."""
    img = np.asarray(PIL.open(path).convert("RGB"), dtype=np.uint8)
    return oracle.resize_lanczos3(img, ts, ts, oracle.prepare_view(img, ts, crop))


def expected(tiles_dir, src, dim, ts, crop):
    paths = cli.find_images(str(tiles_dir), {"jpg", "jpeg"})
    px_an = np.stack([oracle_tile(p, ts, crop) for p in paths])
    px_rd = np.stack([oracle_tile(p, ts, True) for p in paths])
    colors = oracle.analyse_tiles(px_an, dim * dim)
    item, dist = oracle.match(colors, src)
    return paths, colors, item, dist, px_an, px_rd


@pytest.mark.parametrize("mode,dim", [("1", 1), ("4to1", 2), ("3", 3)])
def test_cli_mosaic_and_cache(workdir, mode, dim):
    d, src = workdir
    ts = 12
    out = d / f"out_{mode}.png"
    cpath = d / "tiles" / cache.cache_file_name(dim * dim, False)
    if cpath.exists():
        cpath.unlink()
    argv = ["-s", str(ts), "-o", str(out), str(d / "src.png"), "mosaic", str(d / "tiles"), "-m", mode]
    assert cli.main(argv + ["-f"]) == 0
    paths, colors, item, dist, px_an, px_rd = expected(d / "tiles", src, dim, ts, False)
    # analysis uses the un-cropped tiles, rendering always re-prepares with crop = true (tileset.rs:152-155)
    assert (np.asarray(PIL.open(out)) == oracle.render(px_rd, item)).all()
    # cache bytes are the reference's bincode layout
    assert cpath.read_bytes() == oracle.cache_serialize(colors, np.arange(1, len(paths) + 1), [None] * len(paths), paths)
    assert os.path.exists(str(out)[:-4] + ".stats.png")
    # second run reuses the cache (tiles re-prepared with crop = true, tileset.rs:152-155)
    assert cli.main(argv) == 0
    assert (np.asarray(PIL.open(out)) == oracle.render(px_rd, item)).all()
    # a vanished tile is filtered and the rest renumbered (main.rs:624-654)
    victim = paths[5]
    os.rename(victim, victim + ".bak")
    try:
        assert cli.main(argv) == 0
        keep = [i for i in range(len(paths)) if i != 5]
        it2, _ = oracle.match(colors[keep], src)
        assert (np.asarray(PIL.open(out)) == oracle.render(px_rd[keep], it2)).all()
    finally:
        os.rename(victim + ".bak", victim)


def test_cli_tint_and_random(workdir):
    d, src = workdir
    ts = 8
    out = d / "tint.png"
    assert cli.main(["-s", str(ts), "-o", str(out), "--crop", str(d / "src.png"), "mosaic", str(d / "tiles"), "-t", "0.5", "-f"]) == 0
    paths, colors, item, dist, px_an, px_rd = expected(d / "tiles", src, 1, ts, True)
    got = np.asarray(PIL.open(out))
    assert got.shape == (36 * ts, 48 * ts, 4)
    assert (got == oracle.tint(oracle.render(px_an, item), src, 127)).all()
    assert (d / "tiles" / ".emosaic_1to1_cropped").exists()
    out2 = d / "rand.png"
    assert cli.main(["-s", str(ts), "-o", str(out2), str(d / "src.png"), "mosaic", str(d / "tiles"), "-m", "random", "--seed", "3"]) == 0
    assert np.asarray(PIL.open(out2)).shape == (36 * ts, 48 * ts, 3)  # mod.rs:48-57


def test_cli_rejects(workdir):
    d, _ = workdir
    base = ["-s", "8", str(d / "src.png"), "mosaic", str(d / "tiles")]
    assert cli.main(base + ["--no-repeat", "--greedy"]) == 2     # render_nto1's rayon-order dependent branch (main.rs:663-667)
    assert cli.main(base + ["--randomize", "5"]) == 2
    assert cli.main(base + ["--no-repeat"]) == 1                  # 36 x 48 blocks for 60 tiles: "Insufficient tiles"
    assert cli.main(["-s", "9", str(d / "src.png"), "mosaic", str(d / "tiles"), "-m", "2"]) == 1  # tile size % dim
    assert cli.main(["-s", "8", str(d / "src.png"), "mosaic", str(d / "nope")]) == 1
    assert cli.main(["-s", "0", str(d / "src.png"), "mosaic", str(d / "tiles")]) == 1      # main.rs:272-283
    assert cli.main(["-s", "2048", str(d / "src.png"), "mosaic", str(d / "tiles")]) == 1
    assert cli.main(["-s", "8", str(d / "missing.png"), "mosaic", str(d / "tiles")]) == 1
    assert cli.main(["-s", "8", "-o", str(d / "no_dir" / "o.png"), str(d / "src.png"), "mosaic", str(d / "tiles")]) == 1


def test_cli_tint_uses_original_image_as_overlay(workdir):
    """--downsample 2 with -t: matching runs on the resized source, the tint overlay is the image as opened
    (main.rs:396-398 / :447-466), nearest-resized to the output."""
    d, src = workdir
    ts = 8
    out = d / "tint_ds.png"
    assert cli.main(["-s", str(ts), "-o", str(out), "--crop", str(d / "src.png"), "mosaic", str(d / "tiles"), "-t", "0.5",
                     "--downsample", "2", "-f"]) == 0
    from emosaic_b200 import api
    nw, nh = api.adjust_source_dims(src.shape[1], src.shape[0], 2, 1)
    small = oracle.resize_lanczos3(src, nw, nh)  # main.rs:595
    paths = cli.find_images(str(d / "tiles"), {"jpg", "jpeg"})
    px = np.stack([oracle_tile(p, ts, True) for p in paths])
    colors = oracle.analyse_tiles(px, 1)
    item, _ = oracle.match(colors, small)
    want = oracle.tint(oracle.render(px, item), src, 127)
    got = np.asarray(PIL.open(out))
    assert got.shape == (nh * ts, nw * ts, 4) and (got == want).all()


def test_cli_no_repeat(workdir):
    """--no-repeat (main.rs:663-664 -> render_nto1_no_repeat): a 6 x 8 block source for the 60 tiles; every tile at most once."""
    from oracle import oracle_np as onp
    d, src = workdir
    ts = 12
    small = src[:12, :16]
    PIL.fromarray(small).save(d / "small.png")
    out = d / "out_nr.png"
    assert cli.main(["-s", str(ts), "-o", str(out), str(d / "small.png"), "mosaic", str(d / "tiles"), "-m", "2", "--no-repeat", "-f"]) == 0
    paths, colors, _, _, px_an, px_rd = expected(d / "tiles", small, 2, ts, False)
    item, dist = onp.no_repeat_assign(colors, small)
    assert (np.asarray(PIL.open(out)) == oracle.render(px_rd, item)).all()
    assert len(set(np.abs(item).reshape(-1).tolist())) == item.size


def test_cli_prepare_subcommand(workdir):
    """`emosaic -s N [--crop] IMG prepare` (main.rs:380-386): one prepared tile, saved to the output path."""
    d, _ = workdir
    photo = cli.find_images(str(d / "tiles"), {"jpg", "jpeg"})[0]   # t000: has a white frame
    for crop in (False, True):
        out = d / f"prepared_{int(crop)}.png"
        assert cli.main(["-s", "12", "-o", str(out)] + (["--crop"] if crop else []) + [photo, "prepare"]) == 0
        assert (np.asarray(PIL.open(out)) == oracle_tile(photo, 12, crop)).all()


def test_cli_skips_unreadable_and_undersized_tiles(workdir, capsys):
    """generate_tile_set (main.rs:757-806): a corrupt file and a photo smaller than the tile size are listed under
    'Failed to read the following images(n)' and left out; the survivors are numbered 1..n in walk order; the no-repeat
    statistics image has one pixel per block (output coordinates, rendering.rs:352-365)."""
    d, src = workdir
    bad = d / "tiles_bad"
    bad.mkdir()
    rng = np.random.default_rng(5)
    good = []
    for i in range(30):
        img = np.clip(rng.integers(0, 226, 3) + rng.integers(-20, 21, (30, 30, 3)), 0, 225).astype(np.uint8)
        PIL.fromarray(img).save(bad / f"g{i:02d}.jpg", quality=95)
        good.append(str(bad / f"g{i:02d}.jpg"))
    (bad / "broken.jpg").write_bytes(b"\xff\xd8\xff\xe0 this is not a jpeg")
    PIL.fromarray(rng.integers(0, 200, (5, 40, 3), dtype=np.uint8)).save(bad / "tiny.jpg")      # 5 rows < tile size 8
    small = src[:4, :6]
    PIL.fromarray(small).save(d / "small_bad.png")
    out = d / "out_bad.png"
    assert cli.main(["-s", "8", "-o", str(out), str(d / "small_bad.png"), "mosaic", str(bad), "-f", "--no-repeat"]) == 0
    err = capsys.readouterr().err
    assert "Failed to read the following images(2):" in err and "- broken.jpg" in err and "- tiny.jpg" in err
    assert "Tile set with 30 tiles" in err
    from oracle import oracle_np as onp
    px = np.stack([oracle_tile(p, 8, False) for p in sorted(good)])
    px_rd = np.stack([oracle_tile(p, 8, True) for p in sorted(good)])
    colors = oracle.analyse_tiles(px, 1)
    item, dist = onp.no_repeat_assign(colors, small)
    assert (np.asarray(PIL.open(out)) == oracle.render(px_rd, item)).all()
    stats_img = np.asarray(PIL.open(str(out)[:-4] + ".stats.png"))
    assert stats_img.shape[:2] == (4, 6)                                  # one pixel per block
    md = float(dist.max())
    assert (stats_img[..., 0] == (dist / md * 255.0).astype(np.uint8)).all()
    paths = cache.deserialize_tile_set((bad / ".emosaic_1to1").read_bytes(), 1)[1]
    assert [os.path.basename(p) for p in paths] == [f"g{i:02d}.jpg" for i in range(30)]
