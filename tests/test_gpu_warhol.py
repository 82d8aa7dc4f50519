"""The reference's one real image fixture, example/warhol.png (1600 x 1600 palette PNG; used by its own
src/mosaic/tiles/utils.rs:291-299), through both command lines against the CPU oracle.  Every other input in this suite is
synthetic; a real picture has large flat regions (exact distance ties between tiles), white areas (the border trim of
prepare_tile, utils.rs:93-167) and a palette PNG decode in front of the path.
tests/golden/warhol.png is a byte copy of that file (MIT-licensed example image of the reference repository)."""
import os
import subprocess

import numpy as np
import pytest

import oracle
from emosaic_b200 import api, cli

pytestmark = pytest.mark.gpu
PIL = pytest.importorskip("PIL.Image")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WARHOL = os.path.join(ROOT, "tests", "golden", "warhol.png")


def oracle_tile(img, ts, crop):
    return oracle.resize_lanczos3(img, ts, ts, oracle.prepare_view(img, ts, crop))


@pytest.fixture(scope="module")
def warhol():
    img = np.asarray(PIL.open(WARHOL).convert("RGB"), dtype=np.uint8)
    assert img.shape == (1600, 1600, 3)
    return img


@pytest.fixture(scope="module")
def dirs(tmp_path_factory, warhol):
    d = tmp_path_factory.mktemp("warhol")
    rng = np.random.default_rng(42)
    syn = d / "synthetic"
    syn.mkdir()
    for i in range(300):                           # C1's library size: the single-leaf regime of kiddo (T <= 320)
        base = rng.integers(0, 256, 3)
        img = np.clip(base + rng.integers(-25, 26, (24, 30, 3)), 0, 235).astype(np.uint8)
        PIL.fromarray(img).save(syn / f"s{i:03d}.png")
    crops = d / "crops"
    crops.mkdir()
    k = 0
    for y in range(0, 1600, 200):                  # 64 crops of warhol itself, 200 x 200 (some are mostly white)
        for x in range(0, 1600, 200):
            PIL.fromarray(warhol[y:y + 200, x:x + 200]).save(crops / f"c{k:02d}.png")
            k += 1
    return d


def expected(paths, src, dim, ts, crop):
    decoded = [np.asarray(PIL.open(p).convert("RGB"), dtype=np.uint8) for p in paths]
    px_an, px_rd, keep = [], [], []
    for p, im in zip(paths, decoded):
        try:
            a, r = oracle_tile(im, ts, crop), oracle_tile(im, ts, True)
        except Exception:                          # e.g. an all-white crop: prepare_tile's assertion (utils.rs:157-158)
            continue
        px_an.append(a); px_rd.append(r); keep.append(p)
    px_an, px_rd = np.stack(px_an), np.stack(px_rd)
    colors = oracle.analyse_tiles(px_an, dim * dim)
    item, dist = oracle.match(colors, src)
    return keep, colors, item, dist, px_rd


def test_prepare_tile_of_the_fixture(ctx, warhol):
    """utils.rs:291-299 test_prepare_tile: prepare_tile(example/warhol.png, 32, true) is a 32 x 32 image — here also equal to
    the oracle's pixels, cropped and uncropped."""
    for crop in (True, False):
        t = cli.prepare_tile(WARHOL, 32, crop, ctx)
        assert t.shape == (32, 32, 3)
        assert (t == oracle_tile(warhol, 32, crop)).all()


@pytest.mark.parametrize("mode,dim", [("1", 1), ("4to1", 2)])
def test_python_cli_warhol_source_c1_shape(dirs, warhol, mode, dim, capsys):
    """C1's shape with a real source: --downsample 16 turns the 1600 x 1600 picture into 100 x 100 (Lanczos3 on the GPU,
    main.rs:567-595), 300 tiles of 16 x 16 -> 1600 x 1600 out."""
    out = dirs / f"w_{mode}.png"
    argv = ["-s", "16", "-o", str(out), WARHOL, "mosaic", str(dirs / "synthetic"), "-m", mode, "--downsample", "16",
            "--extensions", "png", "-f"]
    assert cli.main(argv) == 0
    assert "Resizing source image from 1600x1600 to 100x100" in capsys.readouterr().err
    small = oracle.resize_lanczos3(warhol, 100, 100)
    paths = cli.find_images(str(dirs / "synthetic"), {"png"})
    keep, colors, item, dist, px_rd = expected(paths, small, dim, 16, False)
    assert len(keep) == 300
    got = np.asarray(PIL.open(out))
    assert got.shape == (1600 // dim, 1600 // dim, 3)
    assert (got == oracle.render(px_rd, item)).all()
    # the flat regions of the picture produce real ties: many blocks share one colour vector and must share one tile
    q = small.reshape(100 // dim, dim, 100 // dim, dim, 3).transpose(0, 2, 1, 3, 4).reshape(-1, dim * dim * 3)
    _, inv = np.unique(q, axis=0, return_inverse=True)
    assert len(set(zip(inv.tolist(), item.reshape(-1).tolist()))) == inv.max() + 1


def test_both_clis_warhol_tiles_from_the_picture(dirs, warhol):
    """Tiles = crops of warhol itself through prepare_tile (trim view, --crop, Lanczos3 on the GPU); source = the picture
    downsampled by 8.  Python and C++ front ends against the oracle; crops that prepare_tile rejects (no non-white interior)
    are listed and skipped like main.rs:757-806."""
    small = oracle.resize_lanczos3(warhol, 200, 200)
    paths = cli.find_images(str(dirs / "crops"), {"png"})
    keep, colors, item, dist, px_rd = expected(paths, small, 1, 8, True)
    assert 32 <= len(keep) <= 64
    want = oracle.render(px_rd, item)
    out_py = dirs / "crops_py.png"
    assert cli.main(["-s", "8", "-o", str(out_py), "--crop", WARHOL, "mosaic", str(dirs / "crops"), "--downsample", "8",
                     "--extensions", "png", "-f"]) == 0
    assert (np.asarray(PIL.open(out_py)) == want).all()
    out_cpp = dirs / "crops_cpp.png"
    exe = os.path.join(ROOT, "emosaic_b200", "emosaic")
    r = subprocess.run([exe, "-s", "8", "-o", str(out_cpp), "--crop", WARHOL, "mosaic", str(dirs / "crops"), "--downsample", "8",
                        "--extensions", "png", "-f"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert (np.asarray(PIL.open(out_cpp)) == want).all()
    if len(keep) < 64:
        assert f"Failed to read the following images({64 - len(keep)}):" in r.stderr


def test_warhol_tint_overlay(dirs, warhol, ctx):
    """-t 0.5 on the real picture: the overlay is the image as opened (1600 x 1600), nearest-sampled onto the 800 x 800
    output of a --downsample 16 run with 8 x 8 tiles (main.rs:447-478)."""
    out = dirs / "w_tint.png"
    assert cli.main(["-s", "8", "-o", str(out), WARHOL, "mosaic", str(dirs / "synthetic"), "--downsample", "16", "--extensions", "png",
                     "-t", "0.5", "-f"]) == 0
    small = oracle.resize_lanczos3(warhol, 100, 100)
    paths = cli.find_images(str(dirs / "synthetic"), {"png"})
    keep, colors, item, dist, px_rd = expected(paths, small, 1, 8, False)
    got = np.asarray(PIL.open(out))
    assert got.shape == (800, 800, 4)
    assert (got == oracle.tint(oracle.render(px_rd, item), warhol, 127)).all()
