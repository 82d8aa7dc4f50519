"""The C++ host mirror (emosaic_b200/csrc/host): the reference's unit tests restated in C++ (host_tests) and the
C++ `emosaic` command line end to end (PNG tiles) against the CPU oracle."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "emosaic_b200")


def _need(name):
    p = os.path.join(BIN, name)
    if not os.path.exists(p):
        pytest.fail(f"{p} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    return p


def test_host_tests_cpu_subset():
    r = subprocess.run([_need("host_tests"), "--cpu"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "FAIL" not in r.stdout and r.stdout.count("PASS") >= 8


@pytest.mark.gpu
def test_host_tests_full():
    r = subprocess.run([_need("host_tests")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "test_analyse_tiles_consistency_9" in r.stdout and "FAIL" not in r.stdout


@pytest.mark.gpu
def test_cpp_cli_end_to_end(tmp_path):
    import oracle
    from emosaic_b200 import cache
    PIL = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(4)
    ts, T = 8, 50
    tiles_dir = tmp_path / "tiles"
    (tiles_dir / "b").mkdir(parents=True)
    tiles, paths = [], []
    for i in range(T):
        img = np.clip(rng.integers(0, 256, 3) + rng.integers(-40, 41, (ts, ts, 3)), 0, 255).astype(np.uint8)
        img[:, :3] //= 2
        p = tiles_dir / ("b" if i % 4 == 0 else "") / f"t{i:02d}.png"
        PIL.fromarray(img).save(p)
        tiles.append(img)
        paths.append(str(p))
    order = np.argsort(_walk_order(str(tiles_dir), paths))
    src = rng.integers(0, 256, (20, 28, 3), dtype=np.uint8)
    PIL.fromarray(src).save(tmp_path / "src.png")
    exe = _need("emosaic")
    for mode, dim in (("1", 1), ("4to1", 2)):
        out = tmp_path / f"out{dim}.png"
        args = [exe, "-s", str(ts), "-o", str(out), str(tmp_path / "src.png"), "mosaic", str(tiles_dir), "-m", mode, "--extensions", "png"]
        r = subprocess.run(args + ["-f"], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stderr
        found = _find_images(str(tiles_dir))
        px = np.stack([tiles[paths.index(p)] for p in found])
        colors = oracle.analyse_tiles(px, dim * dim)
        item, dist = oracle.match(colors, src)
        assert (np.asarray(PIL.open(out)) == oracle.render(px, item)).all()
        blob = (tiles_dir / cache.cache_file_name(dim * dim, False)).read_bytes()
        assert blob == oracle.cache_serialize(colors, np.arange(1, T + 1), [None] * T, found)
        assert "Average color distance" in r.stderr
        # cache reuse + tint
        r = subprocess.run(args + ["-t", "0.5"], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0 and "Reusing analysis cache" in r.stderr, r.stderr
        got = np.asarray(PIL.open(out))
        assert got.shape[2] == 4 and (got == oracle.tint(oracle.render(px, item), src, 127)).all()
    r = subprocess.run([exe, "-s", "8", str(tmp_path / "src.png"), "mosaic", str(tiles_dir), "--greedy"], capture_output=True, text=True)
    assert r.returncode == 2
    # --no-repeat (main.rs:663-664): 560 blocks for 50 tiles is refused, a 6 x 8 block source gets every tile at most once
    base = [exe, "-s", str(ts), "-o", str(tmp_path / "nr.png"), "mosaic", str(tiles_dir), "-m", "1", "--extensions", "png", "--no-repeat"]
    r = subprocess.run(base[:5] + [str(tmp_path / "src.png")] + base[5:], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "Insufficient tiles for no-repeat mode: need 560 tiles but only have 100" in r.stderr
    from oracle import oracle_np as onp
    small = src[:6, :8]
    PIL.fromarray(small).save(tmp_path / "small.png")
    r = subprocess.run(base[:5] + [str(tmp_path / "small.png")] + base[5:], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    found = _find_images(str(tiles_dir))
    px = np.stack([tiles[paths.index(p)] for p in found])
    colors = oracle.analyse_tiles(px, 1)
    item, dist = onp.no_repeat_assign(colors, small)
    assert (np.asarray(PIL.open(tmp_path / "nr.png")) == oracle.render(px, item)).all()
    assert len(set(np.abs(item).reshape(-1).tolist())) == 48


def _find_images(root):
    out = []
    for n in sorted(os.listdir(root)):
        p = os.path.join(root, n)
        if os.path.isdir(p):
            out += _find_images(p)
        elif n.endswith(".png"):
            out.append(p)
    return out


def _walk_order(root, paths):
    found = _find_images(root)
    return [found.index(p) for p in paths]


@pytest.mark.gpu
def test_cpp_cli_prepares_tiles_and_resizes_source(tmp_path):
    """Tiles that are not tile_size x tile_size go through prepare_tile (trim view, crop, Lanczos3 on the GPU; tiles/utils.rs:63-196),
    a 37 x 45 source in 4to1 mode is matched as its Lanczos3 resize to 36 x 44 (main.rs:567-595), --downsample 2, and the tint
    overlay stays the image as opened (main.rs:447-466)."""
    import oracle
    PIL = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(31)
    ts = 8
    tiles_dir = tmp_path / "tiles"
    tiles_dir.mkdir()
    for i in range(30):
        h, w = 30 + i % 7, 36 + i % 5
        img = np.clip(rng.integers(0, 226, 3) + rng.integers(-25, 26, (h, w, 3)), 0, 225).astype(np.uint8)
        img[:, : w // 3] //= 2
        if i % 3 == 0:
            img[:2] = img[-3:] = img[:, :4] = img[:, -1:] = 255            # white frame (utils.rs:93-167)
        PIL.fromarray(img).save(tiles_dir / f"t{i:02d}.png")
    src = rng.integers(0, 256, (37, 45, 3), dtype=np.uint8)
    PIL.fromarray(src).save(tmp_path / "src.png")
    found = _find_images(str(tiles_dir))

    def tile(p, crop):
        im = np.asarray(PIL.open(p).convert("RGB"), dtype=np.uint8)
        return oracle.resize_lanczos3(im, ts, ts, oracle.prepare_view(im, ts, crop))

    px_an, px_rd = np.stack([tile(p, False) for p in found]), np.stack([tile(p, True) for p in found])
    colors = oracle.analyse_tiles(px_an, 4)
    exe = _need("emosaic")
    for downsample in (1, 2):
        out = tmp_path / f"o{downsample}.png"
        args = [exe, "-s", str(ts), "-o", str(out), str(tmp_path / "src.png"), "mosaic", str(tiles_dir), "-m", "2", "--extensions", "png",
                "--downsample", str(downsample)]
        r = subprocess.run(args + ["-f"], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stderr
        nw, nh = oracle.adjust_dims(45, 37, downsample, 2)
        assert f"Resizing source image from 45x37 to {nw}x{nh}" in r.stderr
        item, _ = oracle.match(colors, oracle.resize_lanczos3(src, nw, nh))
        assert (np.asarray(PIL.open(out)) == oracle.render(px_rd, item)).all()
        r = subprocess.run(args + ["-t", "0.5"], capture_output=True, text=True, timeout=120)   # cache reuse + tint
        assert r.returncode == 0 and "Reusing analysis cache" in r.stderr, r.stderr
        assert (np.asarray(PIL.open(out)) == oracle.tint(oracle.render(px_rd, item), src, 127)).all()


@pytest.mark.gpu
def test_cpp_cli_prepare_subcommand(tmp_path):
    """`emosaic -s N [--crop] IMG prepare` (main.rs:380-386) in the C++ front end."""
    import oracle
    PIL = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(8)
    img = np.full((50, 64, 3), 255, np.uint8)
    img[4:45, 6:60] = rng.integers(0, 220, (41, 54, 3))
    PIL.fromarray(img).save(tmp_path / "photo.png")
    exe = _need("emosaic")
    for crop in (False, True):
        out = tmp_path / f"tile{int(crop)}.png"
        r = subprocess.run([exe, "-s", "16", "-o", str(out)] + (["--crop"] if crop else []) + [str(tmp_path / "photo.png"), "prepare"],
                           capture_output=True, text=True, timeout=60)
        assert r.returncode == 0, r.stderr
        want = oracle.resize_lanczos3(img, 16, 16, oracle.prepare_view(img, 16, crop))
        assert (np.asarray(PIL.open(out)) == want).all()
