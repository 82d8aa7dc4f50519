"""Argument validation of both command lines (src/main.rs:28-155, :272-345) — everything that is decided before a GPU is
touched — and the loud failure without one (no CPU fallback).  CPU only."""
import os
import subprocess

import numpy as np
import pytest

from emosaic_b200 import cli

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PIL = pytest.importorskip("PIL.Image")


@pytest.fixture(scope="module")
def work(tmp_path_factory):
    d = tmp_path_factory.mktemp("cli_cpu")
    (d / "tiles").mkdir()
    rng = np.random.default_rng(0)
    PIL.fromarray(rng.integers(0, 256, (8, 8, 3), dtype=np.uint8)).save(d / "src.png")
    for i in range(3):
        PIL.fromarray(rng.integers(0, 256, (8, 8, 3), dtype=np.uint8)).save(d / "tiles" / f"t{i}.png")
    return d


def test_python_cli_validation(work):
    d = work
    src, tiles = str(d / "src.png"), str(d / "tiles")
    assert cli.main(["-s", "0", src, "mosaic", tiles]) == 1                      # main.rs:272-283 validate_tile_size
    assert cli.main(["-s", "2048", src, "mosaic", tiles]) == 1
    assert cli.main(["-s", "8", str(d / "missing.png"), "mosaic", tiles]) == 1   # validate_input_image
    (d / "notes.txt").write_text("x")
    assert cli.main(["-s", "8", str(d / "notes.txt"), "mosaic", tiles]) == 1     # unsupported format
    assert cli.main(["-s", "8", "-o", str(d / "nodir" / "o.png"), src, "mosaic", tiles]) == 1   # validate_output_path
    assert cli.main(["-s", "8", src]) == 0                                        # main.rs:378-379 `None => ()`: validated, nothing to do
    assert cli.main(["-s", "8", src, "mosaic", tiles, "-t", "1.5"]) == 2          # "Value must be between 0 and 1"
    assert cli.main(["-s", "8", src, "mosaic", tiles, "--no-repeat", "--greedy"]) == 2
    assert cli.main(["-s", "8", src, "mosaic", tiles, "--randomize", "5"]) == 2
    assert cli.main(["-s", "8", src, "mosaic", tiles, "--html"]) == 2
    assert cli.main(["-s", "8", src, "mosaic", str(d / "nope")]) == 1
    with pytest.raises(SystemExit):                                               # clap rejects unknown modes
        cli.main(["-s", "8", src, "mosaic", tiles, "-m", "7"])


def test_find_images_order_and_extensions(work):
    d = work
    (d / "tiles" / "sub").mkdir(exist_ok=True)
    PIL.fromarray(np.zeros((4, 4, 3), np.uint8)).save(d / "tiles" / "sub" / "a.png")
    (d / "tiles" / "readme.md").write_text("x")
    found = cli.find_images(str(d / "tiles"), {"png"})
    assert len(found) == 4 and all(p.endswith(".png") for p in found)
    assert cli.find_images(str(d / "tiles"), {"jpg"}) == []


def test_no_gpu_means_loud_failure(work):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import emosaic_b200
    d = work
    with pytest.raises(emosaic_b200.EmosaicError, match="no CPU path"):
        cli.main(["-s", "8", "-o", str(d / "o.png"), str(d / "src.png"), "mosaic", str(d / "tiles"), "--extensions", "png"])
    with pytest.raises(emosaic_b200.EmosaicError, match="no CPU path"):            # host-layer entry points too
        emosaic_b200.resize_source(np.zeros((5, 5, 3), np.uint8), 1, 2)


def test_cpp_cli_validation_and_no_gpu(work):
    exe = os.path.join(ROOT, "emosaic_b200", "emosaic")
    if not os.path.exists(exe):
        pytest.fail(f"{exe} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    d = work
    src, tiles = str(d / "src.png"), str(d / "tiles")

    def run(*a):
        return subprocess.run([exe, *a], capture_output=True, text=True, timeout=60)

    assert run().returncode == 2 and "usage" in run().stderr
    assert run("-s", "8", src, "mosaic", tiles, "--greedy").returncode == 2
    assert run("-s", "8", src, "mosaic", tiles, "-t", "2").returncode == 2
    assert run("-s", "8", src, "mosaic", tiles, "--downsample", "0").returncode == 2
    import torch
    if not torch.cuda.is_available():
        r = run("-s", "8", "-o", str(d / "o.png"), src, "mosaic", tiles, "--extensions", "png")
        assert r.returncode == 1 and "no CPU path" in r.stderr
        r = run("-s", "8", "-o", str(d / "t.png"), src, "prepare")                  # main.rs:380-386
        assert r.returncode == 1 and "Failed to prepare tile" in r.stderr and "no CPU path" in r.stderr
