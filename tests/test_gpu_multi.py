"""Multi-GPU product path (include/emosaic_cuda.h "multi-GPU") against the CPU oracle and the single-GPU result.

What is sharded are the reference's own rayon tasks: block rows in render() (src/mosaic/rendering.rs:68-101) and tiles in the
analysis build (src/main.rs:760-794).  (b) emo_group_*: one process, one worker thread per GPU, every stripe copied straight
into the caller's one image; (a) emo_comm_*: one process per GPU, library replicated by the root's NCCL broadcast.
Skipped on a box with fewer than 2 GPUs."""
import os
import subprocess
import sys

import numpy as np
import pytest

import emosaic_b200 as emo
import oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def n_gpus() -> int:
    import torch
    return torch.cuda.device_count()


def group_sizes():
    n = n_gpus()
    return sorted({s for s in (2, 3, n) if 2 <= s <= n})


@pytest.fixture(scope="module")
def need2():
    if n_gpus() < 2:
        pytest.skip("needs at least 2 GPUs")


def c2_inputs():
    tiles = np.random.default_rng(1234).integers(0, 256, (10000, 16, 16, 3), dtype=np.uint8)
    src = np.random.default_rng(5678).integers(0, 256, (1024, 1024, 3), dtype=np.uint8)
    return tiles, src


def test_group_c2_full_vs_oracle_and_single_gpu(need2, ctx):
    """C2 at full size (4to1, 10k tiles, 1024x1024 source): the stitched item / dist / image of 2..n GPUs equal the KD-tree
    oracle and the one-GPU result bit for bit."""
    tiles, src = c2_inputs()
    colors = ctx.analyse_tiles(tiles, 2)
    ctx.set_library(colors, tiles)
    out1, item1, dist1 = ctx.mosaic(src, 3, 0)
    ri, rd = oracle.KdTree(colors).match(src)
    assert (item1 == ri).all() and (dist1 == rd).all()
    for n in group_sizes():
        g = emo.Group(n)
        try:
            gc = g.analyse_tiles(tiles, 2)           # tiles sharded over the GPUs, results written in place
            assert (gc == colors).all()
            g.set_library(gc, tiles)                 # H2D on GPU 0 + one NCCL broadcast
            out, item, dist = g.mosaic(src, 3, 0)
            assert (item == ri).all() and (dist == rd).all(), f"{n} GPUs: maps differ from the oracle"
            assert (out == out1).all(), f"{n} GPUs: image differs from one GPU"
            assert all(m.launch_count() > 0 for m in g.members), "a GPU of the group did no work"
        finally:
            g.close()
    assert (out1 == oracle.render(tiles, ri)).all()


def test_group_c4_stripe_pinned_and_tint(need2, ctx):
    """C4 geometry (100k tiles, ts 8, 4096 wide) on a 200-row stripe — rows do not divide by 3 — with a pinned output image, and the
    RGBA tint path; item / dist / image against the oracle and one GPU."""
    T = 100_000
    tiles = np.random.default_rng(1234).integers(0, 256, (T, 8, 8, 3), dtype=np.uint8)
    src = np.random.default_rng(5678).integers(0, 256, (200, 4096, 3), dtype=np.uint8)
    colors = ctx.analyse_tiles(tiles, 1)
    ri, rd = oracle.KdTree(colors).match(src)
    want = oracle.render(tiles, ri)
    want4 = oracle.tint(want[:64 * 8], src[:64], 127)
    for n in group_sizes():
        g = emo.Group(n)
        try:
            g.set_library(colors, tiles)
            out = np.zeros((200 * 8, 4096 * 8, 3), np.uint8)
            emo.host_register(out)
            try:
                _, item, dist = g.mosaic(src, 3, 0, out=out)
            finally:
                emo.host_unregister(out)
            assert (item == ri).all() and (dist == rd).all()
            assert (out == want).all()
            out4, item4, _ = g.mosaic(src[:64], 4, 127)
            assert (item4 == ri[:64]).all() and (out4 == want4).all()
            # the index and the scan give the same stitched maps
            g.set_match_mode("scan")
            _, item_s, dist_s = g.mosaic(src[:8], 3, 0)
            g.set_match_mode("auto")
            assert (item_s == ri[:8]).all() and (dist_s == rd[:8]).all()
        finally:
            g.close()


def test_group_more_gpus_than_rows_and_fused_analysis(need2):
    n = n_gpus()
    rng = np.random.default_rng(7)
    tiles = rng.integers(0, 256, (1001, 64, 64, 3), dtype=np.uint8)     # ragged tile shards
    g = emo.Group(n)
    try:
        o1, o4 = g.analyse_tiles_fused(tiles)
        assert (o1 == oracle.analyse_tiles(tiles, 1)).all() and (o4 == oracle.analyse_tiles(tiles, 4)).all()
        t16 = rng.integers(0, 256, (300, 16, 16, 3), dtype=np.uint8)
        c = g.analyse_tiles(t16, 2)
        g.set_library(c, t16)
        src = rng.integers(0, 256, (2, 40, 3), dtype=np.uint8)          # one block row: every GPU but the first idles
        out, item, dist = g.mosaic(src, 3, 0)
        ri, rd = oracle.match(c, src)
        assert (item == ri).all() and (dist == rd).all() and (out == oracle.render(t16, ri)).all()
    finally:
        g.close()


def test_group_errors(need2):
    with pytest.raises(emo.EmosaicError) as e:
        emo.Group([0, 0])
    assert "twice" in str(e.value)
    g = emo.Group(2)
    try:
        with pytest.raises(emo.EmosaicError) as e:
            g.mosaic(np.zeros((4, 4, 3), np.uint8))
        assert e.value.code == -4                                      # no library yet
        t = np.zeros((10, 8, 8, 3), np.uint8)
        g.set_library(np.zeros((10, 4, 3), np.uint8), t)
        with pytest.raises(emo.EmosaicError) as e:
            g.mosaic(np.zeros((5, 4, 3), np.uint8))
        assert "divisible by 2" in str(e.value)                        # main.rs:603-611
    finally:
        g.close()


WORKER = r"""
import os, sys, time
import numpy as np
sys.path.insert(0, sys.argv[1])
import emosaic_b200 as emo
rank, world, tmp = int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
ctx = emo.Context(rank)
idf = os.path.join(tmp, "nccl_id")
if rank == 0:
    open(idf + ".tmp", "wb").write(emo.Context.comm_unique_id())
    os.rename(idf + ".tmp", idf)
while not os.path.exists(idf):
    time.sleep(0.01)
ctx.comm_init_rank(open(idf, "rb").read(), rank, world)
assert ctx.comm_info()["world"] == world and ctx.comm_info()["nccl_version"] > 20000
src = np.load(os.path.join(tmp, "src.npy"))
if rank == 0:                       # only the root holds the library; the others receive it over NCCL
    tiles = np.load(os.path.join(tmp, "tiles.npy"))
    colors = ctx.analyse_tiles(tiles, 2)
    ctx.comm_set_library(colors, tiles, root=0)
else:
    ctx.comm_set_library(None, None, root=0)
    assert (ctx.T, ctx.N, ctx.dim, ctx.ts) == (3000, 4, 2, 16)
a, b = emo.stripe_bounds(src.shape[0] // 2, world, rank)
out, item, dist = ctx.mosaic(src[a * 2:b * 2], 3, 0)
np.savez(os.path.join(tmp, f"part{rank}.npz"), out=out, item=item, dist=dist)
# sharded analysis + all-gather on the device (ragged: 1001 tiles)
import ctypes as C
t64 = np.load(os.path.join(tmp, "t64.npy"))
ta, tb = emo.stripe_bounds(len(t64), world, rank)
mine = np.ascontiguousarray(t64[ta:tb])
d_in = ctx.dev_alloc(max(mine.nbytes, 1)); d_loc = ctx.dev_alloc((tb - ta) * 12 + 1); d_all = ctx.dev_alloc(len(t64) * 12)
ctx.h2d(d_in, mine)
ctx.analyse_dev(d_in, tb - ta, 64, 2, d_loc)
ctx.comm_allgather_analysis_dev(d_loc, len(t64), 12, d_all)
full = np.zeros((len(t64), 4, 3), np.uint8)
ctx.d2h(full, d_all); ctx.sync()
np.save(os.path.join(tmp, f"gather{rank}.npy"), full)
ctx.close()
"""


@pytest.mark.parametrize("world", [2, 8])
def test_one_process_per_gpu_comm(need2, world, tmp_path):
    """emo_comm_*: `world` processes, one GPU each; rank 0 owns the library and replicates it with emo_comm_set_library
    (NCCL broadcast); every rank renders its stripe; the concatenation equals the oracle.  Plus the sharded analysis build
    with the device-side all-gather (ragged shards)."""
    if n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    tiles = np.random.default_rng(11).integers(0, 256, (3000, 16, 16, 3), dtype=np.uint8)
    src = np.random.default_rng(12).integers(0, 256, (2 * 37, 2 * 50, 3), dtype=np.uint8)   # 37 block rows: ragged stripes
    t64 = np.random.default_rng(13).integers(0, 256, (1001, 64, 64, 3), dtype=np.uint8)
    np.save(tmp_path / "tiles.npy", tiles); np.save(tmp_path / "src.npy", src); np.save(tmp_path / "t64.npy", t64)
    (tmp_path / "worker.py").write_text(WORKER)
    env = dict(os.environ)
    env.pop("OMP_NUM_THREADS", None)
    procs = [subprocess.Popen([sys.executable, str(tmp_path / "worker.py"), ROOT, str(r), str(world), str(tmp_path)], env=env,
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=300)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(o)
    for r, p in enumerate(procs):
        assert p.returncode == 0, f"rank {r} failed:\n{outs[r][-3000:]}"
    parts = [np.load(tmp_path / f"part{r}.npz") for r in range(world)]
    item = np.concatenate([p["item"] for p in parts]); dist = np.concatenate([p["dist"] for p in parts])
    out = np.concatenate([p["out"] for p in parts])
    colors = oracle.analyse_tiles(tiles, 4)
    ri, rd = oracle.match(colors, src)
    assert (item == ri).all() and (dist == rd).all() and (out == oracle.render(tiles, ri)).all()
    want = oracle.analyse_tiles(t64, 4)
    for r in range(world):
        assert (np.load(tmp_path / f"gather{r}.npy") == want).all(), f"rank {r}: gathered analysis differs"


def test_both_command_lines_with_gpus_flag(need2, tmp_path):
    """`--gpus 2` in the Python and the C++ front end: same PNG as one GPU (which the CLI tests pin to the oracle)."""
    PIL = pytest.importorskip("PIL.Image")
    from emosaic_b200 import cli
    rng = np.random.default_rng(21)
    tiles_dir = tmp_path / "tiles"
    tiles_dir.mkdir()
    for i in range(70):
        img = np.clip(rng.integers(0, 256, 3) + rng.integers(-40, 41, (8, 8, 3)), 0, 255).astype(np.uint8)
        img[:, :3] //= 2
        PIL.fromarray(img).save(tiles_dir / f"t{i:02d}.png")
    src = rng.integers(0, 256, (2 * 23, 2 * 31, 3), dtype=np.uint8)       # 23 block rows in 4to1: ragged stripes
    PIL.fromarray(src).save(tmp_path / "src.png")
    exe = os.path.join(ROOT, "emosaic_b200", "emosaic")
    for mode in ("1", "4to1"):
        outs = {}
        for gpus in (1, 2):
            o_py, o_cpp = tmp_path / f"py_{mode}_{gpus}.png", tmp_path / f"cpp_{mode}_{gpus}.png"
            base = ["-s", "8", str(tmp_path / "src.png"), "mosaic", str(tiles_dir), "-m", mode, "--extensions", "png", "-f"]
            assert cli.main(["--gpus", str(gpus), "-o", str(o_py)] + base) == 0
            r = subprocess.run([exe, "--gpus", str(gpus), "-o", str(o_cpp)] + base, capture_output=True, text=True, timeout=300)
            assert r.returncode == 0, r.stderr
            outs[gpus] = (np.asarray(PIL.open(o_py)), np.asarray(PIL.open(o_cpp)))
        assert (outs[1][0] == outs[2][0]).all() and (outs[1][1] == outs[2][1]).all()
        # and against the oracle directly (the C++ front end takes tile_size x tile_size files as prepared tiles)
        px = np.stack([np.asarray(PIL.open(p)) for p in sorted(str(p) for p in tiles_dir.glob("*.png"))])
        N = 1 if mode == "1" else 4
        item, _ = oracle.match(oracle.analyse_tiles(px, N), src)
        assert (outs[2][1] == oracle.render(px, item)).all()
