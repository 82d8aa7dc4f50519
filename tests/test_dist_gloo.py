"""world_size-2 gloo test of the row-stripe sharding used by bench.py --gpus N (CPU only).

Each rank takes its stripe of block rows (emosaic_b200.sharding), the library is broadcast from
rank 0 exactly like the NCCL broadcast in the GPU run, every rank matches + composes its stripe
(with the CPU oracle standing in for the kernels) and rank 0 checks that the concatenated slabs
equal the single-process result.  No collective is needed inside the loop."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, N, q):
    sys.path.insert(0, ROOT)
    import oracle
    from emosaic_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dim = int(N ** 0.5)
    T, ts, H, W = 120, 4 * dim, 14 * dim, 10 * dim
    if rank == 0:
        rng = np.random.default_rng(5)
        tiles = rng.integers(0, 256, (T, ts, ts, 3), dtype=np.uint8)
        colors = oracle.analyse_tiles(tiles, N)
        src = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    else:
        tiles = np.zeros((T, ts, ts, 3), np.uint8)
        colors = np.zeros((T, N, 3), np.uint8)
        src = np.zeros((H, W, 3), np.uint8)
    for a in (tiles, colors, src):  # library + source replicated once, up front
        t = torch.from_numpy(a)
        dist.broadcast(t, 0)
    stripe, (a, b) = sharding.source_stripe(src, dim, world, rank)
    item, dd = oracle.match(colors, stripe)
    slab = oracle.render(tiles, item)
    # rank 0 concatenates the slabs (disjoint row ranges)
    gathered = [None] * world
    dist.gather_object((a, b, item, dd, slab), gathered if rank == 0 else None, dst=0)
    # max-over-ranks timing reduction used by bench.py
    tmax = torch.tensor([float(rank + 1)])
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    if rank == 0:
        gathered.sort(key=lambda g: g[0])
        full_item = np.concatenate([g[2] for g in gathered], 0)
        full_dist = np.concatenate([g[3] for g in gathered], 0)
        full_out = np.concatenate([g[4] for g in gathered], 0)
        ri, rd = oracle.match(colors, src)
        ok = (full_item == ri).all() and (full_dist == rd).all() and (full_out == oracle.render(tiles, ri)).all()
        q.put(bool(ok) and float(tmax) == world)
    dist.destroy_process_group()


@pytest.mark.parametrize("N", [1, 4])
def test_two_rank_stripes(N):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, N, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def _analysis_worker(rank, world, port, T, q):
    sys.path.insert(0, ROOT)
    import oracle
    from emosaic_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ts = 8
    tiles = np.random.default_rng(77).integers(0, 256, (T, ts, ts, 3), dtype=np.uint8)   # same seed on every rank
    a, b = sharding.stripe_bounds(T, world, rank)
    ok = True
    for N in (1, 4):
        local = oracle.analyse_tiles(tiles[a:b], N)            # the oracle stands in for emo_analyse on this rank's range
        full = sharding.gather_analysis(torch.from_numpy(local.reshape(-1)), T, 3 * N, world, rank)
        ok = ok and (full.numpy().reshape(T, N, 3) == oracle.analyse_tiles(tiles, N)).all()
    q.put(bool(ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("T", [64, 101])
def test_two_rank_analysis_build(T):
    """C3 across ranks: each rank analyses its contiguous range of tiles, one all_gather assembles [T, 3N] (even and ragged split)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_analysis_worker, args=(r, 2, port, T, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True and q.get(timeout=5) is True


def _prepare_worker(rank, world, port, T, q):
    sys.path.insert(0, ROOT)
    import oracle
    from emosaic_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ts = 8
    photos = np.random.default_rng(78).integers(0, 230, (T, 40, 44, 3), dtype=np.uint8)   # same seed on every rank
    a, b = sharding.stripe_bounds(T, world, rank)

    def prep(p):  # the oracle stands in for emo_resize on this rank's photos (tiles/utils.rs:93-189)
        return oracle.resize_lanczos3(p, ts, ts, oracle.prepare_view(p, ts, True))

    local = np.stack([prep(p) for p in photos[a:b]]) if b > a else np.zeros((0, ts, ts, 3), np.uint8)
    full = sharding.gather_analysis(torch.from_numpy(local.reshape(-1)), T, ts * ts * 3, world, rank)
    want = np.stack([prep(p) for p in photos])
    q.put(bool((full.numpy().reshape(T, ts, ts, 3) == want).all()))
    dist.destroy_process_group()


@pytest.mark.parametrize("T", [10, 7])
def test_two_rank_tile_preparation(T):
    """Tile preparation across ranks: each rank resizes its contiguous range of photos, the prepared tiles [T, ts, ts, 3] are
    assembled by the same single all_gather as the analysis results (even and ragged split)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_prepare_worker, args=(r, 2, port, T, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True and q.get(timeout=5) is True
