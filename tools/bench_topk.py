"""emo_topk (ranked candidate pages for the no-repeat renderer) and the whole no-repeat render (run under gpurun)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import emosaic_b200 as emo

ctx = emo.Context(0)
dev = torch.device("cuda", 0)
for N, T, S, k in ((1, 20000, 128, 64), (4, 10000, 256, 64), (4, 10000, 256, 1024), (9, 5000, 240, 64), (1, 100000, 256, 64)):
    dim = int(N ** 0.5)
    rng = np.random.default_rng(1)
    colors = torch.from_numpy(rng.integers(0, 256, (T * N * 3,), dtype=np.uint8)).to(dev)
    src = torch.from_numpy(rng.integers(0, 256, (S * S * 3,), dtype=np.uint8)).to(dev)
    Q = (S // dim) ** 2
    item = torch.empty(Q * k, dtype=torch.int32, device=dev); dist = torch.empty(Q * k, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    ctx.set_library_dev(colors.data_ptr(), 0, T, N, 0)
    ctx.topk_dev(src.data_ptr(), S, S, 0, k, item.data_ptr(), dist.data_ptr()); ctx.sync()
    t = []
    for _ in range(5):
        ctx.timer_start(); ctx.topk_dev(src.data_ptr(), S, S, 0, k, item.data_ptr(), dist.data_ptr()); t.append(ctx.timer_stop())
    L = T if N == 1 else 2 * T
    ms = float(np.median(t))
    print(f"N={N} T={T} blocks={Q} k={k}: {ms:8.3f} ms  {Q * L * 9 / ms / 1e6:8.1f} G distance evaluations/s (9 passes over L per block)")

# whole render: 120 x 120 blocks, 20 000 tiles, 1to1, tile size 8
rng = np.random.default_rng(2)
T, ts = 20000, 8
tiles = rng.integers(0, 256, (T, ts, ts, 3), dtype=np.uint8)
colors = ctx.analyse_tiles(tiles, 1)
src = rng.integers(0, 256, (120, 120, 3), dtype=np.uint8)
tset = emo.TileSet.from_arrays(colors, tiles)
t0 = time.perf_counter(); res = emo.render_nto1_no_repeat(src, tset, ts, ctx); dt = time.perf_counter() - t0
u = np.abs(res.item).reshape(-1)
print(f"render_nto1_no_repeat 14 400 blocks / 20 000 tiles: {dt*1e3:.0f} ms wall (python heap merge included), all tiles distinct: {len(set(u.tolist())) == u.size}")
