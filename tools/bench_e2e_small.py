import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
import emosaic_b200 as emo
ctx = emo.Context(0)
for name, T, ts, S, dim in (("C2", 10000, 16, 1024, 2), ("C5-rgb", 4096, 32, 1024, 1), ("C1", 300, 16, 100, 1)):
    tiles = np.random.default_rng(1234).integers(0, 256, (T, ts, ts, 3), dtype=np.uint8)
    ctx.set_library(ctx.analyse_tiles(tiles, dim), tiles)
    src = np.random.default_rng(5678).integers(0, 256, (S, S, 3), dtype=np.uint8)
    sp = ctx.host_alloc(src.nbytes); sp[:] = src.reshape(-1); s_img = sp.reshape(S, S, 3)
    ob = (S // dim * ts) ** 2 * 3
    op = ctx.host_alloc(ob); o_img = op.reshape(S // dim * ts, S // dim * ts, 3)
    ctx.mosaic(s_img, 3, 0, out=o_img, want_maps=False)
    ts_ = []
    for _ in range(5):
        ctx.timer_start(); ctx.mosaic(s_img, 3, 0, out=o_img, want_maps=False); ts_.append(ctx.timer_stop())
    ms = float(np.median(ts_))
    print(f"{name}: e2e {ms:.3f} ms, {S*S/ms/1e3:.1f} M px/s, out {ob/ms/1e6:.1f} GB/s")
    ctx.host_free(sp); ctx.host_free(op)
