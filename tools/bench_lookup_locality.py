"""How the 1to1 index lookup depends on the table footprint the source touches (run under gpurun): uniform-random pixels
restricted to b < B for B = 256, 128, 64, 32, through the 64 MiB table (one gather) and the compact one (two gathers)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import emosaic_b200 as emo

T, W, H = 100_000, 4096, 4096
ctx = emo.Context(0)
dev = torch.device("cuda", 0)
tiles = torch.from_numpy(np.random.default_rng(1234).integers(0, 256, (T * 64 * 3,), dtype=np.uint8)).to(dev)
colors = torch.empty(T * 3, dtype=torch.uint8, device=dev)
ctx.analyse_dev(tiles.data_ptr(), T, 8, 1, colors.data_ptr())
ctx.set_library_dev(colors.data_ptr(), 0, T, 1, 0)
item = torch.empty(H * W, dtype=torch.int32, device=dev); dist = torch.empty(H * W, dtype=torch.int32, device=dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for B in (256, 128, 64, 32, 8):
    src = np.random.default_rng(5678).integers(0, 256, (H, W, 3), dtype=np.uint8)
    src[..., 2] = (src[..., 2].astype(np.uint16) % B).astype(np.uint8)
    src_d = torch.from_numpy(src.reshape(-1)).to(dev)
    for mode in ("index_wide", "index_compact"):
        ctx.set_match_mode(mode)
        ctx.match_dev(src_d.data_ptr(), W, H, item.data_ptr(), dist.data_ptr()); ctx.sync()
        warm, cold = [], []
        for _ in range(5):
            ctx.timer_start(); ctx.match_dev(src_d.data_ptr(), W, H, item.data_ptr(), dist.data_ptr()); warm.append(ctx.timer_stop())
        for _ in range(5):
            flush.fill_(1); torch.cuda.synchronize()
            ctx.timer_start(); ctx.match_dev(src_d.data_ptr(), W, H, item.data_ptr(), dist.data_ptr()); cold.append(ctx.timer_stop())
        fp = B / 256 * (64 if mode == "index_wide" else 32)
        print(f"b < {B:3d} ({fp:5.1f} MiB of table touched) {mode:14s}: back to back {np.median(warm)*1e3:6.1f} us, after an L2 flush {np.median(cold)*1e3:6.1f} us")
