"""Build and lookup times of the 1to1 colour-cube index (index.cu) on the C4 geometry (run under gpurun)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import emosaic_b200 as emo

ctx = emo.Context(0)
dev = torch.device("cuda", 0)
for T, S, kind in ((100_000, 4096, "uniform"), (100_000, 4096, "clustered"), (1000, 4096, "uniform"), (4_000_000, 4096, "uniform"),
                   (100_000, 1024, "uniform"), (100_000, 16384, "uniform")):
    rng = np.random.default_rng(1)
    c = rng.integers(0, 256, (T, 3), dtype=np.uint8) if kind == "uniform" else np.clip(rng.normal(127, 9, (T, 3)), 0, 255).astype(np.uint8)
    colors = torch.from_numpy(c.reshape(-1)).to(dev)
    src = torch.randint(0, 256, (S * S * 3,), dtype=torch.uint8, device=dev)
    Q = S * S
    item = torch.empty(Q, dtype=torch.int32, device=dev); dist = torch.empty(Q, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    ctx.set_library_dev(colors.data_ptr(), 0, T, 1, 0)
    ctx.set_match_mode("index")
    tb, tl = [], []
    for _ in range(7):
        ctx.timer_start(); ctx.build_index(); tb.append(ctx.timer_stop())
    for _ in range(7):
        ctx.timer_start(); ctx.match_dev(src.data_ptr(), S, S, item.data_ptr(), dist.data_ptr()); tl.append(ctx.timer_stop())
    ms = float(np.median(tl))
    byts = Q * 11
    print(f"T={T:8d} {kind:9s} Q={Q:10d}: build {np.median(tb):7.3f} ms (min {min(tb):.3f})  lookup {ms:7.3f} ms  "
          f"{Q/ms/1e6:8.2f} G px/s  {byts/ms/1e6:7.1f} GB/s algorithmic (3 B in + 8 B out per px)")
    if T == 100_000 and S == 4096 and kind == "uniform":
        # photo-like source: a smooth field (bilinear upsample of a 33x33 random grid) + N(0, sigma) noise per channel
        for sigma in (0.0, 2.0, 8.0):
            g = torch.rand(1, 3, 33, 33, device=dev) * 255
            img = torch.nn.functional.interpolate(g, size=(S, S), mode="bilinear", align_corners=True)[0].permute(1, 2, 0)
            img = (img + torch.randn_like(img) * sigma).clamp(0, 255).to(torch.uint8).contiguous().reshape(-1)
            torch.cuda.synchronize()
            tl = []
            for _ in range(7):
                ctx.timer_start(); ctx.match_dev(img.data_ptr(), S, S, item.data_ptr(), dist.data_ptr()); tl.append(ctx.timer_stop())
            print(f"    photo-like source, noise sigma {sigma}: lookup {np.median(tl):7.3f} ms  {Q/np.median(tl)/1e6:8.2f} G px/s")
            del img
    del src, item, dist
