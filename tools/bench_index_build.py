"""Build time of the 1to1 colour-cube index (index.cu) for a few library sizes (run under gpurun); EMO_INDEX_SWEEP=0 selects the
in-place strided sweeps for an A-B comparison."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import emosaic_b200 as emo

ctx = emo.Context(0)
dev = torch.device("cuda", 0)
for T, kind in ((100_000, "clustered"), (100_000, "uniform"), (10_000, "uniform"), (4_000_000, "uniform")):
    rng = np.random.default_rng(1)
    c = rng.integers(0, 256, (T, 3), dtype=np.uint8) if kind == "uniform" else np.clip(rng.normal(127, 9, (T, 3)), 0, 255).astype(np.uint8)
    colors = torch.from_numpy(c.reshape(-1)).to(dev)
    ctx.set_library_dev(colors.data_ptr(), 0, T, 1, 0)
    ctx.set_match_mode("index")
    ctx.build_index(); ctx.sync()
    tb = []
    for _ in range(9):
        ctx.timer_start(); ctx.build_index(); tb.append(ctx.timer_stop())
    print(f"T={T:8d} {kind:9s}: build {np.median(tb)*1e3:7.1f} us (min {min(tb)*1e3:.1f})")
