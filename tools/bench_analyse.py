"""HBM throughput of the analysis kernels at the tile sizes of the BASELINE configs (run under gpurun)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import emosaic_b200 as emo
ctx = emo.Context(0)
dev = torch.device("cuda", 0)
for ts, T in ((8, 16_000_000), (16, 4_000_000), (32, 2_000_000), (64, 500_000)):
    tiles = torch.randint(0, 256, (T * ts * ts * 3,), dtype=torch.uint8, device=dev)
    o1 = torch.empty(T * 3, dtype=torch.uint8, device=dev); o4 = torch.empty(T * 12, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    for name, fn, outb in (("dim1", lambda: ctx.analyse_dev(tiles.data_ptr(), T, ts, 1, o1.data_ptr()), 3),
                           ("dim2", lambda: ctx.analyse_dev(tiles.data_ptr(), T, ts, 2, o4.data_ptr()), 12),
                           ("fused", lambda: ctx.analyse_fused_dev(tiles.data_ptr(), T, ts, o1.data_ptr(), o4.data_ptr()), 15)):
        fn(); ctx.sync()
        t = []
        for _ in range(9):  # single launches vary by up to 1.6x on this pool (0.47 .. 0.78 ms at ts=8): median of 9
            ctx.timer_start(); fn(); t.append(ctx.timer_stop())
        ms = float(np.median(t))
        print(f"ts={ts:3d} T={T:9d} {name:5s}: {ms:7.3f} ms  {T*(ts*ts*3+outb)/(ms*1e-3)/1e9:8.1f} GB/s")
    del tiles
