"""Tuning sweep for the match kernel launch heuristics (run under gpurun)."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import emosaic_b200 as emo

ctx = emo.Context(0)
dev = torch.device("cuda", 0)

def bench(N, T, S, H=None, reps=5):
    H = H or S
    dim = int(N ** 0.5)
    colors = torch.from_numpy(np.random.default_rng(1).integers(0, 256, (T * N * 3,), dtype=np.uint8)).to(dev)
    src = torch.from_numpy(np.random.default_rng(2).integers(0, 256, (H * S * 3,), dtype=np.uint8)).to(dev)
    Q = (S // dim) * (H // dim)
    item = torch.empty(Q, dtype=torch.int32, device=dev); dist = torch.empty(Q, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    ctx.set_library_dev(colors.data_ptr(), 0, T, N, 0)
    out = {}
    for R in (2, 4, 8):
        for sp in (0, 1, 2, 4):
            os.environ["EMO_MATCH_R"] = str(R)
            if sp: os.environ["EMO_MATCH_SPLITS"] = str(sp)
            else: os.environ.pop("EMO_MATCH_SPLITS", None)
            ctx.match_dev(src.data_ptr(), S, H, item.data_ptr(), dist.data_ptr()); ctx.sync()
            ts = []
            for _ in range(reps):
                ctx.timer_start(); ctx.match_dev(src.data_ptr(), S, H, item.data_ptr(), dist.data_ptr()); ts.append(ctx.timer_stop())
            out[(R, sp)] = float(np.median(ts))
    L = T if N == 1 else 2 * T
    best = min(out.values())
    print(f"N={N} T={T} src={S}x{H} Q={Q} L={L}: ideal_ms~{Q*L*(3*N/4*1.0+0.3)/18.6e12*1e3:.3f}")
    for k, v in sorted(out.items()):
        print(f"   R={k[0]} splits={k[1] or 'auto':>4}: {v:8.3f} ms {'<-- best' if v == best else ''}")

bench(1, 100000, 4096, 85)     # C4 e2e chunk
bench(1, 4096, 1024)           # C5 match
bench(1, 20000, 2048)          # mid-size library
bench(1, 100000, 4096, 512)    # C4 at N=8 ranks
