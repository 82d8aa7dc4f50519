"""Throughput of match_wide_kernel (--mode 5..128) vs the VABSDIFF4 issue rate (run under gpurun)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import emosaic_b200 as emo

ctx = emo.Context(0)
dev = torch.device("cuda", 0)
from tools.probe import probe_int_pipe
peak = probe_int_pipe(0, 1)  # VABSDIFF4 thread-instr/s
for N, T, S in ((4, 20000, 2048), (9, 20000, 2040), (16, 20000, 2048), (25, 20000, 2000), (64, 10000, 2048), (256, 5000, 2048), (1024, 4000, 2048), (16384, 2000, 2048)):
    dim = int(N ** 0.5)
    colors = torch.from_numpy(np.random.default_rng(1).integers(0, 256, (T * N * 3,), dtype=np.uint8)).to(dev)
    src = torch.from_numpy(np.random.default_rng(2).integers(0, 256, (S * S * 3,), dtype=np.uint8)).to(dev)
    Q = (S // dim) ** 2
    item = torch.empty(Q, dtype=torch.int32, device=dev); dist = torch.empty(Q, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    ctx.set_library_dev(colors.data_ptr(), 0, T, N, 0)
    Ws = S // dim * dim
    ctx.match_dev(src.data_ptr(), Ws, Ws, item.data_ptr(), dist.data_ptr()); ctx.sync()
    ts = []
    for _ in range(5):
        ctx.timer_start(); ctx.match_dev(src.data_ptr(), Ws, Ws, item.data_ptr(), dist.data_ptr()); ts.append(ctx.timer_stop())
    ms = float(np.median(ts))
    words = (3 * N + 3) // 4
    pairs = Q * 2 * T
    sad = pairs * words
    print(f"mode {dim:3d} N={N:5d} D={3*N:6d} T={T} Q={Q}: {ms:8.3f} ms  {pairs/ms/1e6:9.1f} Gpairs/s  "
          f"VABSDIFF4 rate {sad/(ms*1e-3)/1e12:6.2f} T/s = {sad/(ms*1e-3)/peak:5.2f} of pipe peak; source px/s {Ws*Ws/(ms*1e-3)/1e6:8.1f} M")
