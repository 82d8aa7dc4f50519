"""emo_resize (Lanczos3, image 0.25.2 bit-exact) throughput on a few geometries (run under gpurun); the whole call (vertical +
horizontal pass) is timed with CUDA events, the split per kernel comes from the ncu launch list (tools/summarise_ncu.py launches).
usage: python tools/bench_resize.py [case index ...]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import emosaic_b200 as emo
ctx = emo.Context(0)
dev = torch.device("cuda", 0)
CASES = [  # n, h, w, nh, nw
    (1, 4097, 4098, 4096, 4096),      # source, not divisible (main.rs:567-595), unaligned rows
    (1, 4096, 4100, 4096, 4096),      # same with 4-byte aligned rows
    (1, 8192, 8192, 4096, 4096),      # --downsample 2
    (1, 3000, 4000, 64, 64),          # one 12 MP photo -> tile
    (64, 2048, 2048, 64, 64),         # a batch of photos -> tiles
    (256, 1024, 1024, 16, 16),
    (4096, 256, 256, 8, 8),
    (1, 1024, 1024, 4096, 4096),      # upscale
]
sel = [int(a) for a in sys.argv[1:]]
for n, h, w, nh, nw in ([CASES[i] for i in sel] if sel else CASES):
    imgs = torch.randint(0, 256, (n * h * w * 3,), dtype=torch.uint8, device=dev)
    out = torch.empty(n * nh * nw * 3, dtype=torch.uint8, device=dev)
    fn = lambda: ctx.resize_dev(imgs.data_ptr(), n, w, h, None, nw, nh, out.data_ptr())
    fn(); ctx.sync()
    t = []
    for _ in range(7):
        ctx.timer_start(); fn(); t.append(ctx.timer_stop())
    ms = float(np.median(t))
    byt = n * (h * w + nh * nw) * 3
    print(f"n={n:5d} {w}x{h} -> {nw}x{nh}: {ms:8.3f} ms  {byt/(ms*1e-3)/1e9:8.1f} GB/s algorithmic  {n*h*w/(ms*1e-3)/1e9:7.2f} Gpx/s in")
    del imgs, out
