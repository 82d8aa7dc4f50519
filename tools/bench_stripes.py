"""C4 step time on stripes of H/1, H/2, H/4, H/8 source rows on ONE GPU (the per-rank work of an N-GPU run: the loop has no
collective, so this predicts strong scaling).  Run under gpurun; EMO_L2_HINTS selects experimental cache-hint variants."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import emosaic_b200 as emo

T, ts, W, H = int(os.environ.get("T", "100000")), 8, 4096, 4096
MODE = os.environ.get("MODE", "auto")   # auto | index_wide | index_compact
ctx = emo.Context(0)
dev = torch.device("cuda", 0)
tiles = torch.from_numpy(np.random.default_rng(1234).integers(0, 256, (T * ts * ts * 3,), dtype=np.uint8)).to(dev)
if os.environ.get("SRC", "uniform").startswith("photo"):   # SRC=photo:<noise>: smooth RGB gradient + uniform noise of +-<noise> per channel
    noise = int(os.environ["SRC"].split(":")[1]) if ":" in os.environ["SRC"] else 6
    yy, xx = np.mgrid[0:H, 0:W]
    ph = np.stack([xx * 255 // (W - 1), yy * 255 // (H - 1), (xx + yy) * 255 // (W + H - 2)], -1)
    ph = np.clip(ph + np.random.default_rng(99).integers(-noise, noise + 1, ph.shape), 0, 255).astype(np.uint8)
    src = torch.from_numpy(ph.reshape(-1)).to(dev)
else:
    src = torch.from_numpy(np.random.default_rng(5678).integers(0, 256, (H * W * 3,), dtype=np.uint8)).to(dev)
colors = torch.empty(T * 3, dtype=torch.uint8, device=dev)
item = torch.empty(H * W, dtype=torch.int32, device=dev); dist = torch.empty(H * W, dtype=torch.int32, device=dev)
out = torch.empty(H * ts * W * ts * 3, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
ctx.analyse_dev(tiles.data_ptr(), T, ts, 1, colors.data_ptr())
ctx.set_library_dev(colors.data_ptr(), tiles.data_ptr(), T, 1, ts)
ctx.set_match_mode(MODE)
ctx.build_index(); ctx.sync()
fused = os.environ.get("FUSED", "0") == "1"   # one emo_mosaic_dev call per step instead of match_dev + compose_dev
base = None
for n in ([int(x) for x in os.environ['NS'].split(',')] if os.environ.get('NS') else (1, 2, 4, 8)):
    Hs = H // n
    def step(k=None):
        if fused:
            ctx.mosaic_dev(src.data_ptr(), W, Hs, 3, 0, item.data_ptr(), dist.data_ptr(), out.data_ptr())
            return
        if k is not None: ctx.mark(3 * k)
        ctx.match_dev(src.data_ptr(), W, Hs, item.data_ptr(), dist.data_ptr())
        if k is not None: ctx.mark(3 * k + 1)
        ctx.compose_dev(item.data_ptr(), 0, W, Hs, 3, 0, out.data_ptr())
        if k is not None: ctx.mark(3 * k + 2)
    for _ in range(5): step()
    ctx.sync()
    K = 50
    ctx.timer_start()
    for k in range(K): step(k)
    ms = ctx.timer_stop() / K
    if fused:
        m = c = float("nan")
    else:
        m = np.mean([ctx.mark_elapsed(3 * k, 3 * k + 1) for k in range(K)]); c = np.mean([ctx.mark_elapsed(3 * k + 1, 3 * k + 2) for k in range(K)])
    base = base or ms
    print(f"src={os.environ.get('SRC', 'uniform')} T={T} mode={MODE} fused={int(fused)} rows={Hs:5d} (N={n}): step {ms*1e3:7.1f} us  match {m*1e3:6.1f} us  compose {c*1e3:6.1f} us  "
          f"-> {H*W/n/ms/1e6:6.2f} G px/s per GPU, strong-scaling efficiency {base/(ms*n):.3f}")
