"""Launch-shape sweep for the 4to1 scan (match_kernel<3, R, NT>) on C2 and its row stripes (run under gpurun).
One process per setting: the overrides are read once per process."""
import os, sys, subprocess, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == "one":
    import torch
    import emosaic_b200 as emo
    from tools.probe import probe_int_pipe
    ctx = emo.Context(0)
    dev = torch.device("cuda", 0)
    T, S = 10000, 1024
    colors = torch.from_numpy(np.random.default_rng(1).integers(0, 256, (T * 12,), dtype=np.uint8)).to(dev)
    ctx.set_library_dev(colors.data_ptr(), 0, T, 4, 0)
    sad = probe_int_pipe(0, 1)
    res = {}
    for rows in [int(x) for x in os.environ.get("ROWS", "512,256,128,64").split(",")]:
        H = rows * 2
        src = torch.from_numpy(np.random.default_rng(2).integers(0, 256, (H * S * 3,), dtype=np.uint8)).to(dev)
        Q = (S // 2) * rows
        item = torch.empty(Q, dtype=torch.int32, device=dev); dist = torch.empty(Q, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        for _ in range(3):
            ctx.match_dev(src.data_ptr(), S, H, item.data_ptr(), dist.data_ptr())
        ctx.sync()
        ts = []
        for _ in range(7):
            ctx.timer_start(); ctx.match_dev(src.data_ptr(), S, H, item.data_ptr(), dist.data_ptr()); ts.append(ctx.timer_stop())
        ms = float(np.median(ts))
        res[rows] = (ms, 3 * Q * 2 * T / (ms * 1e-3) / sad, int(item.sum().item()) & 0xffffff)
    print(json.dumps(res))
    sys.exit(0)

base = None
for shape in os.environ.get("SHAPES", "2128,1,2,3,4,5,6,7").split(","):
    for minlen in os.environ.get("MINLENS", "0").split(","):
        env = dict(os.environ, EMO_MATCH_SHAPE3=shape, EMO_MATCH_MINLEN=minlen)
        r = subprocess.run([sys.executable, __file__, "one"], env=env, capture_output=True, text=True)
        try:
            res = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception:
            print(shape, minlen, "FAILED", r.stderr[-300:]); continue
        chk = {k: v[2] for k, v in res.items()}
        base = base or chk
        print(f"R,NT={shape} minlen={minlen:>5}: " + "  ".join(f"rows {k}: {v[0]*1e3:7.1f} us frac {v[1]:.3f}" for k, v in res.items()) +
              ("" if chk == base else "  CHECKSUM MISMATCH"), flush=True)
