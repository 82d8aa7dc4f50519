"""Host copy ceiling of the box: pinned-memory D2H / H2D with 1..n GPUs copying at once (one process, one thread per GPU).
Run under gpurun --gpus N.  Prints one JSON line per configuration."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402  (device count only)

from tools import probe  # noqa: E402

n_all = torch.cuda.device_count()
size = 400 << 20
for n in sorted({1, 2, 4, n_all} & set(range(1, n_all + 1))):
    for direction in ("d2h", "h2d"):
        for piece in (16 << 20, 64 << 20, 256 << 20):
            agg, per = probe.host_copy(list(range(n)), size, piece, direction, 3)
            print(json.dumps({"gpus": n, "dir": direction, "piece_mb": piece >> 20, "aggregate_gbs": round(agg, 1),
                              "per_gpu_gbs": [round(x, 1) for x in per]}), flush=True)
if n_all >= 2:  # which pairs share an uplink: GPU 0 together with each other GPU
    for j in range(1, n_all):
        agg, per = probe.host_copy([0, j], size, 64 << 20, "d2h", 3)
        print(json.dumps({"pair": [0, j], "dir": "d2h", "aggregate_gbs": round(agg, 1), "per_gpu_gbs": [round(x, 1) for x in per]}), flush=True)
