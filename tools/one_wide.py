import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import emosaic_b200 as emo
ctx = emo.Context(0); dev = torch.device("cuda", 0)
N, T, S = 64, 10000, 2048
colors = torch.from_numpy(np.random.default_rng(1).integers(0, 256, (T * N * 3,), dtype=np.uint8)).to(dev)
src = torch.from_numpy(np.random.default_rng(2).integers(0, 256, (S * S * 3,), dtype=np.uint8)).to(dev)
Q = (S // 8) ** 2
item = torch.empty(Q, dtype=torch.int32, device=dev); dist = torch.empty(Q, dtype=torch.int32, device=dev)
torch.cuda.synchronize()
ctx.set_library_dev(colors.data_ptr(), 0, T, N, 0)
for _ in range(3):
    ctx.match_dev(src.data_ptr(), S, S, item.data_ptr(), dist.data_ptr())
ctx.sync()
print("ok")
