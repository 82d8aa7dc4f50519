timeout 600 python -m pytest tests/test_gpu_no_repeat.py tests/test_host_cpp.py -x -q -s 2>&1 | grep -v "^$" | tail -12
ROWS=512,64 timeout 300 python tools/sweep_match3.py 2>&1 | tee gpurun_out/sweep_match3_variants.txt
# source-level capture of the 4to1 scan
cat > /tmp/c2once.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import emosaic_b200 as emo
ctx = emo.Context(0); dev = torch.device("cuda", 0)
T, S = 10000, 1024
colors = torch.from_numpy(np.random.default_rng(1).integers(0, 256, (T * 12,), dtype=np.uint8)).to(dev)
ctx.set_library_dev(colors.data_ptr(), 0, T, 4, 0)
src = torch.from_numpy(np.random.default_rng(2).integers(0, 256, (S * S * 3,), dtype=np.uint8)).to(dev)
Q = (S // 2) ** 2
item = torch.empty(Q, dtype=torch.int32, device=dev); dist = torch.empty(Q, dtype=torch.int32, device=dev)
for _ in range(3):
    ctx.match_dev(src.data_ptr(), S, S, item.data_ptr(), dist.data_ptr())
ctx.sync()
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:match_kernel -s 2 -c 1 -f -o gpurun_out/prof_match_kernel_c2 python /tmp/c2once.py > gpurun_out/ncu_c2.log 2>&1; echo ncu rc=$?
