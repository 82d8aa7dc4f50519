timeout 600 python -m pytest tests/test_gpu_index.py tests/test_gpu_warhol.py tests/test_gpu_cli.py "tests/test_gpu_parity.py::test_config4_full_map" -x -q --durations=5 2>&1 | tail -12
for T in 100000 30000; do T=$T MODE=auto timeout 120 python tools/bench_stripes.py 2>&1 | tail -4; done
for P in 40 64; do EMO_L2_PERSIST=$P T=100000 MODE=auto NS=1,8 timeout 120 python tools/bench_stripes.py 2>&1 | tail -2; done
T=100000 NS=8 timeout 300 ncu --set full --clock-control none --import-source on -k regex:match_index16 -s 30 -c 1 -f -o gpurun_out/prof_match_index16_rows512 python tools/bench_stripes.py > gpurun_out/ncu_idx16.log 2>&1; echo ncu rc=$?
T=100000 NS=8 timeout 300 ncu --set full --clock-control none --import-source on -k regex:compose_tile -s 30 -c 1 -f -o gpurun_out/prof_compose_tile_rows512 python tools/bench_stripes.py > gpurun_out/ncu_comp512.log 2>&1; echo ncu rc=$?
