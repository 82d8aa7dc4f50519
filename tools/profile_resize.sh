set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 100 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 100 python tools/bench_resize.py > gpurun_out/bench_resize_v3.txt 2>&1; cat gpurun_out/bench_resize_v3.txt
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/resize_launches_v3.csv python tools/bench_resize.py > /dev/null 2>&1; echo "list rc=$?"
for spec in "resize_vertical_kernel:4:vt" "resize_horizontal_t_kernel:4:ht" "resize_vertical_kernel:0:v0" "resize_horizontal_kernel:0:h0"; do
  K=${spec%%:*}; rest=${spec#*:}; C=${rest%%:*}; TAG=${rest#*:}
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/prof_resize_$TAG python tools/bench_resize.py $C > gpurun_out/ncu_resize_$TAG.log 2>&1; echo "$TAG rc=$?"
done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:topk_kernel -s 1 -c 1 -f -o gpurun_out/prof_topk python tools/bench_topk.py > gpurun_out/ncu_topk.log 2>&1; echo "topk rc=$?"
timeout 400 python bench.py > gpurun_out/r01h_bench_n1.json 2> gpurun_out/r01h_bench_n1.err; echo "bench rc=$?"; head -c 600 gpurun_out/r01h_bench_n1.json
