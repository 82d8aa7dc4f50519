timeout 600 python -m pytest tests/test_gpu_index.py -x -q 2>&1 | tail -2
echo "new build"; timeout 200 python tools/bench_index_build.py
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:index_ -c 20 --csv --log-file gpurun_out/r02j_index_launches.csv python tools/bench_index_build.py > /dev/null 2>&1; echo "list rc=$?"
python tools/summarise_ncu.py launches gpurun_out/r02j_index_launches.csv
