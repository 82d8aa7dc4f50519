timeout 600 python -m pytest tests/test_gpu_resize.py -x -q 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_index.py -x -q -k "mosaic_dev" 2>&1 | tail -2
for G in 0 1; do echo "EMO_GRAPH=$G"; EMO_GRAPH=$G FUSED=1 timeout 300 python tools/bench_stripes.py; done
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:resize_ -c 40 --csv --log-file gpurun_out/r02j_resize_launches.csv python tools/bench_resize.py 4 > /dev/null 2>&1; echo "list rc=$?"
python tools/summarise_ncu.py launches gpurun_out/r02j_resize_launches.csv
