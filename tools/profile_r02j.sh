#!/bin/bash
# Round-2 final profiling pass (under gpurun, 1 GPU): plain bench first, then the ncu launch list, then one --set full capture per
# hot kernel of the bench and of the resize.  tools/summarise_ncu.py turns the reports into profiles/*.txt and profiles/traffic.json.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
timeout 600 $CMD > gpurun_out/plain.json 2> gpurun_out/plain.err
rc=$?; echo "plain rc=$rc"; [ $rc -ne 0 ] && { tail -5 gpurun_out/plain.err; exit $rc; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
cap() {  # name, demangled-name regex, launches to skip
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 -f -o gpurun_out/prof_$1 $CMD > gpurun_out/ncu_$1.log 2>&1
  echo "$1 rc=$?"
}
cap compose_tile_kernel 'compose_tile_kernel' 2
cap match_index16_kernel 'match_index16_kernel' 2
cap match_kernel_c4 'match_kernel<(\(int\))?1, (\(int\))?8, (\(int\))?256' 1
cap match_kernel_c2 'match_kernel<(\(int\))?3, (\(int\))?2, (\(int\))?128' 2
cap analyse_fast_kernel 'analyse_fast_kernel' 1
cap compose_tint_kernel 'compose_tint_kernel' 1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:resize_vertical2 -s 1 -c 1 -f -o gpurun_out/prof_resize_vertical2 python tools/bench_resize.py 4 > gpurun_out/ncu_resize_v2.log 2>&1; echo "resize rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:index_sweep_reg -s 2 -c 1 -f -o gpurun_out/prof_index_sweep_reg python tools/bench_index_build.py > gpurun_out/ncu_index_sweep.log 2>&1; echo "index sweep rc=$?"
timeout 200 python tools/bench_resize.py > gpurun_out/bench_resize.txt 2>&1; cat gpurun_out/bench_resize.txt
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
