#!/bin/bash
# Run under gpurun (1 GPU): plain bench first, then the ncu launch list and one --set full capture per hot kernel.
# Outputs land in gpurun_out/; copy the summaries you want judged into profiles/ (tools/summarise_ncu.py).
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu"
timeout 600 $CMD > gpurun_out/plain.json 2> gpurun_out/plain.err
rc=$?; echo "plain rc=$rc"; [ $rc -ne 0 ] && { tail -5 gpurun_out/plain.err; exit $rc; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
for K in compose_tile_kernel match_index_kernel index_sweep_r_kernel index_sweep_kernel match_kernel compose_copy_kernel analyse_fast_kernel compose_tint_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/prof_$K $CMD > gpurun_out/ncu_$K.log 2>&1
  echo "$K rc=$?"
done
ls -la gpurun_out/
