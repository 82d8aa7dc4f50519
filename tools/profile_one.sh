#!/bin/bash
# usage: bash tools/profile_one.sh <kernel-regex> [<kernel-regex> ...]   (under gpurun, 1 GPU)
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu"
timeout 600 $CMD > gpurun_out/plain.json 2> gpurun_out/plain.err
rc=$?; echo "plain rc=$rc"; [ $rc -ne 0 ] && { tail -5 gpurun_out/plain.err; exit $rc; }
for K in "$@"; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/prof_$K $CMD > gpurun_out/ncu_$K.log 2>&1
  echo "$K rc=$?"
done
