#!/bin/bash
# Runs on the GPU box: ncu launch list of the bench command + one full capture of the top kernel.
# usage: bash tools/gpu_profile.sh <tag> [workload] [kernel-regex]
TAG=${1:-r01}; WL=${2:-reddit64}; KRE=${3:-spmm_rowseg}
mkdir -p gpurun_out
CMD="python bench.py --workload $WL --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_${TAG}_${WL}.csv $CMD > gpurun_out/ncu_list_${TAG}_${WL}.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 3 -c 2 \
    -o gpurun_out/prof_${TAG}_${WL} -f $CMD > gpurun_out/ncu_full_${TAG}_${WL}.log 2>&1
ls -la gpurun_out | tail -8
