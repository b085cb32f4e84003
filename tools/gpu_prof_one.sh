#!/bin/bash
# usage: gpu_prof_one.sh <tag> <workload> <kernel-regex>   -> gpurun_out/prof_<tag>_<workload>.ncu-rep
TAG=$1; WL=$2; KRE=$3
mkdir -p gpurun_out
CMD="python bench.py --workload $WL --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-ref-cuda"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 3 -c 1 \
    -o gpurun_out/prof_${TAG}_${WL} -f $CMD > gpurun_out/ncu_full_${TAG}_${WL}.log 2>&1
tail -3 gpurun_out/ncu_full_${TAG}_${WL}.log
