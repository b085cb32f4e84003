#!/bin/bash
# GPU box: ncu launch list + full capture of the spconv tile kernel on the MinkUNet layer maps.
TAG=${1:-r01}
mkdir -p gpurun_out
CMD="python tools/bench_spconv.py --reps 2"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file gpurun_out/launches_${TAG}_spconv.csv $CMD > gpurun_out/ncu_list_${TAG}_spconv.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spconv_fgms_pipe -s 12 -c 2 \
    -o gpurun_out/prof_${TAG}_spconv -f $CMD > gpurun_out/ncu_full_${TAG}_spconv.log 2>&1
ls -la gpurun_out | tail -5
