#!/bin/bash
mkdir -p gpurun_out
for wl in reddit64 products128 arxiv256; do
  timeout 900 python bench.py --workload $wl --steps 20 --warmup 5 > gpurun_out/bench_$wl.log 2> gpurun_out/bench_$wl.err
  cat gpurun_out/bench_$wl.log; tail -3 gpurun_out/bench_$wl.err
done
