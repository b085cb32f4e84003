#!/bin/bash
# on a box with N GPUs: (optional) 2-rank parity test, then one default bench run at N (what the driver's scaling step launches)
N=${1:-8}; TEST=${2:-0}
mkdir -p gpurun_out
if [ "$TEST" = "1" ]; then
  timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/mgpu_test.log 2>&1
  tail -5 gpurun_out/mgpu_test.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 \
     bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}.log 2> gpurun_out/bench_n${N}.err
grep '^{' gpurun_out/bench_n${N}.log; tail -3 gpurun_out/bench_n${N}.err
