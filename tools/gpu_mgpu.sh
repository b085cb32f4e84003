#!/bin/bash
# on a box with N GPUs: multi-GPU parity + bench at N for every exchange mode
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/mgpu_test.log 2>&1
tail -5 gpurun_out/mgpu_test.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29650 tests/mgpu_worker.py 2>&1 | grep -v Warning | tail -14
for mode in mcast peer nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29700 \
     bench.py --gpus $N --steps 20 --warmup 5 --mode $mode --no-e2e > gpurun_out/bench_n${N}_$mode.log 2> gpurun_out/bench_n${N}_$mode.err
  cut -c1-700 gpurun_out/bench_n${N}_$mode.log; tail -2 gpurun_out/bench_n${N}_$mode.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 \
     bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}.log 2> gpurun_out/bench_n${N}.err
cat gpurun_out/bench_n${N}.log; tail -3 gpurun_out/bench_n${N}.err
