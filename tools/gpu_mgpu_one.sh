#!/bin/bash
# on a box with N GPUs: one default bench run at N (what the driver's scaling step launches)
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 \
     bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}.log 2> gpurun_out/bench_n${N}.err
cat gpurun_out/bench_n${N}.log; tail -3 gpurun_out/bench_n${N}.err
