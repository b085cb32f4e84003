#!/bin/bash
# 8-GPU box: the default bench line (multicast epilogue, with e2e) and the peer-store epilogue beside it
N=${1:-8}
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 \
     bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}.log 2> gpurun_out/bench_n${N}.err
cat gpurun_out/bench_n${N}.log; tail -3 gpurun_out/bench_n${N}.err
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29700 \
     bench.py --gpus $N --steps 20 --warmup 5 --mode peer --no-e2e > gpurun_out/bench_n${N}_peer.log 2> gpurun_out/bench_n${N}_peer.err
cut -c1-300 gpurun_out/bench_n${N}_peer.log; tail -2 gpurun_out/bench_n${N}_peer.err
