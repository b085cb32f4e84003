#!/bin/bash
# round 2, GPU call D (2 GPUs): multi-GPU parity (incl. the slow-consumer stress), bench at N=2 in every exchange mode
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
timeout 900 python -m pytest tests/test_multigpu_gpu.py -x -q -m gpu > gpurun_out/pytest_d.log 2>&1
tail -15 gpurun_out/pytest_d.log
for MODE in mcast peer nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 \
      bench.py --gpus 2 --steps 20 --warmup 5 --mode $MODE --no-e2e > gpurun_out/bench_n2_$MODE.log 2> gpurun_out/bench_n2_$MODE.err
  python - <<PY
import json
l=[x for x in open('gpurun_out/bench_n2_$MODE.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('$MODE', d['config']['parallelism'], d['value'], d['ms_per_step'], d.get('parity_ok'), d['roofline']['kernel_ms_avg'], d['roofline']['fixup_ms_avg'], d.get('legs'))
else:
    print('$MODE: no line'); print(open('gpurun_out/bench_n2_$MODE.err').read()[-1500:])
PY
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29656 \
    bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2_default.log 2> gpurun_out/bench_n2_default.err
tail -c 400 gpurun_out/bench_n2_default.log
