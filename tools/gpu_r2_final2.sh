#!/bin/bash
# round 2, last multi-GPU check on the final sources (2 GPUs): N=1 default bench (traffic now keyed to the final sources),
# the 2-rank parity test, the N=2 bench as the driver launches it, and the reference arm
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
timeout 900 python bench.py > gpurun_out/bench_final_n1.log 2> gpurun_out/bench_final_n1.err
timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/mgpu_test.log 2>&1; tail -3 gpurun_out/mgpu_test.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 \
     bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_final_n2.log 2> gpurun_out/bench_final_n2.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_ref.log 2> gpurun_out/bench_final_ref.err
python - <<'PY'
import json
for f in ('bench_final_n1','bench_final_n2','bench_final_ref'):
    l=[x for x in open(f'gpurun_out/{f}.log') if x.startswith('{')]
    if l:
        d=json.loads(l[-1]); print(f, d.get('value'), d.get('ms_per_step'), d.get('parity_ok'), (d.get('roofline') or {}).get('traffic'), (d.get('e2e') or {}).get('value'), d.get('clocks'))
    else:
        print(f, open(f'gpurun_out/{f}.err').read()[-1500:])
PY
