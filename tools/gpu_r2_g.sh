#!/bin/bash
# round 2, GPU call G (8 GPUs): the config-5 bench line (weak scaling, parity_ok, legs incl. the strong-scaling leg)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29671 \
    bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8.log 2> gpurun_out/bench_n8.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_n8.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(d['config']['parallelism'], 'value', d['value'], 'ms', d['ms_per_step'], 'parity', d.get('parity_ok'), 'kernel', d['roofline']['kernel_ms_avg'], 'fixup', d['roofline']['fixup_ms_avg'])
    print(d.get('legs')); print(d.get('e2e'))
else:
    print(open('gpurun_out/bench_n8.err').read()[-3000:])
PY
