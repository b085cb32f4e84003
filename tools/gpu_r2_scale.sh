#!/bin/bash
# usage: gpu_r2_scale.sh N   -> the driver's SCALE command at N GPUs, final sources
N=$1
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29681 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_n$N.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('N=$N', d['config']['parallelism'], 'value', d['value'], 'ms', d['ms_per_step'], 'parity', d.get('parity_ok'), 'kernel', d['roofline']['kernel_ms_avg'], 'fixup', d['roofline']['fixup_ms_avg'], 'e2e', d['e2e']['ms_per_step'])
    print({k: v for k, v in d['legs'].items() if k != 'note'})
else:
    print(open('gpurun_out/bench_n$N.err').read()[-3000:])
PY
