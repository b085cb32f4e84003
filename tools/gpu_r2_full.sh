#!/bin/bash
# round 2 checkpoint: smoke + the whole GPU suite + default bench on the current sources
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_full.log 2>&1
tail -3 gpurun_out/pytest_full.log
timeout 900 python bench.py > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_full.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(d['value'], d['ms_per_step'], d['parity_ok'], d['roofline']['traffic'], d['e2e']['ms_per_step'], d['clocks'])
    for k,v in d['secondary'].items(): print(k, v.get('ms_per_step'), v.get('parity_ok'), (v.get('reference_cuda') or {}).get('ms_per_step'), v.get('error'))
else:
    print(open('gpurun_out/bench_full.err').read()[-2000:])
PY
