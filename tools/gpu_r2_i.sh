#!/bin/bash
# round 2, GPU call I (1 GPU): full suite after the wave-quantisation changes, bench line, op sweep, small graphs
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_i.log 2>&1
tail -4 gpurun_out/pytest_i.log
timeout 900 python bench.py > gpurun_out/bench_i.log 2> gpurun_out/bench_i.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_i.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(d['value'], d['ms_per_step'], d['parity_ok'], d['roofline']['kernel_ms_avg'], d['roofline']['fixup_ms_avg'], d['roofline']['traffic'], d['e2e']['ms_per_step'], d['e2e_resident_csr']['ms_per_step'])
    for k,v in d['secondary'].items(): print(k, v.get('ms_per_step'), v.get('parity_ok'), v.get('roofline',{}).get('kernel_ms_avg'), v.get('roofline',{}).get('traffic'), (v.get('reference_cuda') or {}).get('ms_per_step'), v.get('error'))
else:
    print(open('gpurun_out/bench_i.err').read()[-2000:])
PY
DGS_SPMM_PANEL=64 timeout 300 python tools/exp_panels.py reddit 128 2>/dev/null | head -4 | cut -c1-200
DGS_SPMM_PANEL=64 timeout 400 python tools/exp_panels.py products 128 2>/dev/null | head -4 | cut -c1-200
timeout 300 python tools/bench_vs_ref.py --small --reps 50 > gpurun_out/small_default.jsonl 2> gpurun_out/small_default.err
python - <<'PY'
import json
for l in open('gpurun_out/small_default.jsonl'):
    d=json.loads(l); print(' ',d['op'], d['graph'][:14], d.get('N',d.get('K')), 'ours %.1f us ref %.1f us x%.2f'%(d['ours_ms']*1e3, d['reference_cuda_ms']*1e3, d['speedup']))
PY
