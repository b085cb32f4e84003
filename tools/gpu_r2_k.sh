#!/bin/bash
# round 2, GPU call K (1 GPU): smoke, full suite, final ncu captures (full + launch lists) on the final sources, default bench
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_k.log 2>&1
tail -3 gpurun_out/pytest_k.log
for SPEC in "reddit64 spmm_rowseg" "products128 spmm_rowseg" "arxiv256 sddmm_ring"; do
  set -- $SPEC
  CMD="python bench.py --workload $1 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-ref-cuda --no-secondary --no-legs"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file gpurun_out/r02_launches_$1.csv $CMD > gpurun_out/ncu_list_$1.log 2>&1
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 \
      -o gpurun_out/r02_prof_$1 -f $CMD > gpurun_out/ncu_full_$1.log 2>&1
  tail -1 gpurun_out/ncu_full_$1.log
done
timeout 900 python bench.py > gpurun_out/bench_k.log 2> gpurun_out/bench_k.err
timeout 600 python bench.py --workload products128 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_products_k.log 2>&1
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_k.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(d['value'], d['ms_per_step'], d['parity_ok'], d['roofline']['kernel_ms_avg'], d['roofline']['fixup_ms_avg'], d['e2e']['ms_per_step'], d['e2e_resident_csr']['ms_per_step'], d['clocks'])
    for k,v in d['secondary'].items(): print(k, v.get('ms_per_step'), v.get('parity_ok'), (v.get('reference_cuda') or {}).get('ms_per_step'), v.get('error'))
else:
    print(open('gpurun_out/bench_k.err').read()[-2000:])
PY
