#!/bin/bash
# round 2, GPU call C (1 GPU): full suite, default bench line (with configs 3 / 4 appended), op latency, locality experiment
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_c.log 2>&1
tail -6 gpurun_out/pytest_c.log
timeout 900 python bench.py > gpurun_out/bench_c.log 2> gpurun_out/bench_c.err
tail -c 1500 gpurun_out/bench_c.err
timeout 300 python tools/bench_call_latency.py > gpurun_out/call_latency.jsonl 2> gpurun_out/call_latency.err
timeout 300 python tools/bench_vs_ref.py --small --reps 50 > gpurun_out/small_default.jsonl 2> gpurun_out/small_default.err
timeout 600 python tools/exp_locality.py 64 > gpurun_out/locality64.jsonl 2> gpurun_out/locality64.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,lts__throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:spmm_rowseg --csv --log-file gpurun_out/ncu_locality64.csv \
    python tools/exp_locality.py 64 > gpurun_out/ncu_locality64.log 2>&1
cat gpurun_out/call_latency.jsonl
cat gpurun_out/locality64.jsonl
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_c.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1])
    print({k:d[k] for k in ('value','ms_per_step','parity_ok') if k in d}, d.get('roofline',{}).get('frac'), d.get('e2e',{}).get('ms_per_step'))
    for k,v in (d.get('secondary') or {}).items(): print(k, {kk:v.get(kk) for kk in ('ms_per_step','parity_ok','error')}, (v.get('roofline') or {}).get('frac'), (v.get('reference_cuda') or {}).get('ms_per_step'))
    print(d.get('legs'))
PY
