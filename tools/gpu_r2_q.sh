#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python tools/bench_vs_ref.py --reps 20 > gpurun_out/vs_ref_full.jsonl 2> gpurun_out/vs_ref_full.err
python - <<'PY'
import json
for l in open('gpurun_out/vs_ref_full.jsonl'):
    d=json.loads(l); print(d['op'], d['graph'], d.get('N', d.get('K')), 'ours %.3f ms ref %.3f ms x%.2f'%(d['ours_ms'], d['reference_cuda_ms'], d['speedup']))
PY
timeout 300 python tools/bench_vs_ref.py --small --reps 50 > gpurun_out/small_default.jsonl 2>/dev/null
