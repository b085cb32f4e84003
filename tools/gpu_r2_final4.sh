#!/bin/bash
# final tables on the final sources: ours vs the reference's CUDA across widths (full-size graphs), and the default bench line
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
timeout 900 python tools/bench_vs_ref.py --reps 20 > gpurun_out/vs_ref_final4.jsonl 2> gpurun_out/vs_ref_final4.err
python - <<'PY'
import json
for l in open('gpurun_out/vs_ref_final4.jsonl'):
    d=json.loads(l); print(d['op'], d['graph'], d.get('N', d.get('K')), 'ours %.4f ms ref %.4f ms x%.2f'%(d['ours_ms'], d['reference_cuda_ms'], d['speedup']))
PY
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final4.log 2> gpurun_out/bench_final4.err; echo "bench rc=$?"
