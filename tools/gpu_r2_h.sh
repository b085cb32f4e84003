#!/bin/bash
# round 2, GPU call H (1 GPU): segment-count sweep on reddit@64 (one wave of long segments?), resident-CSR test + bench legs
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
timeout 600 python -m pytest tests/test_spmm_gpu.py -x -q -m gpu -k "resident or host" > gpurun_out/pytest_h.log 2>&1; tail -3 gpurun_out/pytest_h.log
for CFG in "2 8192" "1 8192" "1 16384" "1 32768" "2 16384" "3 8192"; do
  set -- $CFG
  DGS_SPMM_SEGS=$1 DGS_SPMM_CHUNK_CAP=$2 timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-ref-cuda --no-secondary --no-legs > gpurun_out/bench_segs_$1_$2.log 2>&1
  python - <<PY
import json
l=[x for x in open('gpurun_out/bench_segs_$1_$2.log') if x.startswith('{')]
d=json.loads(l[-1]) if l else None
print('segs $1 cap $2', (d['ms_per_step'], d['roofline']['kernel_ms_avg'], d['roofline']['fixup_ms_avg'], d['parity_ok']) if d else open('gpurun_out/bench_segs_$1_$2.log').read()[-800:])
PY
done
timeout 900 python bench.py > gpurun_out/bench_h.log 2> gpurun_out/bench_h.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_h.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(d['value'], d['ms_per_step'], d['parity_ok'], d['e2e']['ms_per_step'], d.get('e2e_resident_csr'), d['cpu_baseline'].get('torch_sparse_mm'))
else:
    print(open('gpurun_out/bench_h.err').read()[-2000:])
PY
