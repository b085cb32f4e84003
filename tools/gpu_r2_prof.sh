#!/bin/bash
# round 2, final profiling call (1 GPU): launch lists + one `ncu --set full` capture per headline kernel, SDDMM chunk sweep
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
for SPEC in "reddit64 spmm_rowseg" "products128 spmm_rowseg" "arxiv256 sddmm_ring"; do
  set -- $SPEC
  CMD="python bench.py --workload $1 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-ref-cuda --no-secondary --no-legs"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file gpurun_out/r02_launches_$1.csv $CMD > gpurun_out/ncu_list_$1.log 2>&1
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 \
      -o gpurun_out/r02_prof_$1 -f $CMD > gpurun_out/ncu_full_$1.log 2>&1
  tail -2 gpurun_out/ncu_full_$1.log
done
# the latency-regime kernel on the reference's published configuration
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_rowpar -s 5 -c 1 -o gpurun_out/r02_prof_gnutella32_rowpar -f \
    python tools/bench_vs_ref.py --small --reps 5 > gpurun_out/ncu_full_gnutella.log 2>&1
for CH in 32 64 96 128; do
  DGS_SDDMM_CHUNK=$CH timeout 300 python tools/bench_vs_ref.py --small --reps 50 2>/dev/null | grep sddmm > gpurun_out/small_sddmm_chunk$CH.jsonl
done
python - <<'PY'
import json
for ch in (32,64,96,128):
    print('chunk',ch, [ (json.loads(l)['graph'][:6], json.loads(l)['K'], round(json.loads(l)['ours_ms']*1e3,1), round(json.loads(l)['reference_cuda_ms']*1e3,1)) for l in open(f'gpurun_out/small_sddmm_chunk{ch}.jsonl')])
PY
ls -la gpurun_out/*.ncu-rep
