#!/bin/bash
# round 2, GPU call F (1 GPU): full suite (-s to keep the printed reference comparisons), spconv vs reference, small graphs
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
timeout 1800 python -m pytest tests -x -q -m gpu -s > gpurun_out/pytest_f.log 2>&1
tail -4 gpurun_out/pytest_f.log
grep -E "driver-reported|reference backward" gpurun_out/pytest_f.log | cut -c1-600
timeout 600 python tools/bench_spconv.py --reps 20 --channels "128,128" > gpurun_out/spconv_vs_ref.jsonl 2> gpurun_out/spconv_vs_ref.err
timeout 300 python tools/bench_vs_ref.py --small --reps 50 > gpurun_out/small_default.jsonl 2> gpurun_out/small_default.err
DGS_SDDMM_STAGES=3 timeout 300 python tools/bench_vs_ref.py --small --reps 50 > gpurun_out/small_stages3.jsonl 2> gpurun_out/small_stages3.err
python - <<'PY'
import json
for l in open('gpurun_out/spconv_vs_ref.jsonl'):
    d=json.loads(l); print(d['case'], d['precision'], 'fwd %.3f dx %.3f dw %.3f'%(d['fwd_ms'],d['dx_ms'],d['dw_ms']), 'ref fwd', d.get('reference_fwd_ms'), 'ref bwd', d.get('reference_bwd_dx_plus_dw_ms'))
for f in ('small_default','small_stages3'):
    print(f)
    for l in open(f'gpurun_out/{f}.jsonl'):
        d=json.loads(l)
        if d['op']=='sddmm_csr' or 'Gnutella' in d['graph']: print(' ',d['op'], d['graph'][:14], d.get('N',d.get('K')), 'ours %.1f us ref %.1f us x%.2f'%(d['ours_ms']*1e3, d['reference_cuda_ms']*1e3, d['speedup']))
PY
