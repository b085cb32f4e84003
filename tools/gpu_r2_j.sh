#!/bin/bash
# round 2, GPU call J (1 GPU): SDDMM whole-wave chunking (tests + arxiv@256 bench line + K sweep), then the sanitizer pass
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
timeout 900 python -m pytest tests/test_sddmm_csr2csc_gpu.py tests/test_vs_reference_cuda_gpu.py tests/test_torch_face_gpu.py -x -q -m gpu > gpurun_out/pytest_j.log 2>&1
tail -3 gpurun_out/pytest_j.log
timeout 300 python bench.py --workload arxiv256 --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_arxiv_j.log 2>&1
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_arxiv_j.log') if x.startswith('{')]
d=json.loads(l[-1]); print('arxiv256', d['ms_per_step'], d['roofline']['kernel_ms_avg'], d['roofline']['frac'], d['reference_cuda']['ms_per_step'], d['reference_cuda']['max_rel_diff_ours_vs_reference_cuda'])
PY
timeout 300 python tools/bench_vs_ref.py --reps 30 2>/dev/null | grep sddmm | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(' ', d['graph'], d['K'], 'ours %.4f ref %.4f x%.2f'%(d['ours_ms'], d['reference_cuda_ms'], d['speedup']))"
bash tools/gpu_r2_sanitize.sh
