#!/bin/bash
# round 2, GPU call N: SDDMM ring, 2 CTAs x 11 warps per SM for rows <= 512 B vs 1 x 16
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
timeout 600 python -m pytest tests/test_sddmm_csr2csc_gpu.py tests/test_vs_reference_cuda_gpu.py -x -q -m gpu > gpurun_out/pytest_n.log 2>&1; tail -2 gpurun_out/pytest_n.log
for WPC in 0 16; do
  echo "sddmm_wpc=$WPC (0 = default rule)"
  DGS_SDDMM_WPC=$WPC timeout 300 python tools/bench_vs_ref.py --small --reps 50 2>/dev/null | grep sddmm | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('   ', d['graph'][:12], d['K'], 'ours %.1f ref %.1f x%.2f'%(d['ours_ms']*1e3, d['reference_cuda_ms']*1e3, d['speedup']))"
  DGS_SDDMM_WPC=$WPC timeout 300 python tools/exp_sddmm_k.py 2>/dev/null
done
