#!/bin/bash
# SDDMM ring: warps per CTA x CTAs per SM sweep with the residency taken from the occupancy API
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_sddmm_csr2csc_gpu.py tests/test_spmm_gpu.py -m gpu -q -x -p no:cacheprovider -k "sddmm or colmajor" > gpurun_out/pytest_t.log 2>&1
tail -3 gpurun_out/pytest_t.log
timeout 600 python tools/exp_sddmm_ring.py --chunks > gpurun_out/exp_sddmm_ring.jsonl 2> gpurun_out/exp_sddmm_ring.err
tail -3 gpurun_out/exp_sddmm_ring.err
python - <<'PY'
import json
for l in open('gpurun_out/exp_sddmm_ring.jsonl'):
    d = json.loads(l)
    print(d['graph'][:10], d['K'], str(d['setting'])[:44].ljust(44), 'wpc', d['wpc'], 'x', d['ctas_per_sm'], 'chunk', d['edges_per_warp'], '%.4f ms' % d['ms'], '' if d['bit_identical_to_default'] else 'DIFF')
PY
