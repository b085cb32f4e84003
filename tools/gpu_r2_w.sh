#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
timeout 900 python -m pytest tests/test_sddmm_csr2csc_gpu.py -m gpu -q -x -p no:cacheprovider -k "sddmm" > gpurun_out/pytest_w.log 2>&1
tail -5 gpurun_out/pytest_w.log
timeout 600 python tools/exp_sddmm_ring.py --d1 > gpurun_out/exp_sddmm_ring_d1.jsonl 2> gpurun_out/exp_sddmm_ring_d1.err
tail -3 gpurun_out/exp_sddmm_ring_d1.err
python - <<'PY'
import json
for l in open('gpurun_out/exp_sddmm_ring_d1.jsonl'):
    d = json.loads(l)
    print(d['graph'][:10], d['K'], str(d['setting'])[:48].ljust(48), 'wpc', d['wpc'], 'x', d['ctas_per_sm'], 'chunk', d['edges_per_warp'], '%.4f ms' % d['ms'], '' if d['bit_identical_to_default'] else 'DIFF')
PY
