#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
timeout 900 python -m pytest tests/test_sddmm_csr2csc_gpu.py tests/test_torch_face_gpu.py -m gpu -q -x -p no:cacheprovider -k "sddmm or backward or masked" > gpurun_out/pytest_x.log 2>&1
tail -3 gpurun_out/pytest_x.log
timeout 600 python tools/exp_sddmm_ring.py --d1 > gpurun_out/exp_sddmm_ring_d1b.jsonl 2> gpurun_out/exp_sddmm_ring_d1b.err
tail -3 gpurun_out/exp_sddmm_ring_d1b.err
python - <<'PY'
import json
for l in open('gpurun_out/exp_sddmm_ring_d1b.jsonl'):
    d = json.loads(l)
    if 'sddmm_chunk' in str(d['setting']) and d['graph'][:5] != 'arxiv': continue
    print(d['graph'][:10], d['K'], str(d['setting'])[:48].ljust(48), 'wpc', d['wpc'], 'x', d['ctas_per_sm'], 'chunk', d['edges_per_warp'], '%.4f ms' % d['ms'], '' if d['bit_identical_to_default'] else 'DIFF')
PY
timeout 300 python tools/exp_sddmm_threads.py > gpurun_out/exp_sddmm_threads_b.jsonl 2>/dev/null; cut -c1-250 gpurun_out/exp_sddmm_threads_b.jsonl
timeout 300 python tools/exp_sddmm_k.py 2>&1 | tail -4
