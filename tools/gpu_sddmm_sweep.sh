#!/bin/bash
mkdir -p gpurun_out
for ps in 1 2 3 4; do
  echo "PASSES=$ps"
  DGS_SDDMM_PASSES=$ps timeout 300 python bench.py --workload arxiv256 --steps 50 --warmup 5 --no-ref-cuda 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms_avg'])"
done
echo default; timeout 300 python bench.py --workload arxiv256 --steps 50 --warmup 5 --no-ref-cuda 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms_avg'])"
DGS_SDDMM_PASSES=3 timeout 600 python -m pytest tests/test_sddmm_csr2csc_gpu.py -m gpu -q -k sddmm -p no:cacheprovider 2>&1 | tail -2
timeout 600 python -m pytest tests/test_sddmm_csr2csc_gpu.py tests/test_vs_reference_cuda_gpu.py tests/test_torch_face_gpu.py -m gpu -q -k "sddmm or forward_backward" -p no:cacheprovider 2>&1 | tail -2
