#!/bin/bash
# GPU box: SDDMM parity + config-4 bench (with the reference CUDA leg) + ncu capture of the SDDMM kernel.
TAG=${1:-r01c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sddmm_csr2csc_gpu.py tests/test_vs_reference_cuda_gpu.py tests/test_torch_face_gpu.py -m gpu -q -k "sddmm or forward_backward" --maxfail=20 -p no:cacheprovider > gpurun_out/pytest_sddmm.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_sddmm.log
tail -15 gpurun_out/pytest_sddmm.log
timeout 600 python bench.py --workload arxiv256 --steps 50 --warmup 5 > gpurun_out/bench_arxiv256.log 2> gpurun_out/bench_arxiv256.err
cat gpurun_out/bench_arxiv256.log; tail -3 gpurun_out/bench_arxiv256.err
DGS_SDDMM_NO_RING=1 timeout 600 python bench.py --workload arxiv256 --steps 50 --warmup 5 --no-ref-cuda > gpurun_out/bench_arxiv256_noring.log 2>&1
cat gpurun_out/bench_arxiv256_noring.log | tail -1
CMD="python bench.py --workload arxiv256 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv \
    --log-file gpurun_out/launches_${TAG}_arxiv256.csv $CMD > gpurun_out/ncu_list_${TAG}_arxiv256.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sddmm -s 3 -c 2 \
    -o gpurun_out/prof_${TAG}_arxiv256 -f $CMD > gpurun_out/ncu_full_${TAG}_arxiv256.log 2>&1
