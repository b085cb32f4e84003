#!/bin/bash
# GPU box: spconv + reference-CUDA parity tests, then the spconv timing table.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_spconv_gpu.py tests/test_vs_reference_cuda_gpu.py -m gpu -q --maxfail=${MAXFAIL:-30} -p no:cacheprovider > gpurun_out/pytest_spconv.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_spconv.log
tail -40 gpurun_out/pytest_spconv.log
timeout 600 python tools/bench_spconv.py --reps 20 --channels "128,128;256,256" > gpurun_out/bench_spconv.log 2>&1
cat gpurun_out/bench_spconv.log | tail -20
