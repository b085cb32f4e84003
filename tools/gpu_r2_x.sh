#!/bin/bash
# round 2, GPU call X (1 GPU): spconv with TMA bulk-copied weight panels — parity, timing against the reference, sanitizer;
# plus the CUDA-graph capture test of SpMM
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
timeout 1500 python -m pytest tests/test_spconv_gpu.py tests/test_vs_reference_spconv_gpu.py tests/test_nn_gpu.py -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/pytest_x.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_x.log
timeout 300 python -m pytest tests/test_spmm_gpu.py -m gpu -q -x --timeout 120 -k "cuda_graph or graph_note" -p no:cacheprovider 2>&1 | tail -3
timeout 900 python tools/bench_spconv.py --reps 20 > gpurun_out/bench_spconv_x.log 2>&1; tail -24 gpurun_out/bench_spconv_x.log | cut -c1-400
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 0 python -m pytest tests/test_spconv_gpu.py -x -q -m gpu -k "forward_random_maps or half_inputs" -p no:cacheprovider > gpurun_out/sanitizer_r02_spconv_tma.log 2>&1
echo "memcheck exit $?" >> gpurun_out/sanitizer_r02_spconv_tma.log
grep -E "ERROR SUMMARY|passed|failed|memcheck exit" gpurun_out/sanitizer_r02_spconv_tma.log | tail -4
