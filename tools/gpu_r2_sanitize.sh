#!/bin/bash
# round 2: compute-sanitizer memcheck over the kernels added this round (row-parallel SpMM, device-side segment layout of the
# legacy entry, resident-CSR host path, tiled exact-fp32 spconv, fp16 spconv operands, kmap range flag, SDDMM tiny cases)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 0 \
  python -m pytest tests/test_spmm_gpu.py tests/test_sddmm_csr2csc_gpu.py tests/test_spconv_gpu.py tests/test_kmap_gpu.py -x -q -m gpu \
  -k "row_parallel_kernel_forced or legacy_spmm_cuda or tiny_and_hub or forward_random_maps or half_inputs or out_of_range or narrow_panels" \
  > gpurun_out/sanitizer_r02_mem.log 2>&1
echo "memcheck exit $?" >> gpurun_out/sanitizer_r02_mem.log
grep -E "ERROR SUMMARY|passed|failed|memcheck exit" gpurun_out/sanitizer_r02_mem.log | tail -5
