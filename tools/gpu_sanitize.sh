#!/bin/bash
# compute-sanitizer (memcheck) over a representative subset of the GPU tests: small problems, every kernel family.
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest -m gpu -q -x -p no:cacheprovider \
    "tests/test_spconv_gpu.py::test_forward_random_maps" "tests/test_spconv_gpu.py::test_empty_offsets_and_ragged_tail" \
    "tests/test_spconv_gpu.py::test_torch_op_autograd" \
    "tests/test_sddmm_csr2csc_gpu.py::test_sddmm_widths" "tests/test_sddmm_csr2csc_gpu.py::test_masked_kernels" \
    "tests/test_kmap_gpu.py::test_kernel_map_bit_exact" "tests/test_kmap_gpu.py::test_edge_cases" \
    "tests/test_spmm_gpu.py::test_widths_and_reduces" "tests/test_spmm_gpu.py::test_host_buffer_entry" \
    > gpurun_out/sanitizer.log 2>&1
echo "sanitizer rc=$?" >> gpurun_out/sanitizer.log
grep -E "ERROR SUMMARY|Invalid|rc=|passed|failed" gpurun_out/sanitizer.log | head -20
