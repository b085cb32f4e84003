#!/bin/bash
# compute-sanitizer over the kernels added late in round 1: the rebuilt csr2csc (memcheck + racecheck: shared-memory tile
# sort, ballot ranking, per-warp atomics), the kernel-map expand branch, the batched fix-up fold.
mkdir -p gpurun_out
timeout 300 python -m pytest -m gpu -q -p no:cacheprovider tests/test_sddmm_csr2csc_gpu.py -k "csr2csc" 2>&1 | tail -2
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest -m gpu -q -x -p no:cacheprovider \
    "tests/test_sddmm_csr2csc_gpu.py::test_csr2csc_shapes" "tests/test_kmap_gpu.py::test_expand_branch_bit_exact" \
    "tests/test_kmap_gpu.py::test_expand_empty_input" "tests/test_spmm_gpu.py::test_widths_and_reduces" \
    > gpurun_out/sanitizer2_mem.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/sanitizer2_mem.log
grep -E "ERROR SUMMARY|Invalid|rc=|passed|failed" gpurun_out/sanitizer2_mem.log | head -8
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 \
  python -m pytest -m gpu -q -x -p no:cacheprovider "tests/test_sddmm_csr2csc_gpu.py::test_csr2csc_shapes" \
    > gpurun_out/sanitizer2_race.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/sanitizer2_race.log
grep -E "RACECHECK SUMMARY|hazard|rc=|passed|failed" gpurun_out/sanitizer2_race.log | head -8
