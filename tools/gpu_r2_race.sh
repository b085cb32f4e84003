#!/bin/bash
# round 2: compute-sanitizer racecheck (shared-memory hazards) over the kernels with new shared-memory traffic this round:
# the tiled exact-fp32 spconv kernel (transposed A tile + W chunk) and the row-parallel / row-segment SpMM staging
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 10 --launch-timeout 0 \
  python -m pytest tests/test_spconv_gpu.py tests/test_spmm_gpu.py -x -q -m gpu -p no:cacheprovider \
  -k "(forward_random_maps and fp32) or forward_matches_c_oracle_small or row_parallel_kernel_forced or test_edge_cases" \
  > gpurun_out/sanitizer_r02_race.log 2>&1
echo "racecheck exit $?" >> gpurun_out/sanitizer_r02_race.log
grep -E "RACECHECK SUMMARY|hazard|passed|failed|racecheck exit" gpurun_out/sanitizer_r02_race.log | tail -6
