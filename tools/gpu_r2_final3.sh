#!/bin/bash
# final call of the round: memcheck + racecheck over this session's kernels (one-warp SDDMM rings with two D1 slots, the G-ary row
# search, 64-thread register SDDMM, the transposing column-major SpMM), smoke, full GPU suite, bench, arxiv@256 capture
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
SEL="two_d1_slots or tiny_and_hub or colmajor or sddmm_golden or masked_kernels"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 0 \
  python -m pytest tests/test_spmm_gpu.py tests/test_sddmm_csr2csc_gpu.py -x -q -m gpu -p no:cacheprovider -k "$SEL" > gpurun_out/sanitizer_r02_mem2.log 2>&1
echo "memcheck exit $?" >> gpurun_out/sanitizer_r02_mem2.log
grep -E "ERROR SUMMARY|passed|failed|memcheck exit" gpurun_out/sanitizer_r02_mem2.log | tail -4
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --launch-timeout 0 \
  python -m pytest tests/test_spmm_gpu.py tests/test_sddmm_csr2csc_gpu.py -x -q -m gpu -p no:cacheprovider -k "two_d1_slots and (256 or 100) or colmajor_both" > gpurun_out/sanitizer_r02_race2.log 2>&1
echo "racecheck exit $?" >> gpurun_out/sanitizer_r02_race2.log
grep -E "RACECHECK SUMMARY|passed|failed|racecheck exit" gpurun_out/sanitizer_r02_race2.log | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_final3.log 2>&1; tail -2 gpurun_out/pytest_final3.log
CMD="python bench.py --workload arxiv256 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-ref-cuda --no-secondary --no-legs"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_arxiv256.csv $CMD > gpurun_out/ncu_list_arxiv256.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sddmm_ring -s 3 -c 1 -o gpurun_out/r02_prof_arxiv256 -f $CMD > gpurun_out/ncu_full_arxiv256.log 2>&1
tail -1 gpurun_out/ncu_full_arxiv256.log
timeout 300 python tools/exp_sddmm_k.py 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final3.log 2> gpurun_out/bench_final3.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final3.log 2>&1; echo "reference arm rc=$?"
