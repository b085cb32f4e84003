#!/bin/bash
# round 2, GPU call W (1 GPU): csr2csc rewrite — parity, timing against the reference op, launch list, sanitizer (memcheck + racecheck)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
timeout 900 python -m pytest tests/test_sddmm_csr2csc_gpu.py tests/test_torch_face_gpu.py tests/test_compiled_ops_gpu.py tests/test_vs_reference_torch_face_gpu.py -x -q -m gpu 2>&1 | tail -2
timeout 900 python tools/bench_csr2csc.py --reps 10 2> gpurun_out/csr2csc.err | tee gpurun_out/csr2csc_w.jsonl | cut -c1-330
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_csr2csc_w.csv python tools/bench_csr2csc.py --reps 1 --impl ours --graph reddit-like > /dev/null 2>&1
python tools/launch_table.py gpurun_out/launches_csr2csc_w.csv 2>/dev/null | tail -14
for TOOL in memcheck racecheck; do
timeout 1200 compute-sanitizer --tool $TOOL --error-exitcode 9 --launch-timeout 0 python -m pytest tests/test_sddmm_csr2csc_gpu.py -x -q -m gpu -k "csr2csc_shapes or csr2csc_bit_exact" > gpurun_out/sanitizer_r02_csr2csc_$TOOL.log 2>&1
echo "$TOOL exit $?" >> gpurun_out/sanitizer_r02_csr2csc_$TOOL.log
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit" gpurun_out/sanitizer_r02_csr2csc_$TOOL.log | tail -4
done
