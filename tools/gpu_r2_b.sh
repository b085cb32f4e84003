#!/bin/bash
# round 2, GPU call B: full GPU suite, latency regime vs the reference's kernels, L2 capacity curve, op-call latency
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_b.log 2>&1
tail -6 gpurun_out/pytest_b.log
timeout 300 python tools/bench_vs_ref.py --small --reps 50 > gpurun_out/small_default.jsonl 2> gpurun_out/small_default.err
DGS_SPMM_ROWPAR=0 DGS_SDDMM_NO_RING=1 timeout 300 python tools/bench_vs_ref.py --small --reps 50 > gpurun_out/small_norowpar_noring.jsonl 2> gpurun_out/small_norowpar_noring.err
timeout 300 python tools/bench_call_latency.py > gpurun_out/call_latency.jsonl 2> gpurun_out/call_latency.err
timeout 600 python tools/exp_l2_capacity.py > gpurun_out/l2_capacity.jsonl 2> gpurun_out/l2_capacity.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct \
    --clock-control none -k regex:spmm_rowseg --csv --log-file gpurun_out/ncu_l2_capacity.csv \
    python tools/exp_l2_capacity.py > gpurun_out/ncu_l2_capacity.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_b.log 2> gpurun_out/bench_b.err
tail -c 600 gpurun_out/bench_b.log
cat gpurun_out/small_default.jsonl | cut -c1-250
cat gpurun_out/call_latency.jsonl
cat gpurun_out/l2_capacity.jsonl
