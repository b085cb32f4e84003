#!/bin/bash
# round 2, GPU call E (1 GPU): full suite, SM-affine locality experiment, spconv vs the reference's own CUDA
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_e.log 2>&1
tail -6 gpurun_out/pytest_e.log
timeout 600 python tools/exp_locality.py 64 > gpurun_out/locality64_affine.jsonl 2> gpurun_out/locality64_affine.err
DGS_SPMM_SM_AFFINE=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,lts__throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:spmm_rowseg --csv --log-file gpurun_out/ncu_locality64_affine.csv \
    python tools/exp_locality.py 64 > gpurun_out/ncu_locality64_affine.log 2>&1
timeout 600 python tools/bench_spconv.py --reps 20 --channels "128,128" > gpurun_out/spconv_vs_ref.jsonl 2> gpurun_out/spconv_vs_ref.err
timeout 300 python tools/bench_vs_ref.py --small --reps 50 > gpurun_out/small_default.jsonl 2> gpurun_out/small_default.err
cat gpurun_out/locality64_affine.jsonl | cut -c1-260
cat gpurun_out/spconv_vs_ref.jsonl | cut -c1-420
tail -3 gpurun_out/spconv_vs_ref.err
