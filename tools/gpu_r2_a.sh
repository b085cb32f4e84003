#!/bin/bash
# round 2, GPU call A: new parity tests + the panel-width experiment + ncu DRAM traffic per panel width
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu -k "narrow or tiny or spconv or kmap or products or sddmm or masked" > gpurun_out/pytest_a.log 2>&1
tail -5 gpurun_out/pytest_a.log
timeout 600 python tools/exp_panels.py products 128 > gpurun_out/exp_panels_products128.jsonl 2> gpurun_out/exp_panels_products128.err
timeout 300 python tools/exp_panels.py reddit 128 > gpurun_out/exp_panels_reddit128.jsonl 2> gpurun_out/exp_panels_reddit128.err
for P in 64 16 8; do
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct \
    --clock-control none -k regex:"spmm_(narrow|rowseg)" -c 2 --csv --log-file gpurun_out/ncu_panel_$P.csv \
    python tools/exp_panels.py products 128 --once $P > gpurun_out/ncu_panel_$P.log 2>&1
done
cat gpurun_out/exp_panels_products128.jsonl | cut -c1-220
