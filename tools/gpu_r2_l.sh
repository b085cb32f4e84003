#!/bin/bash
# round 2, GPU call L (1 GPU): column-slab SpMM — parity (forced small + products full size), then products@128 timings
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
timeout 900 python -m pytest tests/test_spmm_gpu.py -x -q -m gpu -k "slab or products or widths or edge_cases" > gpurun_out/pytest_l.log 2>&1
tail -5 gpurun_out/pytest_l.log
for ROWS in 0 160000 220000 280000 340000; do
  DGS_SPMM_SLAB_ROWS=$ROWS timeout 400 python bench.py --workload products128 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-ref-cuda --no-legs > gpurun_out/bench_products_slab_$ROWS.log 2>&1
  python - <<PY
import json
l=[x for x in open('gpurun_out/bench_products_slab_$ROWS.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('slab rows $ROWS', d['ms_per_step'], d['roofline']['kernel_ms_avg'], d['roofline']['fixup_ms_avg'], d['parity_ok'], d['gpu_launches'])
else:
    print('slab rows $ROWS FAILED', open('gpurun_out/bench_products_slab_$ROWS.log').read()[-1500:])
PY
done
DGS_SPMM_SLAB=0 timeout 400 python bench.py --workload products128 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-ref-cuda --no-legs 2>&1 | grep "^{" | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('no slab', d['ms_per_step'], d['roofline']['kernel_ms_avg'], d['parity_ok'])"
DGS_SPMM_SLAB_ROWS=220000 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:"spmm_rowseg|slab_|spmm_fixup" -c 40 --csv --log-file gpurun_out/ncu_slab_products.csv \
  python bench.py --workload products128 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-ref-cuda --no-legs > gpurun_out/ncu_slab_products.log 2>&1
python tools/launch_table.py gpurun_out/ncu_slab_products.csv 40 | tail -34
