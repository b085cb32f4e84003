#!/bin/bash
# fresh launch lists + ncu --set full captures of the two SpMM headline kernels on the final sources
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
for SPEC in "reddit64 spmm_rowseg" "products128 spmm_rowseg"; do
  set -- $SPEC
  CMD="python bench.py --workload $1 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-ref-cuda --no-secondary --no-legs"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file gpurun_out/r02_launches_$1.csv $CMD > gpurun_out/ncu_list_$1.log 2>&1
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 \
      -o gpurun_out/r02_prof_$1 -f $CMD > gpurun_out/ncu_full_$1.log 2>&1
  tail -2 gpurun_out/ncu_full_$1.log
done
