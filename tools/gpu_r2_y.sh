#!/bin/bash
# checkpoint on the final SDDMM sources: smoke, full GPU suite, bench, SDDMM tables, fresh ncu capture of the arxiv@256 kernel
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
timeout 300 python tools/exp_sddmm_k.py > gpurun_out/sd_arxiv_final.txt 2>&1; tail -4 gpurun_out/sd_arxiv_final.txt
timeout 300 python tools/bench_vs_ref.py --small --reps 50 > gpurun_out/small_final.jsonl 2>/dev/null
grep sddmm gpurun_out/small_final.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['graph'][:10], d['K'], round(d['ours_ms']*1e3,1), round(d['reference_cuda_ms']*1e3,1), round(d['speedup'],2))"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_y.log 2>&1; tail -3 gpurun_out/pytest_y.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_y.log 2> gpurun_out/bench_y.err; echo "bench rc=$?"
CMD="python bench.py --workload arxiv256 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-ref-cuda --no-secondary --no-legs"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_arxiv256.csv $CMD > gpurun_out/ncu_list_arxiv256.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sddmm_ring -s 3 -c 1 -o gpurun_out/r02_prof_arxiv256 -f $CMD > gpurun_out/ncu_full_arxiv256.log 2>&1
tail -2 gpurun_out/ncu_full_arxiv256.log
