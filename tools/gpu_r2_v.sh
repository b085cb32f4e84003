#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
timeout 600 python tools/exp_ab_variant.py t64 --products > gpurun_out/exp_ab_t64.jsonl 2> gpurun_out/exp_ab_t64.err
cat gpurun_out/exp_ab_t64.jsonl; tail -3 gpurun_out/exp_ab_t64.err
