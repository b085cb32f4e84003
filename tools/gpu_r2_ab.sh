#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python tools/exp_ab_r1.py 2> gpurun_out/ab.err | tee gpurun_out/ab_r1_vs_now.jsonl
tail -3 gpurun_out/ab.err
