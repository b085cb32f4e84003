#!/bin/bash
# round 2, GPU call M: does asking for 3 CTAs/SM help the arg-tracking SpMM flavours?  (before: reddit N=128 max+arg 5.56 ms, products 10.9 ms)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
timeout 600 python -m pytest tests/test_spmm_gpu.py tests/test_torch_face_gpu.py -x -q -m gpu > gpurun_out/pytest_m.log 2>&1; tail -2 gpurun_out/pytest_m.log
timeout 300 python tools/exp_panels.py reddit 128 2>/dev/null | head -4 | cut -c1-170
timeout 300 python tools/exp_panels.py reddit 64 2>/dev/null | head -4 | cut -c1-170
timeout 400 python tools/exp_panels.py products 128 2>/dev/null | head -4 | cut -c1-170
