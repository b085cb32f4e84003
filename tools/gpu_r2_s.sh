#!/bin/bash
# column-major gespmmCsrSpMM through the row-major kernel + SDDMM ring geometry sweep
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_spmm_gpu.py tests/test_vs_reference_cuda_gpu.py -m gpu -q -x -p no:cacheprovider -k "colmajor or older_api or descr" > gpurun_out/pytest_s.log 2>&1
tail -3 gpurun_out/pytest_s.log
timeout 600 python tools/exp_colmajor.py --reps 20 --scale 0.25 > gpurun_out/exp_colmajor.jsonl 2> gpurun_out/exp_colmajor.err
cat gpurun_out/exp_colmajor.jsonl; tail -3 gpurun_out/exp_colmajor.err
timeout 600 python tools/exp_sddmm_ring.py > gpurun_out/exp_sddmm_ring.jsonl 2> gpurun_out/exp_sddmm_ring.err
cat gpurun_out/exp_sddmm_ring.jsonl; tail -3 gpurun_out/exp_sddmm_ring.err
