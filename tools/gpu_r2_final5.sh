#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
timeout 600 python -m pytest tests/test_spmm_gpu.py tests/test_vs_reference_cuda_gpu.py tests/test_cabi_symbols.py tests/test_reference_drivers_gpu.py -q -x -p no:cacheprovider -k "colmajor or older_api or descr or symbols or drivers" > gpurun_out/pytest_final5.log 2>&1
tail -3 gpurun_out/pytest_final5.log
