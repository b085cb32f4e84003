#!/bin/bash
# Builds the reference's OWN gspmm-fp pybind module (src/gspmm-fp/gspmm.cu + gspmm.cc, UNMODIFIED, where they lie under
# /root/reference) for sm_100a into oracle/_ref/spmm.so — the module example/gspmm-fp/util.py:7-14 JIT-loads as `spmm`.
# TEST / BENCH INFRASTRUCTURE ONLY: tests/test_vs_reference_gspmm_gpu.py compares our generalized SpMM with it op by op
# and bench.py times it beside ours on the gspmm workload.  ~6.5 minutes (torch headers), so it is NOT part of the
# default oracle build; run it once in the build container, the .so travels to the GPU box with gpurun.
set -e
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
[ -f "$REF/src/gspmm-fp/gspmm.cu" ] || { echo "reference tree absent"; exit 0; }
T=$(python -c "import torch, os; print(os.path.dirname(torch.__file__))")
PYI=$(python -c "import sysconfig; print(sysconfig.get_paths()['include'])")
ABI=$(python -c "import torch; print(int(torch._C._GLIBCXX_USE_CXX11_ABI))")
W=$(mktemp -d)
DEFS="-DTORCH_EXTENSION_NAME=spmm -DTORCH_API_INCLUDE_EXTENSION_H -D_GLIBCXX_USE_CXX11_ABI=$ABI"
INC="-I$T/include -I$T/include/torch/csrc/api/include -I$PYI -I$REF/src/gspmm-fp"
nvcc -O2 -std=c++17 -w -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a $DEFS $INC -c "$REF/src/gspmm-fp/gspmm.cu" -o "$W/gspmm_cu.o"
g++ -O2 -std=c++17 -w -fPIC $DEFS $INC -I/usr/local/cuda/include -c "$REF/src/gspmm-fp/gspmm.cc" -o "$W/gspmm_cc.o"
mkdir -p "$HERE/_ref"
g++ -shared -o "$HERE/_ref/spmm.so" "$W/gspmm_cu.o" "$W/gspmm_cc.o" -L"$T/lib" -lc10 -lc10_cuda -ltorch_cpu -ltorch_cuda -ltorch \
    -ltorch_python -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,"$T/lib"
rm -rf "$W"
echo "built $HERE/_ref/spmm.so"
