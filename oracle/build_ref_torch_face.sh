#!/bin/bash
# Builds the reference's OWN torch extension dgsparse._spmm_cuda (src/spmm.cpp + src/cuda/spmm_cuda.cu, UNMODIFIED, where
# they lie under /root/reference; the CUDAExtension of setup.py:26-84) for sm_100a into oracle/_ref/_spmm_cuda.so.
# TEST INFRASTRUCTURE ONLY: oracle/run_ref_torch_face.py loads it in a process of its own and
# tests/test_vs_reference_torch_face_gpu.py compares torch.ops.dgsparse_spmm.* of the two implementations.
# ~10 minutes (spmm_cuda.cu instantiates every kernel under torch headers), so it is NOT part of the default oracle build;
# run it once in the build container, the .so travels to the GPU box with gpurun.
set -e
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
[ -f "$REF/src/cuda/spmm_cuda.cu" ] || { echo "reference tree absent"; exit 0; }
T=$(python -c "import torch, os; print(os.path.dirname(torch.__file__))")
PYI=$(python -c "import sysconfig; print(sysconfig.get_paths()['include'])")
ABI=$(python -c "import torch; print(int(torch._C._GLIBCXX_USE_CXX11_ABI))")
W=$(mktemp -d)
DEFS="-DWITH_PYTHON -DWITH_CUDA -DTORCH_EXTENSION_NAME=_spmm_cuda -DTORCH_API_INCLUDE_EXTENSION_H -D_GLIBCXX_USE_CXX11_ABI=$ABI"
INC="-I$T/include -I$T/include/torch/csrc/api/include -I$PYI"
nvcc -O2 -std=c++17 -w -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a $DEFS $INC -c "$REF/src/cuda/spmm_cuda.cu" -o "$W/spmm_cuda.o"
g++ -O2 -std=c++17 -w -fPIC $DEFS $INC -I/usr/local/cuda/include -c "$REF/src/spmm.cpp" -o "$W/spmm.o"
mkdir -p "$HERE/_ref"
g++ -shared -o "$HERE/_ref/_spmm_cuda.so" "$W/spmm_cuda.o" "$W/spmm.o" -L"$T/lib" -lc10 -lc10_cuda -ltorch_cpu -ltorch_cuda -ltorch \
    -ltorch_python -L/usr/local/cuda/lib64 -lcudart -lcusparse -Wl,-rpath,"$T/lib"
rm -rf "$W"
echo "built $HERE/_ref/_spmm_cuda.so"
