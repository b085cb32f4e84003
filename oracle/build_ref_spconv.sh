#!/bin/bash
# Builds the reference's OWN sparse-convolution CUDA (src/cuda/sparse_mapping.cu: kernel-map construction,
# src/cuda/spconv_cuda.cu: fused gather-GEMM-scatter forward / backward; UNMODIFIED, compiled where they lie under
# /root/reference) for sm_100a into oracle/_ref/_ref_spconv.so behind a no-algorithm pybind shim (ref_spconv_shim.cpp).
# TEST INFRASTRUCTURE ONLY: pins csrc/kmap.cu and csrc/spconv.cu against the reference (tests/test_vs_reference_spconv_gpu.py)
# and gives tools/bench_spconv.py a timed baseline.  Several minutes (torch headers + every spconv template), so it is NOT
# part of the default oracle build; run it once in the build container, the .so travels to the GPU box with gpurun.
set -e
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
[ -f "$REF/src/cuda/spconv_cuda.cu" ] || { echo "reference tree absent"; exit 0; }
T=$(python -c "import torch, os; print(os.path.dirname(torch.__file__))")
PYI=$(python -c "import sysconfig; print(sysconfig.get_paths()['include'])")
ABI=$(python -c "import torch; print(int(torch._C._GLIBCXX_USE_CXX11_ABI))")
W=$(mktemp -d)
DEFS="-DWITH_PYTHON -DWITH_CUDA -DTORCH_EXTENSION_NAME=_ref_spconv -DTORCH_API_INCLUDE_EXTENSION_H -D_GLIBCXX_USE_CXX11_ABI=$ABI"
INC="-I$T/include -I$T/include/torch/csrc/api/include -I$PYI"
NV="nvcc -O2 -std=c++17 -w -Xcompiler -fPIC --expt-relaxed-constexpr --extended-lambda -gencode arch=compute_100a,code=sm_100a $DEFS $INC"
$NV -c "$REF/src/cuda/sparse_mapping.cu" -o "$W/sparse_mapping.o" &
$NV -c "$REF/src/cuda/spconv_cuda.cu" -o "$W/spconv_cuda.o" &
g++ -O2 -std=c++17 -w -fPIC $DEFS $INC -I"$REF" -I/usr/local/cuda/include -c "$HERE/ref_spconv_shim.cpp" -o "$W/shim.o" &
wait
mkdir -p "$HERE/_ref"
g++ -shared -o "$HERE/_ref/_ref_spconv.so" "$W/sparse_mapping.o" "$W/spconv_cuda.o" "$W/shim.o" -L"$T/lib" -lc10 -lc10_cuda \
    -ltorch_cpu -ltorch_cuda -ltorch -ltorch_python -L/usr/local/cuda/lib64 -lcudart -lcublas -Wl,-rpath,"$T/lib"
rm -rf "$W"
echo "built $HERE/_ref/_ref_spconv.so"
