// ref_spconv_shim.cpp — TEST INFRASTRUCTURE ONLY.  A pybind11 binding with NO algorithm in it: it exposes the reference's
// OWN sparse-convolution entry points, compiled unmodified from /root/reference/src/cuda/{sparse_mapping,spconv_cuda}.cu
// where they lie (oracle/build_ref_spconv.sh), as the module oracle/_ref/_ref_spconv.so.  The reference itself never
// registers sparse_mapping as an op and never builds its spconv extension (setup.py:58-59), so this shim is the only way to
// run them; tests/test_vs_reference_spconv_gpu.py compares dgs_kmap_build / dgs_spconv_fwd with them on the same inputs and
// tools/bench_spconv.py times spconv_fwd_fused beside ours.  Never linked into, imported by or shipped with the product.
#include <torch/extension.h>
#include "include/cuda/sparse_mapping.h"   // at::Tensor sparse_mapping(...)             (src/cuda/sparse_mapping.cu:20-161)
#include "include/cuda/spconv_cuda.h"      // spconv_fwd_fused / spconv_bwd_fused        (src/cuda/spconv_cuda.cu:18-253)

PYBIND11_MODULE(_ref_spconv, m) {
  m.def("sparse_mapping", &sparse_mapping);
  m.def("spconv_fwd_fused", &spconv_fwd_fused);
  m.def("spconv_bwd_fused", &spconv_bwd_fused);
}
