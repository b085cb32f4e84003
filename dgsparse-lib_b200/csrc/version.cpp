// version.cpp — dgsparse/_C.so: the pybind module of src/version.cpp:11-21 (`cuda_version()`, checked by
// dgsparse/__init__.py:28-42 against torch.version.cuda).  The CUDA version is the one libdgsparse_b200.so was built with.
#include <pybind11/pybind11.h>

#include "../../include/dgsparse_b200.h"

static long long cuda_version() noexcept { return (long long)dgs_cuda_version(); }

PYBIND11_MODULE(_C, m) { m.def("cuda_version", &cuda_version, "cuda_version"); }
