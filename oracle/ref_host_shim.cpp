// ref_host_shim.cpp — exposes the REFERENCE's own host functions through a C ABI.
//
// TEST INFRASTRUCTURE ONLY (see oracle.c).  This file contains no algorithm: it #includes the
// reference header where it lies (-I/root/reference/example/util, never copied into this repo) and
// forwards to its templates, so oracle/_ref/libref_host.so *is* the reference's CPU code:
//   spmm_reference_host<int,float>   /root/reference/example/util/sp_util.hpp:62-84
//   sddmm_reference_host<int,float>  /root/reference/example/util/sp_util.hpp:87-112
// Built by oracle/Makefile into oracle/_ref/ (git-ignored, travels to the GPU box with gpurun).
#include <cstdint>
#include <cstring>
#include "sp_util.hpp"

extern "C" {

// Whole-matrix call, exactly as example/ge-spmm/spmm.cu:99 makes it.
void ref_spmm_host(int M, int N, int K, const int *rowptr, const int *col, const float *val,
                   const float *B, float *C) {
  spmm_reference_host<int, float>(M, N, K, rowptr, col, val, B, C);
}

// Row-block call used to spread the (single-threaded) reference loop over host cores:
// rows [r0, r1) of the same matrix; rowptr entries stay absolute offsets into col/val.
void ref_spmm_host_rows(int r0, int r1, int N, int K, const int *rowptr, const int *col,
                        const float *val, const float *B, float *C) {
  spmm_reference_host<int, float>(r1 - r0, N, K, rowptr + r0, col, val, B, C + (int64_t)r0 * N);
}

void ref_sddmm_host(int M, int N, int K, int nnz, const int *rowptr, const int *col,
                    const float *A, const float *B, float *C) {
  sddmm_reference_host<int, float>(M, N, K, nnz, rowptr, col, A, B, C);
}
}
