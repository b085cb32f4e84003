// spconv.h — internal C++ interface between the C ABI (cabi.cu) and the spconv kernels.  Torch-free.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace dgs {

enum SpconvPrecision { SPCONV_FP32 = 0, SPCONV_TF32 = 1, SPCONV_BF16 = 2, SPCONV_FP16 = 3 };

// One gather-GEMM-scatter pass:  out[omap[p], :ndim] += in[imap[p], :kdim] @ Wk   (Wk = kdim x ndim view of W[k])
// The forward uses (kdim, ndim) = (c_in, c_out) with W[k][c][n]; the dX backward swaps the maps and uses
// (kdim, ndim) = (c_out, c_in) with W[k][n][c]: the strides w_sc / w_sn say which.
struct SpconvProblem {
  int k_vol = 0, kdim = 0, ndim = 0;
  const int *kpos = nullptr;    // [k_vol + 1] device: pair ranges per kernel offset
  const int *qkpos = nullptr;   // [k_vol + 1] device: the same ranges rounded up to multiples of 128 (tile table)
  const int *imap = nullptr;    // [kpos[k_vol]] gather rows
  const int *omap = nullptr;    // [kpos[k_vol]] scatter rows
  int64_t sum_nnz = 0;          // qkpos[k_vol]
  const float *in = nullptr;
  int64_t ld_in = 0;
  int in_rows = 0;
  const float *W = nullptr;
  int64_t w_sc = 0, w_sn = 0, w_sk = 0;   // element strides of W along the GEMM K index, N index and offset k
  float *out = nullptr;
  int64_t ld_out = 0;
  int out_rows = 0;
  int precision = SPCONV_TF32;
  int separate_mid = 0;         // also apply the centre offset to the identity map (in_rows == out_rows)
  int accumulate = 0;           // 0: out is zeroed first
};

size_t spconv_workspace_bytes(int rows, int k_vol, int c_in, int c_out, int precision);   // rows = max(in_nnz, out_nnz)
cudaError_t spconv_gemm(const SpconvProblem &p, void *workspace, size_t workspace_bytes, cudaStream_t stream);

// dW[k] = sum_p in[imap[p], :]^T (x) dout[omap[p], :]   (kernel gradient), fp32 accumulate
cudaError_t spconv_wgrad(int k_vol, int c_in, int c_out, const int *kpos, const int *qkpos, const int *imap, const int *omap,
                         int64_t sum_nnz, const float *in, int64_t ld_in, int in_rows, const float *dout, int64_t ld_dout,
                         float *dW, int precision, int separate_mid, cudaStream_t stream);

}  // namespace dgs
