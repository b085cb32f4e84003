// kmap.h — internal C++ interface of the kernel-map builder (kmap.cu).  Torch-free.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace dgs {

size_t kmap_workspace_bytes(int in_nnz, int out_nnz, int k_vol);
// out_coords[<= in_nnz, 4] = sorted unique of (batch, floor(x / sx), floor(y / sy), floor(z / sz)); *out_nnz_dev = count
cudaError_t kmap_downsample(int in_nnz, const int *in_coords, int sx, int sy, int sz, int *out_coords, int *out_nnz_dev,
                            void *workspace, size_t workspace_bytes, cudaStream_t stream);
// pair lists grouped by kernel offset (capacity k_vol * out_nnz each), knnz[k_vol], kpos[k_vol + 1], qkpos[k_vol + 1]
cudaError_t kmap_build(int in_nnz, const int *in_coords, int out_nnz, const int *out_coords, int ksx, int ksy, int ksz, int sx,
                       int sy, int sz, int q, int skip_mid, int *imap, int *omap, int *knnz, int *kpos, int *qkpos,
                       void *workspace, size_t workspace_bytes, cudaStream_t stream);
// general layer (any stride, padding, output bounds): out_coords[<= in_nnz * k_vol, 4] = sorted unique of the voxels
// (in - tap_offset + padding) / stride that divide exactly and lie in [lo, hi] (3 host ints each)
size_t kmap_expand_workspace_bytes(int in_nnz, int k_vol);
cudaError_t kmap_downsample_expand(int in_nnz, const int *in_coords, int ksx, int ksy, int ksz, int sx, int sy, int sz, int px,
                                   int py, int pz, const int *lo, const int *hi, int *out_coords, int *out_nnz_dev,
                                   void *workspace, size_t workspace_bytes, cudaStream_t stream);
// kmap_build with padding; subm != 0: centred taps around out (stride ignored), else input = out * stride - padding + tap offset
cudaError_t kmap_build_ex(int in_nnz, const int *in_coords, int out_nnz, const int *out_coords, int ksx, int ksy, int ksz, int sx,
                          int sy, int sz, int px, int py, int pz, int subm, int q, int skip_mid, int *imap, int *omap, int *knnz,
                          int *kpos, int *qkpos, void *workspace, size_t workspace_bytes, cudaStream_t stream);

}  // namespace dgs
