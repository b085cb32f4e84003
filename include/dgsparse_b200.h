/*
 * dgsparse_b200.h — extended C ABI of libdgsparse_b200.so (stream-taking, error-returning).
 *
 * This is what the PyTorch face (dgsparse-lib_b200/dgsparse) binds through ctypes and what a C/C++
 * caller should prefer over the legacy symbols of dgsparse.h.  Plain pointers and sizes only.
 * Every function returns 0 on success or a cudaError_t value (dgs_last_error() gives the text).
 * `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  All data pointers are
 * device pointers unless the function name ends in _host.
 */
#ifndef DGSPARSE_B200_H
#define DGSPARSE_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* REDUCEOP / COMPUTEOP of the reference, include/gspmm.h:13-14 (order is part of the gspmm-fp pybind
 * surface, src/gspmm-fp/gspmm.cc:31-42). */
enum dgsReduce { DGS_SUM = 0, DGS_MAX = 1, DGS_MIN = 2, DGS_MEAN = 3 };
enum dgsCompute { DGS_ADD = 0, DGS_SUB = 1, DGS_MUL = 2, DGS_DIV = 3, DGS_COPY = 4, DGS_MASKMUL = 5 };

int dgs_version(void);                 /* library version, major*10000 + minor*100 + patch */
int dgs_cuda_version(void);            /* CUDA_VERSION the library was built with: src/version.cpp:11-21 */
const char *dgs_last_error(void);      /* text of the last failure on this thread ("" if none) */
int dgs_sm_count(void);

/* Scratch the SpMM needs for rows cut by a segment boundary (see csrc/spmm_rowseg.cuh). */
size_t dgs_spmm_workspace_bytes(int N, int64_t nnz, int with_arg);

/* Generalized CSR SpMM:  C[r, :] = REDUCE_{p in row r} COMPUTE(val[p], B[col[p], :]).
 * Replaces spmm_cuda(Tensor...) src/cuda/spmm_cuda.cu:14-253 (algorithm 0 semantics,
 * include/cuda/spmm_cuda.cuh:10-55) and GSpMM_cuda / GSpMM_no_value_cuda src/gspmm-fp/gspmm.cu:442-473.
 *   val NULL  -> no edge value (compute is ignored, COPY)
 *   E non-NULL (MAX/MIN only) -> E[r, c] = column index of the winning nonzero, -1 for empty rows
 *   empty row -> 0;  MEAN divides by the row's nnz count
 *   n_dst > 1 -> the finished C rows are also stored to dst[1..] (NVLink peer pointers): the fused
 *                column-shard epilogue.  dst[0] is the local C.  All share ldc. */
int dgs_spmm_csr(int M, int N, int64_t nnz, const int *rowptr, const int *col, const float *val, const float *B,
                 int64_t ldb, float *C, int64_t ldc, int *E, int64_t lde, int reduce, int compute, void *workspace,
                 size_t workspace_bytes, void *stream);
/* Experiment / test knobs (csrc/options.h lists them; each is also read once from the environment, DGS_<NAME>):
 * dgs_set_option("spmm_rowpar", 1) etc.; value < 0 clears the override.  Returns 0, or -1 for an unknown name. */
int dgs_set_option(const char *name, int value);

/* Which kernel family the calling thread's last SpMM launched: 0 = row-segment kernel + fix-up (two launches, any matrix),
 * 1 = row-parallel single launch (latency regime, matrices the library has seen to have short rows only). */
int dgs_spmm_last_path(void);
/* Forget what the library has learnt about the matrices it has seen (which ones may take the row-parallel kernel):
 * the next call on any matrix starts from the row-segment path again.  For tests and benchmarks. */
void dgs_spmm_forget_graph_notes(void);
/* The legacy entry points of dgsparse.h take no workspace argument: the library keeps ONE grow-only device scratch per device for
 * them (segment partials; for column-major gespmmCsrSpMM also the transposed copies of B and C), used on stream 0 only.  This
 * waits for stream 0 of the current device and frees that scratch; the next legacy call allocates what it needs again.
 * Returns 0 or a cudaError_t value. */
int dgs_legacy_scratch_release(void);
/* Geometry of the calling thread's last SDDMM launch, for tools and tests: warps per CTA and CTAs per SM of the shared-memory
 * ring kernel (both 0 when the register-staged kernel ran) and the edges per warp / lane group.  Null pointers are skipped. */
void dgs_sddmm_last_geometry(int *warps_per_cta, int *ctas_per_sm, int *edges_per_warp);

/* The same with K = rows of B (columns of A) stated.  K only sizes the column panels (csrc/spmm.cu pick_panel): when a
 * 64-column panel of B, K x 256 B, would not stay L2-resident the feature axis is processed in narrower panels, one after
 * the other.  dgs_spmm_csr takes K = M (square adjacency). */
int dgs_spmm_csr_k(int M, int K, int N, int64_t nnz, const int *rowptr, const int *col, const float *val, const float *B,
                   int64_t ldb, float *C, int64_t ldc, int *E, int64_t lde, int reduce, int compute, void *workspace,
                   size_t workspace_bytes, void *stream);
int dgs_spmm_csr_multi(int M, int N, int64_t nnz, const int *rowptr, const int *col, const float *val, const float *B,
                       int64_t ldb, int n_dst, float *const *dst, int64_t ldc, int reduce, int compute, void *workspace,
                       size_t workspace_bytes, void *stream);

/* The same epilogue through an NVLS MULTICAST address (NVSwitch, sm_90+): mc_dst is the multicast mapping of a buffer
 * bound on every rank of the box (e.g. the multicast_ptr of a torch symmetric-memory rendezvous) offset to this rank's
 * column panel; each finished row is sent ONCE with multimem.st and the switch replicates it into every rank's C,
 * this rank's included.  Same completion rule as dgs_spmm_csr_multi: a stream-ordered barrier across the ranks. */
int dgs_spmm_csr_mcast(int M, int N, int64_t nnz, const int *rowptr, const int *col, const float *val, const float *B,
                       int64_t ldb, float *mc_dst, int64_t ldc, int reduce, int compute, void *workspace,
                       size_t workspace_bytes, void *stream);

/* End-of-step barrier of the multicast exchange: one multimem.red (release) adds 1 to every rank's copy of a 4-byte arrival
 * counter in symmetric memory (mc_counter = its multicast address), then the kernel spins (acquire) on this rank's own copy
 * (local_counter) until it reaches `target` = epoch * world.  Stream-ordered behind the SpMM of the step; when it returns on
 * a stream, every rank's multicast stores of the step are visible to later work on that stream.  The counter must be zeroed
 * on all ranks before the first epoch. */
int dgs_mcast_barrier(void *mc_counter, const void *local_counter, unsigned target, void *stream);

/* Masked SpMM of the max/min backward (grad wrt dense), called on the CSC arrays:
 *   out[j, v] = sum_{p in ptr[j]..ptr[j+1]} [E[idx[p], v] == j] * val[p] * G[idx[p], v]
 * Replaces spmm_cuda_with_mask src/cuda/spmm_cuda.cu:255-303 (intended semantics of
 * include/cuda/spmm_cuda.cuh:400-433). */
int dgs_spmm_csr_mask(int M, int N, int64_t nnz, const int *ptr, const int *idx, const float *val, const float *G,
                      int64_t ldg, const int *E, int64_t lde, float *out, int64_t ldo, void *workspace,
                      size_t workspace_bytes, void *stream);

/* SDDMM: out[e] = dot(D1[row(e), :K], D2[col(e), :K]).  mean != 0 divides by the row degree
 * (sddmmCSR{1,2}Scale<MEAN> include/cuda/sddmm_cuda.cuh:222-401); E non-NULL restricts the dot to the
 * feature positions with E[row(e), c] == col(e) (sddmmCSR1Scale_with_mask :403-507).
 * Replaces sddmm_cuda_csr / sddmm_cuda_coo src/cuda/spmm_cuda.cu:305-382 and src/sddmm/sddmm.cu:8-41. */
int dgs_sddmm_csr(int M, int K, int64_t nnz, const int *rowptr, const int *col, const float *D1, int64_t ld1,
                  const float *D2, int64_t ld2, const int *E, int mean, float *out, void *stream);
int dgs_sddmm_coo(int K, int64_t nnz, const int *row, const int *col, const float *D1, int64_t ld1, const float *D2,
                  int64_t ld2, float *out, void *stream);

/* Exact stable CSR -> CSC.  Replaces csr2csc_cuda src/cuda/spmm_cuda.cu:384-414 (cuSPARSE) and the
 * float permutation of dgsparse/storage.py:159-174.  val/val_t/row/perm may be NULL. */
size_t dgs_csr2csc_workspace_bytes(int M, int ncols, int64_t nnz);
int dgs_csr2csc(int M, int ncols, int64_t nnz, const int *rowptr, const int *col, const float *val, int *colptr,
                int *row, float *val_t, int *perm, void *workspace, size_t workspace_bytes, void *stream);

/* Per row and head: softmax over the row's nonzeros of values[p*head + h] (see dgsparse.h). */
int dgs_edge_softmax(int M, int head, const int *rowptr, const float *values, float *out, void *stream);

/* Sparse 3-D convolution, fused gather-GEMM-scatter (the reference's unbuilt torch.ops.dgsparse_spconv.spconv,
 * src/spconv.cpp:18-74; spconv_fwd_fused / spconv_bwd_fused, src/cuda/spconv_cuda.cu:18-253):
 *   forward   out_feats[out_map[p], :] += in_feats[in_map[p], :] @ kernel[k]     p in [kpos[k], kpos[k+1])
 *   backward  in_grad[in_map[p], :]    += out_grad[out_map[p], :] @ kernel[k]^T
 *             kernel_grad[k]           += in_feats[in_map[p], :]^T (x) out_grad[out_map[p], :]
 * in_feats [in_nnz, c_in], kernel [k_vol, c_in, c_out], out_feats [out_nnz, c_out], all fp32 row-major contiguous.
 * kpos / qkpos: device int32 [k_vol + 1]; qkpos = kpos with every offset's count rounded up to a multiple of 128
 * (test/test_spconv.py:5-14), sum_nnz = qkpos[k_vol].  separate_mid != 0: the centre offset (k_vol / 2) is NOT in
 * the maps and is applied to the identity map (needs in_nnz == out_nnz), spconv_cuda.cu:52-78.
 * precision: DGS_SPCONV_FP32 = exact fp32 FMA (the reference's arch80 = false kernels), DGS_SPCONV_TF32 = tcgen05
 * kind::tf32 with fp32 accumulation in TMEM (arch80 = true), DGS_SPCONV_BF16 = tcgen05 kind::f16 on bf16-rounded
 * operands.  Outputs are fully overwritten (the reference leaves out_feats uninitialised).
 * in_grad or kernel_grad may be NULL to skip that half of the backward.  kernel_grad is always fp32 FMA. */
enum dgsSpconvPrecision { DGS_SPCONV_FP32 = 0, DGS_SPCONV_TF32 = 1, DGS_SPCONV_BF16 = 2, DGS_SPCONV_FP16 = 3 };
size_t dgs_spconv_workspace_bytes(int rows /* max(in_nnz, out_nnz) */, int k_vol, int c_in, int c_out, int precision);
int dgs_spconv_fwd(int in_nnz, int out_nnz, int k_vol, int c_in, int c_out, const int *kpos, const int *qkpos,
                   const int *in_map, const int *out_map, int64_t sum_nnz, const float *in_feats, const float *kernel,
                   float *out_feats, int separate_mid, int precision, void *workspace, size_t workspace_bytes,
                   void *stream);
int dgs_spconv_bwd(int in_nnz, int out_nnz, int k_vol, int c_in, int c_out, const int *kpos, const int *qkpos,
                   const int *in_map, const int *out_map, int64_t sum_nnz, const float *out_grad, const float *in_feats,
                   const float *kernel, float *in_grad, float *kernel_grad, int separate_mid, int precision,
                   void *workspace, size_t workspace_bytes, void *stream);

/* Kernel-map construction for sparse convolution (the reference's unregistered sparse_mapping,
 * src/cuda/sparse_mapping.cu:20-161): from integer coordinates (batch, x, y, z) int32 [n, 4] to the pair lists the
 * spconv op consumes.  |x|, |y|, |z| < 32768, 0 <= batch < 65536.
 *   dgs_kmap_downsample: out_coords = sorted unique of (batch, floor(x/sx), floor(y/sy), floor(z/sz)), capacity in_nnz
 *                        rows; the count is written to the DEVICE int *out_nnz_dev (coordsDownsample + sort + unique).
 *   dgs_kmap_build:      for every kernel tap k = (kx*ksy + ky)*ksz + kz and output o, the input at
 *                        out + (tap - (ks-1)/2) (all strides 1, _queryhash_subm) or out*stride + tap (_queryhash_sp,
 *                        padding 0) if it exists.  imap / omap (capacity ksx*ksy*ksz*out_nnz each) are grouped by k and
 *                        ordered by o inside a group (deterministic); knnz[k_vol], kpos[k_vol+1] and qkpos[k_vol+1]
 *                        (counts rounded up to q, 128 for dgs_spconv_*) are device arrays; kpos[k_vol] = pair count.
 *                        skip_mid != 0 leaves the centre tap out (for spconv's separate_mid). */
size_t dgs_kmap_workspace_bytes(int in_nnz, int out_nnz, int k_vol);
int dgs_kmap_downsample(int in_nnz, const int *in_coords, int sx, int sy, int sz, int *out_coords, int *out_nnz_dev,
                        void *workspace, size_t workspace_bytes, void *stream);
int dgs_kmap_build(int in_nnz, const int *in_coords, int out_nnz, const int *out_coords, int ksx, int ksy, int ksz,
                   int sx, int sy, int sz, int q, int skip_mid, int *imap, int *omap, int *knnz, int *kpos, int *qkpos,
                   void *workspace, size_t workspace_bytes, void *stream);

/* General layers (stride neither 1 nor the kernel size, padding, output bounds): the coordsDownsampleExpand branch of
 * the reference (src/cuda/sparse_mapping.cu:98-137, include/cuda/sparse_mapping.cuh:326-401).
 *   dgs_kmap_downsample_expand: out_coords = sorted unique of the voxels (in - off(tap) + padding) / stride over all
 *                        (input, tap) whose division is exact and whose result lies in [lo, hi] (3 HOST ints each,
 *                        output resolution); off(k) = k - (ks-1)/2 (+1 for odd ks > 1), the reference's tap decode.
 *                        Capacity in_nnz * k_vol rows; count to the DEVICE int *out_nnz_dev.  Padding is per axis (the
 *                        reference reads padding[0] for all three axes in this kernel and padding[0..2] in the query).
 *   dgs_kmap_build_ex:   dgs_kmap_build with padding: subm != 0 -> input = out + (tap - (ks-1)/2); otherwise
 *                        input = out * stride - padding + off(tap) (_queryhash_sp, :141-220).  dgs_kmap_build is
 *                        dgs_kmap_build_ex with padding 0 and subm = (all strides 1). */
size_t dgs_kmap_expand_workspace_bytes(int in_nnz, int k_vol);
int dgs_kmap_downsample_expand(int in_nnz, const int *in_coords, int ksx, int ksy, int ksz, int sx, int sy, int sz, int px,
                               int py, int pz, const int *lo, const int *hi, int *out_coords, int *out_nnz_dev,
                               void *workspace, size_t workspace_bytes, void *stream);
int dgs_kmap_build_ex(int in_nnz, const int *in_coords, int out_nnz, const int *out_coords, int ksx, int ksy, int ksz,
                      int sx, int sy, int sz, int px, int py, int pz, int subm, int q, int skip_mid, int *imap, int *omap,
                      int *knnz, int *kpos, int *qkpos, void *workspace, size_t workspace_bytes, void *stream);

/* Peer memory for the fused column-shard epilogue: export a device allocation to the other ranks of
 * the box (CUDA IPC), open theirs; the opened base + offset is passed in dst[] of dgs_spmm_csr_multi. */
int dgs_ipc_export(const void *dptr, void *handle64, int64_t *offset);
int dgs_ipc_open(const void *handle64, void **base);
int dgs_ipc_close(void *base);

/* Per-launch device timing (bench.py's roofline leg): when enabled every kernel this library launches
 * is bracketed by CUDA events on its stream.  dgs_profile_collect synchronises those events and returns
 * the number of records written: kernel_ids[i] (1 = SpMM row-segment, 2 = SpMM fix-up, 3 = SDDMM,
 * 4 = csr2csc passes, 5 = spconv gather-GEMM-scatter, 6 = spconv kernel gradient) and ms[i]; it then clears the list. */
int dgs_profile_enable(int on);
int dgs_profile_collect(int max_records, int *kernel_ids, float *ms);

/* HOST-buffer entry point (what a CPU-side caller of the reference's host functions would switch
 * to): copies the CSR and B to the device, runs dgs_spmm_csr, copies C (and E) back, synchronises.
 * All pointers are host pointers; pinned memory makes the copies asynchronous.  Device staging is
 * cached per thread between calls. */
int dgs_spmm_csr_host(int M, int K, int N, int64_t nnz, const int *rowptr, const int *col, const float *val,
                      const float *B, float *C, int *E, int reduce, int compute);
/* The same for a matrix that stays the same across calls (a GNN's adjacency): dgs_csr_upload copies the host CSR to the
 * device ONCE and returns a handle; dgs_spmm_csr_resident_host then moves only B (host [K, N]) in and C (host [M, N], and E)
 * out per call, the row blocks of A multiplied while the finished rows of C are already crossing PCIe.  Synchronises before
 * returning.  One handle is used by one thread at a time. */
int dgs_csr_upload(int M, int K, int64_t nnz, const int *rowptr, const int *col, const float *val /* may be NULL */, void **handle);
int dgs_spmm_csr_resident_host(void *handle, int N, const float *B, float *C, int *E /* may be NULL */, int reduce, int compute);
int dgs_csr_free(void *handle);
int dgs_sddmm_csr_host(int M, int Kdim, int ncols, int64_t nnz, const int *rowptr, const int *col, const float *D1,
                       const float *D2, float *out);

#ifdef __cplusplus
}
#endif
#endif /* DGSPARSE_B200_H */
