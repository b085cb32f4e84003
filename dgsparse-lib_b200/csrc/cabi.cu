// cabi.cu — the C ABI of libdgsparse_b200.so: include/dgsparse.h (legacy dgSPARSE symbols) and
// include/dgsparse_b200.h (extended, stream-taking).  No torch types anywhere.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <cuda.h>
#include "../../include/dgsparse.h"
#include "../../include/dgsparse_b200.h"
#include "common.cuh"
#include "spmm.h"
#include "options.h"
#include "spconv.h"
#include "kmap.h"

namespace {

thread_local char g_err[512] = "";

int fail(cudaError_t e, const char *where) {
  snprintf(g_err, sizeof(g_err), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
  return (int)e;
}
int ok_or(cudaError_t e, const char *where) {
  if (e == cudaSuccess) return 0;
  return fail(e, where);
}

// Grow-only device scratch for the legacy (workspace-less) entry points; one per device, used on
// stream 0 only, so reuse across calls is ordered by the stream.
struct Scratch {
  void *ptr = nullptr;
  size_t bytes = 0;
};
std::mutex g_mu;
Scratch g_scratch[64];

cudaError_t legacy_scratch(size_t need, void **out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  std::lock_guard<std::mutex> lk(g_mu);
  Scratch &s = g_scratch[dev];
  if (s.bytes < need) {
    if (s.ptr) {
      if ((e = cudaStreamSynchronize(0)) != cudaSuccess) return e;
      cudaFree(s.ptr);
      s.ptr = nullptr; s.bytes = 0;
    }
    size_t want = need + need / 4 + (1u << 20);
    if ((e = cudaMalloc(&s.ptr, want)) != cudaSuccess) return e;
    s.bytes = want;
  }
  *out = s.ptr;
  return cudaSuccess;
}

void legacy_report(int rc, const char *fn) {
  if (rc != 0) fprintf(stderr, "[dgsparse_b200] %s failed: %s\n", fn, g_err);
}

// nnz of a device CSR when the caller does not pass it (legacy spmm_cuda(m, k, rowptr, ...)).  The reference never needs
// it (src/ge-spmm/gespmm.cc:114-129 is fully asynchronous); our segment scheme does, to size the grid.  A blocking 4-byte
// copy of rowptr[m] on every call would serialise the caller's stream, so it happens ONCE per (device, rowptr, m): later
// calls size the grid from a hint — the nnz the kernels themselves found last time, reported through a mapped host word —
// and the kernels re-derive the segment layout from the true rowptr[m] on the device (seg_layout, spmm_rowseg.cuh), so
// a CSR rewritten in place under the same pointer still gives the right answer, only with a stale grid size for one call.
struct NnzHint {
  const int *rowptr = nullptr;
  int m = -1, dev = -1, nnz = -1;
  int *mail_host = nullptr, *mail_dev = nullptr;   // mapped pinned word, allocated once per slot
};
NnzHint g_hints[32];
int g_hint_clock = 0;

// -> hint nnz (> 0) and the device address of the report word; hint <= 0: the caller must not use the hinted path
cudaError_t legacy_nnz_hint(const int *rowptr, int m, int *nnz, int **report) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lk(g_mu);
  NnzHint *h = nullptr;
  for (auto &c : g_hints)
    if (c.rowptr == rowptr && c.m == m && c.dev == dev) { h = &c; break; }
  if (h == nullptr) {
    h = &g_hints[g_hint_clock++ % 32];
    if (h->mail_host == nullptr) {
      if ((e = cudaHostAlloc((void **)&h->mail_host, sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable)) != cudaSuccess) return e;
      if ((e = cudaHostGetDevicePointer((void **)&h->mail_dev, h->mail_host, 0)) != cudaSuccess) return e;
    }
    *(volatile int *)h->mail_host = -1;
    h->rowptr = rowptr; h->m = m; h->dev = dev;
    if ((e = cudaMemcpy(&h->nnz, rowptr + m, sizeof(int), cudaMemcpyDeviceToHost)) != cudaSuccess) { h->rowptr = nullptr; return e; }   // first sight only
  } else {
    const int seen = *(volatile int *)h->mail_host;   // whatever the latest finished call found; no synchronisation
    if (seen >= 0) h->nnz = seen;
    if (h->nnz <= 0 && (e = cudaMemcpy(&h->nnz, rowptr + m, sizeof(int), cudaMemcpyDeviceToHost)) != cudaSuccess) return e;
  }
  *nnz = h->nnz;
  *report = h->mail_dev;
  return cudaSuccess;
}

// ---- column-major SpMM (gespmmCsrSpMM with transpose_BC = false) ---------------------------------
// B[K,N] and C[M,N] column-major (ldB = K, ldC = M), src/ge-spmm/csrspmm_non_transpose.cu:479-600.
// One thread per output element, rows fastest so that C stores and rowptr loads coalesce.
__global__ void __launch_bounds__(256) spmm_colmajor_kernel(int M, int N, int K, const int *__restrict__ rowptr,
                                                            const int *__restrict__ col, const float *__restrict__ val,
                                                            const float *__restrict__ B, float *__restrict__ C) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (r >= M) return;
  const float *Bn = B + (size_t)n * K;
  float acc = 0.0f;
  const int end = __ldg(rowptr + r + 1);
  for (int p = __ldg(rowptr + r); p < end; p++) acc += (val ? __ldg(val + p) : 1.0f) * __ldg(Bn + __ldg(col + p));
  C[(size_t)n * M + r] = acc;
}

// out[c, r] = in[r, c] for a row-major in[rows, cols]: 32 x 32 tiles through shared memory, both sides coalesced.  The
// tile index is linear in blockIdx.x (either dimension may exceed the 65 535 limit of grid.y: products-like M = 2.4 M).
__global__ void __launch_bounds__(256) transpose_tiles_kernel(const float *__restrict__ in, int rows, int cols, float *__restrict__ out) {
  __shared__ float tile[32][33];
  const int tiles_c = (cols + 31) / 32;
  const int tr = blockIdx.x / tiles_c, tc = blockIdx.x - tr * tiles_c;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int r = tr * 32 + ty + i, c = tc * 32 + tx;
    if (r < rows && c < cols) tile[ty + i][tx] = __ldg(in + (size_t)r * cols + c);
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int c = tc * 32 + ty + i, r = tr * 32 + tx;
    if (r < rows && c < cols) out[(size_t)c * rows + r] = tile[tx][ty + i];
  }
}
cudaError_t transpose_tiles(const float *in, int rows, int cols, float *out, cudaStream_t s) {
  const int64_t tiles = (int64_t)((rows + 31) / 32) * ((cols + 31) / 32);
  if (tiles <= 0) return cudaSuccess;
  if (tiles > 0x7fffffffLL) return cudaErrorInvalidValue;
  transpose_tiles_kernel<<<(unsigned)tiles, 256, 0, s>>>(in, rows, cols, out);
  return cudaGetLastError();
}

// rowptr[r] = first position p with rowIdx[p] >= r (rowIdx ascending): one binary search per row
__global__ void __launch_bounds__(256) coo_to_rowptr_kernel(const int *__restrict__ rowIdx, int nnz, int nr, int *__restrict__ rowptr) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > nr) return;
  int lo = 0, hi = nnz;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(rowIdx + mid) < r) lo = mid + 1; else hi = mid;
  }
  rowptr[r] = lo;
}

// ---- completion barrier of the fused multicast epilogue ---------------------------------------------------------
// One thread per rank: add 1 to EVERY rank's arrival counter with a single multimem.red through the NVSwitch (release:
// ordered behind this rank's multicast stores of the step, which the stream placed before this kernel), then spin on the
// LOCAL copy until all `world` arrivals of this epoch are in (acquire).  Replaces a one-int ncclAllReduce (~40 us at 8
// GPUs) as the end-of-step barrier.  The counter only grows (epoch * world), compared wrap-safe.  A rank that never
// arrives would hang the others: the spin is bounded (~2 s) and then traps, loudly.
__global__ void mcast_barrier_kernel(unsigned *mc_ctr, const unsigned *local_ctr, unsigned target) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  asm volatile("multimem.red.release.sys.global.add.u32 [%0], %1;" ::"l"(mc_ctr), "r"(1u) : "memory");
  const long long t0 = clock64();
  unsigned v;
  while (true) {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(local_ctr) : "memory");
    if ((int)(v - target) >= 0) break;
    if (clock64() - t0 > 4000000000LL) { asm volatile("trap;"); }
  }
}

// ---- edge softmax --------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) edge_softmax_kernel(int M, int head, const int *__restrict__ rowptr,
                                                           const float *__restrict__ v, float *__restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M) return;
  const int s = __ldg(rowptr + warp), e = __ldg(rowptr + warp + 1);
  for (int h = 0; h < head; h++) {
    float mx = -INFINITY;
    for (int p = s + lane; p < e; p += 32) mx = fmaxf(mx, v[(size_t)p * head + h]);
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.0f;
    for (int p = s + lane; p < e; p += 32) sum += expf(v[(size_t)p * head + h] - mx);
#pragma unroll
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    for (int p = s + lane; p < e; p += 32) out[(size_t)p * head + h] = expf(v[(size_t)p * head + h] - mx) / sum;
  }
}

}  // namespace

extern "C" {

int dgs_version(void) { return 100; }
int dgs_cuda_version(void) { return CUDA_VERSION; }
const char *dgs_last_error(void) { return g_err; }
int dgs_sm_count(void) { return dgs::device_sm_count(); }
int dgs_spmm_last_path(void) { return dgs::spmm_last_path(); }
int dgs_set_option(const char *name, int value) { return dgs::set_option(name, value); }
void dgs_spmm_forget_graph_notes(void) { dgs::spmm_forget_graph_notes(); }
int dgs_legacy_scratch_release(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail(e, "dgs_legacy_scratch_release");
  if (dev < 0 || dev >= 64) return fail(cudaErrorInvalidDevice, "dgs_legacy_scratch_release");
  std::lock_guard<std::mutex> lk(g_mu);
  Scratch &s = g_scratch[dev];
  if (s.ptr == nullptr) return 0;
  if ((e = cudaStreamSynchronize(0)) != cudaSuccess) return fail(e, "dgs_legacy_scratch_release(sync)");
  e = cudaFree(s.ptr);
  s.ptr = nullptr; s.bytes = 0;
  return ok_or(e, "dgs_legacy_scratch_release(free)");
}
void dgs_sddmm_last_geometry(int *warps_per_cta, int *ctas_per_sm, int *edges_per_warp) {
  dgs::sddmm_last_geometry(warps_per_cta, ctas_per_sm, edges_per_warp);
}

size_t dgs_spmm_workspace_bytes(int N, int64_t nnz, int with_arg) {
  return dgs::spmm_workspace_bytes(N, nnz, with_arg != 0);
}

int dgs_spmm_csr_multi(int M, int N, int64_t nnz, const int *rowptr, const int *col, const float *val, const float *B,
                       int64_t ldb, int n_dst, float *const *dst, int64_t ldc, int reduce, int compute, void *workspace,
                       size_t workspace_bytes, void *stream) {
  if (n_dst < 1 || n_dst > dgs::kMaxDst || dst == nullptr) return fail(cudaErrorInvalidValue, "dgs_spmm_csr_multi(n_dst)");
  dgs::SpmmProblem p;
  p.M = M; p.N = N; p.nnz = nnz; p.rowptr = rowptr; p.col = col; p.val = val; p.B = B; p.ldb = ldb;
  p.n_dst = n_dst;
  for (int d = 0; d < n_dst; d++) p.dst[d] = dst[d];
  p.ldc = ldc; p.reduce = reduce; p.compute = compute;
  return ok_or(dgs::spmm_csr(p, workspace, workspace_bytes, (cudaStream_t)stream), "dgs_spmm_csr_multi");
}

int dgs_spmm_csr_mcast(int M, int N, int64_t nnz, const int *rowptr, const int *col, const float *val, const float *B,
                       int64_t ldb, float *mc_dst, int64_t ldc, int reduce, int compute, void *workspace,
                       size_t workspace_bytes, void *stream) {
  if (mc_dst == nullptr) return fail(cudaErrorInvalidValue, "dgs_spmm_csr_mcast(mc_dst)");
  dgs::SpmmProblem p;
  p.M = M; p.N = N; p.nnz = nnz; p.rowptr = rowptr; p.col = col; p.val = val; p.B = B; p.ldb = ldb;
  p.n_dst = 1; p.mcast = 1; p.dst[0] = mc_dst;
  p.ldc = ldc; p.reduce = reduce; p.compute = compute;
  return ok_or(dgs::spmm_csr(p, workspace, workspace_bytes, (cudaStream_t)stream), "dgs_spmm_csr_mcast");
}

int dgs_spmm_csr(int M, int N, int64_t nnz, const int *rowptr, const int *col, const float *val, const float *B,
                 int64_t ldb, float *C, int64_t ldc, int *E, int64_t lde, int reduce, int compute, void *workspace,
                 size_t workspace_bytes, void *stream) {
  dgs::SpmmProblem p;
  p.M = M; p.N = N; p.nnz = nnz; p.rowptr = rowptr; p.col = col; p.val = val; p.B = B; p.ldb = ldb;
  p.n_dst = 1; p.dst[0] = C; p.ldc = ldc; p.E = E; p.lde = lde; p.reduce = reduce; p.compute = compute;
  if (compute == DGS_MASKMUL) return fail(cudaErrorInvalidValue, "dgs_spmm_csr(compute): use dgs_spmm_csr_mask");
  return ok_or(dgs::spmm_csr(p, workspace, workspace_bytes, (cudaStream_t)stream), "dgs_spmm_csr");
}

int dgs_spmm_csr_k(int M, int K, int N, int64_t nnz, const int *rowptr, const int *col, const float *val, const float *B,
                   int64_t ldb, float *C, int64_t ldc, int *E, int64_t lde, int reduce, int compute, void *workspace,
                   size_t workspace_bytes, void *stream) {
  dgs::SpmmProblem p;
  p.M = M; p.K = K; p.N = N; p.nnz = nnz; p.rowptr = rowptr; p.col = col; p.val = val; p.B = B; p.ldb = ldb;
  p.n_dst = 1; p.dst[0] = C; p.ldc = ldc; p.E = E; p.lde = lde; p.reduce = reduce; p.compute = compute;
  if (compute == DGS_MASKMUL) return fail(cudaErrorInvalidValue, "dgs_spmm_csr_k(compute): use dgs_spmm_csr_mask");
  return ok_or(dgs::spmm_csr(p, workspace, workspace_bytes, (cudaStream_t)stream), "dgs_spmm_csr_k");
}

int dgs_spmm_csr_mask(int M, int N, int64_t nnz, const int *ptr, const int *idx, const float *val, const float *G,
                      int64_t ldg, const int *E, int64_t lde, float *out, int64_t ldo, void *workspace,
                      size_t workspace_bytes, void *stream) {
  dgs::SpmmProblem p;
  p.M = M; p.N = N; p.nnz = nnz; p.rowptr = ptr; p.col = idx; p.val = val; p.B = G; p.ldb = ldg;
  p.n_dst = 1; p.dst[0] = out; p.ldc = ldo; p.reduce = dgs::R_SUM; p.compute = dgs::C_MASK;
  p.mask = E; p.ldm = lde;
  return ok_or(dgs::spmm_csr(p, workspace, workspace_bytes, (cudaStream_t)stream), "dgs_spmm_csr_mask");
}

int dgs_mcast_barrier(void *mc_counter, const void *local_counter, unsigned target, void *stream) {
  if (mc_counter == nullptr || local_counter == nullptr) return fail(cudaErrorInvalidValue, "dgs_mcast_barrier(pointers)");
  mcast_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((unsigned *)mc_counter, (const unsigned *)local_counter, target);
  return ok_or(cudaGetLastError(), "dgs_mcast_barrier");
}

int dgs_sddmm_csr(int M, int K, int64_t nnz, const int *rowptr, const int *col, const float *D1, int64_t ld1,
                  const float *D2, int64_t ld2, const int *E, int mean, float *out, void *stream) {
  dgs::SddmmProblem p;
  p.M = M; p.K = K; p.nnz = nnz; p.rowptr = rowptr; p.col = col; p.D1 = D1; p.D2 = D2; p.ld1 = ld1; p.ld2 = ld2;
  p.E = E; p.mean = mean; p.out = out;
  if (rowptr == nullptr) return fail(cudaErrorInvalidValue, "dgs_sddmm_csr(rowptr)");
  return ok_or(dgs::sddmm(p, (cudaStream_t)stream), "dgs_sddmm_csr");
}

int dgs_sddmm_coo(int K, int64_t nnz, const int *row, const int *col, const float *D1, int64_t ld1, const float *D2,
                  int64_t ld2, float *out, void *stream) {
  dgs::SddmmProblem p;
  p.M = 0; p.K = K; p.nnz = nnz; p.row = row; p.col = col; p.D1 = D1; p.D2 = D2; p.ld1 = ld1; p.ld2 = ld2; p.out = out;
  if (row == nullptr) return fail(cudaErrorInvalidValue, "dgs_sddmm_coo(row)");
  return ok_or(dgs::sddmm(p, (cudaStream_t)stream), "dgs_sddmm_coo");
}

size_t dgs_csr2csc_workspace_bytes(int M, int ncols, int64_t nnz) { return dgs::csr2csc_workspace_bytes(M, ncols, nnz); }

int dgs_csr2csc(int M, int ncols, int64_t nnz, const int *rowptr, const int *col, const float *val, int *colptr,
                int *row, float *val_t, int *perm, void *workspace, size_t workspace_bytes, void *stream) {
  return ok_or(dgs::csr2csc(M, ncols, nnz, rowptr, col, val, colptr, row, val_t, perm, workspace, workspace_bytes,
                            (cudaStream_t)stream), "dgs_csr2csc");
}

int dgs_edge_softmax(int M, int head, const int *rowptr, const float *values, float *out, void *stream) {
  if (M <= 0 || head <= 0) return 0;
  const int blocks = (int)(((int64_t)M * 32 + 255) / 256);
  edge_softmax_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(M, head, rowptr, values, out);
  return ok_or(cudaGetLastError(), "dgs_edge_softmax");
}

// ---- sparse convolution ---------------------------------------------------------------------------
size_t dgs_spconv_workspace_bytes(int rows, int k_vol, int c_in, int c_out, int precision) {
  return dgs::spconv_workspace_bytes(rows, k_vol, c_in, c_out, precision);
}

int dgs_spconv_fwd(int in_nnz, int out_nnz, int k_vol, int c_in, int c_out, const int *kpos, const int *qkpos,
                   const int *in_map, const int *out_map, int64_t sum_nnz, const float *in_feats, const float *kernel,
                   float *out_feats, int separate_mid, int precision, void *workspace, size_t workspace_bytes,
                   void *stream) {
  if (separate_mid && in_nnz != out_nnz) return fail(cudaErrorInvalidValue, "dgs_spconv_fwd(separate_mid needs in_nnz == out_nnz)");
  dgs::SpconvProblem p;
  p.k_vol = k_vol; p.kdim = c_in; p.ndim = c_out; p.kpos = kpos; p.qkpos = qkpos; p.imap = in_map; p.omap = out_map;
  p.sum_nnz = sum_nnz; p.in = in_feats; p.ld_in = c_in; p.in_rows = in_nnz;
  p.W = kernel; p.w_sc = c_out; p.w_sn = 1; p.w_sk = (int64_t)c_in * c_out;
  p.out = out_feats; p.ld_out = c_out; p.out_rows = out_nnz; p.precision = precision; p.separate_mid = separate_mid;
  return ok_or(dgs::spconv_gemm(p, workspace, workspace_bytes, (cudaStream_t)stream), "dgs_spconv_fwd");
}

int dgs_spconv_bwd(int in_nnz, int out_nnz, int k_vol, int c_in, int c_out, const int *kpos, const int *qkpos,
                   const int *in_map, const int *out_map, int64_t sum_nnz, const float *out_grad, const float *in_feats,
                   const float *kernel, float *in_grad, float *kernel_grad, int separate_mid, int precision,
                   void *workspace, size_t workspace_bytes, void *stream) {
  if (separate_mid && in_nnz != out_nnz) return fail(cudaErrorInvalidValue, "dgs_spconv_bwd(separate_mid needs in_nnz == out_nnz)");
  if (in_grad != nullptr) {   // the forward with the maps swapped and W[k] read transposed
    dgs::SpconvProblem p;
    p.k_vol = k_vol; p.kdim = c_out; p.ndim = c_in; p.kpos = kpos; p.qkpos = qkpos; p.imap = out_map; p.omap = in_map;
    p.sum_nnz = sum_nnz; p.in = out_grad; p.ld_in = c_out; p.in_rows = out_nnz;
    p.W = kernel; p.w_sc = 1; p.w_sn = c_out; p.w_sk = (int64_t)c_in * c_out;
    p.out = in_grad; p.ld_out = c_in; p.out_rows = in_nnz; p.precision = precision; p.separate_mid = separate_mid;
    int rc = ok_or(dgs::spconv_gemm(p, workspace, workspace_bytes, (cudaStream_t)stream), "dgs_spconv_bwd(in_grad)");
    if (rc) return rc;
  }
  if (kernel_grad != nullptr) {
    int rc = ok_or(dgs::spconv_wgrad(k_vol, c_in, c_out, kpos, qkpos, in_map, out_map, sum_nnz, in_feats, c_in, in_nnz,
                                     out_grad, c_out, kernel_grad, precision, separate_mid, (cudaStream_t)stream),
                   "dgs_spconv_bwd(kernel_grad)");
    if (rc) return rc;
  }
  return 0;
}

// ---- kernel maps ------------------------------------------------------------------------------------
size_t dgs_kmap_workspace_bytes(int in_nnz, int out_nnz, int k_vol) { return dgs::kmap_workspace_bytes(in_nnz, out_nnz, k_vol); }

int dgs_kmap_downsample(int in_nnz, const int *in_coords, int sx, int sy, int sz, int *out_coords, int *out_nnz_dev,
                        void *workspace, size_t workspace_bytes, void *stream) {
  return ok_or(dgs::kmap_downsample(in_nnz, in_coords, sx, sy, sz, out_coords, out_nnz_dev, workspace, workspace_bytes,
                                    (cudaStream_t)stream), "dgs_kmap_downsample");
}

int dgs_kmap_build(int in_nnz, const int *in_coords, int out_nnz, const int *out_coords, int ksx, int ksy, int ksz,
                   int sx, int sy, int sz, int q, int skip_mid, int *imap, int *omap, int *knnz, int *kpos, int *qkpos,
                   void *workspace, size_t workspace_bytes, void *stream) {
  return ok_or(dgs::kmap_build(in_nnz, in_coords, out_nnz, out_coords, ksx, ksy, ksz, sx, sy, sz, q, skip_mid, imap, omap,
                               knnz, kpos, qkpos, workspace, workspace_bytes, (cudaStream_t)stream), "dgs_kmap_build");
}

size_t dgs_kmap_expand_workspace_bytes(int in_nnz, int k_vol) { return dgs::kmap_expand_workspace_bytes(in_nnz, k_vol); }

int dgs_kmap_downsample_expand(int in_nnz, const int *in_coords, int ksx, int ksy, int ksz, int sx, int sy, int sz, int px,
                               int py, int pz, const int *lo, const int *hi, int *out_coords, int *out_nnz_dev,
                               void *workspace, size_t workspace_bytes, void *stream) {
  return ok_or(dgs::kmap_downsample_expand(in_nnz, in_coords, ksx, ksy, ksz, sx, sy, sz, px, py, pz, lo, hi, out_coords,
                                           out_nnz_dev, workspace, workspace_bytes, (cudaStream_t)stream),
               "dgs_kmap_downsample_expand");
}

int dgs_kmap_build_ex(int in_nnz, const int *in_coords, int out_nnz, const int *out_coords, int ksx, int ksy, int ksz,
                      int sx, int sy, int sz, int px, int py, int pz, int subm, int q, int skip_mid, int *imap, int *omap,
                      int *knnz, int *kpos, int *qkpos, void *workspace, size_t workspace_bytes, void *stream) {
  return ok_or(dgs::kmap_build_ex(in_nnz, in_coords, out_nnz, out_coords, ksx, ksy, ksz, sx, sy, sz, px, py, pz, subm, q,
                                  skip_mid, imap, omap, knnz, kpos, qkpos, workspace, workspace_bytes, (cudaStream_t)stream),
               "dgs_kmap_build_ex");
}

// ---- host-buffer entry points --------------------------------------------------------------------
namespace {
struct HostStage {
  void *dev = nullptr;
  size_t bytes = 0;
  cudaStream_t stream = nullptr;      // H2D copies
  cudaStream_t s_compute = nullptr;   // kernels
  cudaStream_t s_d2h = nullptr;       // D2H copies
  cudaEvent_t ev_in[16] = {nullptr}, ev_done[16] = {nullptr};
  int device = -1;
};
thread_local HostStage g_stage;

cudaError_t stage_reserve(size_t need) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  HostStage &s = g_stage;
  if (s.device != dev) {
    if (s.dev) { cudaFree(s.dev); s.dev = nullptr; s.bytes = 0; }
    if (s.stream) { cudaStreamDestroy(s.stream); s.stream = nullptr; }
    if (s.s_compute) {
      cudaStreamDestroy(s.s_compute); cudaStreamDestroy(s.s_d2h);
      s.s_compute = s.s_d2h = nullptr;
      for (int i = 0; i < 16; i++) { cudaEventDestroy(s.ev_in[i]); cudaEventDestroy(s.ev_done[i]); }
    }
    if ((e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking)) != cudaSuccess) return e;
    s.device = dev;
  }
  if (s.bytes < need) {
    if (s.dev) { cudaFree(s.dev); s.dev = nullptr; s.bytes = 0; }
    if ((e = cudaMalloc(&s.dev, need)) != cudaSuccess) return e;
    s.bytes = need;
  }
  return cudaSuccess;
}
inline size_t up256(size_t x) { return (x + 255) / 256 * 256; }

// rowptr[i] -= base for a row block copied verbatim from the host CSR (so col/val can be addressed block-relative)
__global__ void __launch_bounds__(256) rebase_rowptr_kernel(int *rp, int n, int base) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) rp[i] -= base;
}
}  // namespace

int dgs_spmm_csr_host(int M, int K, int N, int64_t nnz, const int *rowptr, const int *col, const float *val,
                      const float *B, float *C, int *E, int reduce, int compute) {
  if (M < 0 || K < 0 || N < 0 || nnz < 0) return fail(cudaErrorInvalidValue, "dgs_spmm_csr_host(sizes)");
  if (M == 0 || N == 0) return 0;
  const bool with_arg = E != nullptr;
  // Row blocks of ~equal nnz, pipelined over three streams: while block i is multiplied, block i+1's CSR slice is
  // still crossing PCIe and block i-1's rows of C are already on their way back.  B goes first (every block needs it).
  int64_t blk_nnz = 8ll << 20;           // >= ~8 M nonzeros (64 MB of col+val) per block
  if (const char *env = getenv("DGS_HOST_BLOCK_NNZ")) { if (atoll(env) > 0) blk_nnz = atoll(env); }
  int nblk = (int)(nnz / blk_nnz);
  if (nblk < 1) nblk = 1;
  if (nblk > 16) nblk = 16;
  if (nblk > M) nblk = M;
  int r_begin[17];
  r_begin[0] = 0;
  for (int b = 1; b < nblk; b++) {   // first row whose start offset reaches the b-th nnz quantile
    const int64_t target = nnz * b / nblk;
    int lo = r_begin[b - 1], hi = M;
    while (lo < hi) {
      const int mid = lo + (hi - lo) / 2;
      if (rowptr[mid] < target) lo = mid + 1; else hi = mid;
    }
    r_begin[b] = lo;
  }
  r_begin[nblk] = M;
  size_t ws_need = 256;
  for (int b = 0; b < nblk; b++) {
    const size_t w = dgs::spmm_workspace_bytes(N, (int64_t)rowptr[r_begin[b + 1]] - rowptr[r_begin[b]], with_arg);
    if (w > ws_need) ws_need = w;
  }
  const size_t b_rowptr = up256(4 * ((size_t)M + 1 + nblk)), b_col = up256(4 * (size_t)nnz), b_val = val ? b_col : 0;
  const size_t b_B = up256(4 * (size_t)K * N), b_C = up256(4 * (size_t)M * N), b_E = with_arg ? b_C : 0;
  const size_t b_ws = up256(ws_need);
  cudaError_t e = stage_reserve(b_rowptr + b_col + b_val + b_B + b_C + b_E + b_ws);
  if (e != cudaSuccess) return fail(e, "dgs_spmm_csr_host(alloc)");
  HostStage &st = g_stage;
  if (st.s_compute == nullptr) {
    if ((e = cudaStreamCreateWithFlags(&st.s_compute, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "stream");
    if ((e = cudaStreamCreateWithFlags(&st.s_d2h, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "stream");
    for (int i = 0; i < 16; i++) {
      if ((e = cudaEventCreateWithFlags(&st.ev_in[i], cudaEventDisableTiming)) != cudaSuccess) return fail(e, "event");
      if ((e = cudaEventCreateWithFlags(&st.ev_done[i], cudaEventDisableTiming)) != cudaSuccess) return fail(e, "event");
    }
  }
  char *d = static_cast<char *>(st.dev);
  cudaStream_t s_in = st.stream, s_k = st.s_compute, s_out = st.s_d2h;
  int *d_rowptr = (int *)d; d += b_rowptr;
  int *d_col = (int *)d; d += b_col;
  float *d_val = val ? (float *)d : nullptr; d += b_val;
  float *d_B = (float *)d; d += b_B;
  float *d_C = (float *)d; d += b_C;
  int *d_E = with_arg ? (int *)d : nullptr; d += b_E;
  void *d_ws = d;
  if ((e = cudaMemcpyAsync(d_B, B, 4 * (size_t)K * N, cudaMemcpyHostToDevice, s_in)) != cudaSuccess) return fail(e, "h2d B");
  for (int b = 0; b < nblk; b++) {
    const int r0 = r_begin[b], r1 = r_begin[b + 1], rows = r1 - r0;
    const int64_t p0 = rowptr[r0], p1 = rowptr[r1], n = p1 - p0;
    int *d_rp = d_rowptr + r0 + b;   // block b's rowptr copy holds rows + 1 entries: keep the copies disjoint
    if ((e = cudaMemcpyAsync(d_rp, rowptr + r0, 4 * ((size_t)rows + 1), cudaMemcpyHostToDevice, s_in)) != cudaSuccess) return fail(e, "h2d rowptr");
    if (n > 0) {
      if ((e = cudaMemcpyAsync(d_col + p0, col + p0, 4 * (size_t)n, cudaMemcpyHostToDevice, s_in)) != cudaSuccess) return fail(e, "h2d col");
      if (val && (e = cudaMemcpyAsync(d_val + p0, val + p0, 4 * (size_t)n, cudaMemcpyHostToDevice, s_in)) != cudaSuccess) return fail(e, "h2d val");
    }
    if ((e = cudaEventRecord(st.ev_in[b], s_in)) != cudaSuccess) return fail(e, "event record");
    if ((e = cudaStreamWaitEvent(s_k, st.ev_in[b], 0)) != cudaSuccess) return fail(e, "event wait");
    if (rows > 0) {
      if (p0 != 0) rebase_rowptr_kernel<<<(rows + 1 + 255) / 256, 256, 0, s_k>>>(d_rp, rows + 1, (int)p0);
      int rc = dgs_spmm_csr_k(rows, K, N, n, d_rp, d_col + p0, d_val ? d_val + p0 : nullptr, d_B, N, d_C + (size_t)r0 * N, N,
                              d_E ? d_E + (size_t)r0 * N : nullptr, N, reduce, compute, d_ws, b_ws, s_k);
      if (rc) return rc;
    }
    if ((e = cudaEventRecord(st.ev_done[b], s_k)) != cudaSuccess) return fail(e, "event record");
    if ((e = cudaStreamWaitEvent(s_out, st.ev_done[b], 0)) != cudaSuccess) return fail(e, "event wait");
    if (rows > 0) {
      if ((e = cudaMemcpyAsync(C + (size_t)r0 * N, d_C + (size_t)r0 * N, 4 * (size_t)rows * N, cudaMemcpyDeviceToHost, s_out)) != cudaSuccess) return fail(e, "d2h C");
      if (with_arg && (e = cudaMemcpyAsync(E + (size_t)r0 * N, d_E + (size_t)r0 * N, 4 * (size_t)rows * N, cudaMemcpyDeviceToHost, s_out)) != cudaSuccess) return fail(e, "d2h E");
    }
  }
  if ((e = cudaStreamSynchronize(s_out)) != cudaSuccess) return fail(e, "dgs_spmm_csr_host(sync out)");
  if ((e = cudaStreamSynchronize(s_k)) != cudaSuccess) return fail(e, "dgs_spmm_csr_host(sync compute)");
  return ok_or(cudaStreamSynchronize(s_in), "dgs_spmm_csr_host(sync in)");
}

// ---- resident CSR: upload the matrix once, multiply many times from host memory -----------------------------------
// The real GNN use of the host path: A is fixed across layers / epochs, B changes every step.  dgs_spmm_csr_host re-sends
// 977 of its 1 037 MB (reddit@64) on every call; here the CSR crosses PCIe once (dgs_csr_upload), and a step moves only B
// in and C out.  A step: B host->device, then the row blocks of A (equal nnz) are multiplied one after the other while the
// finished rows of C are already on their way back (three streams, as in dgs_spmm_csr_host).
namespace {
struct ResidentCsr {
  int device = 0, M = 0, K = 0, nblk = 0;
  int64_t nnz = 0;
  bool has_val = false;
  int r_begin[17] = {0};
  int64_t p_begin[17] = {0};
  char *base = nullptr;          // one allocation: per-block rebased rowptr copies | col | val
  int *rowptr = nullptr, *col = nullptr;
  float *val = nullptr;
  char *scratch = nullptr;       // B | C | E | workspace for the current N (grown on demand)
  size_t scratch_bytes = 0;
  cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
  cudaEvent_t ev_in = nullptr, ev_done[16] = {nullptr};
};
}  // namespace

int dgs_csr_upload(int M, int K, int64_t nnz, const int *rowptr, const int *col, const float *val, void **handle) {
  if (M <= 0 || K <= 0 || nnz < 0 || rowptr == nullptr || (nnz > 0 && col == nullptr) || handle == nullptr)
    return fail(cudaErrorInvalidValue, "dgs_csr_upload(arguments)");
  ResidentCsr *h = new ResidentCsr();
  cudaError_t e = cudaGetDevice(&h->device);
  if (e != cudaSuccess) { delete h; return fail(e, "dgs_csr_upload(device)"); }
  h->M = M; h->K = K; h->nnz = nnz; h->has_val = val != nullptr;
  int nblk = (int)(nnz / (8ll << 20));   // >= ~8 M nonzeros per block, as the per-call host path
  if (nblk < 1) nblk = 1;
  if (nblk > 16) nblk = 16;
  if (nblk > M) nblk = M;
  h->nblk = nblk;
  for (int b = 1; b < nblk; b++) {
    const int64_t target = nnz * b / nblk;
    int lo = h->r_begin[b - 1], hi = M;
    while (lo < hi) {
      const int mid = lo + (hi - lo) / 2;
      if (rowptr[mid] < target) lo = mid + 1; else hi = mid;
    }
    h->r_begin[b] = lo;
  }
  h->r_begin[nblk] = M;
  for (int b = 0; b <= nblk; b++) h->p_begin[b] = rowptr[h->r_begin[b]];
  const size_t b_rowptr = up256(4 * ((size_t)M + 1 + nblk)), b_col = up256(4 * (size_t)nnz), b_val = h->has_val ? b_col : 0;
  auto bail = [&](cudaError_t err, const char *where) {
    if (h->base) cudaFree(h->base);
    if (h->s_in) cudaStreamDestroy(h->s_in);
    if (h->s_k) cudaStreamDestroy(h->s_k);
    if (h->s_out) cudaStreamDestroy(h->s_out);
    delete h;
    return fail(err, where);
  };
  if ((e = cudaMalloc((void **)&h->base, b_rowptr + b_col + b_val + 256)) != cudaSuccess) return bail(e, "dgs_csr_upload(alloc)");
  h->rowptr = (int *)h->base; h->col = (int *)(h->base + b_rowptr); h->val = h->has_val ? (float *)(h->base + b_rowptr + b_col) : nullptr;
  if ((e = cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "dgs_csr_upload(stream)");
  if ((e = cudaStreamCreateWithFlags(&h->s_k, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "dgs_csr_upload(stream)");
  if ((e = cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "dgs_csr_upload(stream)");
  if ((e = cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "dgs_csr_upload(event)");
  for (int i = 0; i < 16; i++)
    if ((e = cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming)) != cudaSuccess) return bail(e, "dgs_csr_upload(event)");
  for (int b = 0; b < nblk; b++) {   // block b's rowptr copy holds rows + 1 entries, rebased to the block's first nonzero
    const int r0 = h->r_begin[b], rows = h->r_begin[b + 1] - r0;
    int *d_rp = h->rowptr + r0 + b;
    if ((e = cudaMemcpyAsync(d_rp, rowptr + r0, 4 * ((size_t)rows + 1), cudaMemcpyHostToDevice, h->s_in)) != cudaSuccess) return bail(e, "dgs_csr_upload(h2d rowptr)");
    if (h->p_begin[b] != 0) rebase_rowptr_kernel<<<(rows + 1 + 255) / 256, 256, 0, h->s_in>>>(d_rp, rows + 1, (int)h->p_begin[b]);
  }
  if (nnz > 0) {
    if ((e = cudaMemcpyAsync(h->col, col, 4 * (size_t)nnz, cudaMemcpyHostToDevice, h->s_in)) != cudaSuccess) return bail(e, "dgs_csr_upload(h2d col)");
    if (h->has_val && (e = cudaMemcpyAsync(h->val, val, 4 * (size_t)nnz, cudaMemcpyHostToDevice, h->s_in)) != cudaSuccess) return bail(e, "dgs_csr_upload(h2d val)");
  }
  if ((e = cudaStreamSynchronize(h->s_in)) != cudaSuccess) return bail(e, "dgs_csr_upload(sync)");
  *handle = h;
  return 0;
}

int dgs_csr_free(void *handle) {
  ResidentCsr *h = static_cast<ResidentCsr *>(handle);
  if (h == nullptr) return 0;
  cudaStreamSynchronize(h->s_in); cudaStreamSynchronize(h->s_k); cudaStreamSynchronize(h->s_out);
  if (h->scratch) cudaFree(h->scratch);
  if (h->base) cudaFree(h->base);
  cudaEventDestroy(h->ev_in);
  for (int i = 0; i < 16; i++) cudaEventDestroy(h->ev_done[i]);
  cudaStreamDestroy(h->s_in); cudaStreamDestroy(h->s_k); cudaStreamDestroy(h->s_out);
  delete h;
  return 0;
}

int dgs_spmm_csr_resident_host(void *handle, int N, const float *B, float *C, int *E, int reduce, int compute) {
  ResidentCsr *h = static_cast<ResidentCsr *>(handle);
  if (h == nullptr || N <= 0 || B == nullptr || C == nullptr) return fail(cudaErrorInvalidValue, "dgs_spmm_csr_resident_host(arguments)");
  const bool with_arg = E != nullptr;
  size_t ws_need = 256;
  for (int b = 0; b < h->nblk; b++) {
    const size_t w = dgs::spmm_workspace_bytes(N, h->p_begin[b + 1] - h->p_begin[b], with_arg);
    if (w > ws_need) ws_need = w;
  }
  const size_t b_B = up256(4 * (size_t)h->K * N), b_C = up256(4 * (size_t)h->M * N), b_E = with_arg ? b_C : 0, b_ws = up256(ws_need);
  cudaError_t e;
  if (h->scratch_bytes < b_B + b_C + b_E + b_ws) {
    if (h->scratch) { cudaStreamSynchronize(h->s_k); cudaStreamSynchronize(h->s_out); cudaFree(h->scratch); h->scratch = nullptr; h->scratch_bytes = 0; }
    if ((e = cudaMalloc((void **)&h->scratch, b_B + b_C + b_E + b_ws)) != cudaSuccess) return fail(e, "dgs_spmm_csr_resident_host(alloc)");
    h->scratch_bytes = b_B + b_C + b_E + b_ws;
  }
  float *d_B = (float *)h->scratch, *d_C = (float *)(h->scratch + b_B);
  int *d_E = with_arg ? (int *)(h->scratch + b_B + b_C) : nullptr;
  void *d_ws = h->scratch + b_B + b_C + b_E;
  if ((e = cudaMemcpyAsync(d_B, B, 4 * (size_t)h->K * N, cudaMemcpyHostToDevice, h->s_in)) != cudaSuccess) return fail(e, "h2d B");
  if ((e = cudaEventRecord(h->ev_in, h->s_in)) != cudaSuccess) return fail(e, "event record");
  if ((e = cudaStreamWaitEvent(h->s_k, h->ev_in, 0)) != cudaSuccess) return fail(e, "event wait");
  for (int b = 0; b < h->nblk; b++) {
    const int r0 = h->r_begin[b], rows = h->r_begin[b + 1] - r0;
    const int64_t p0 = h->p_begin[b], n = h->p_begin[b + 1] - p0;
    if (rows > 0) {
      int rc = dgs_spmm_csr_k(rows, h->K, N, n, h->rowptr + r0 + b, h->col + p0, h->val ? h->val + p0 : nullptr, d_B, N,
                              d_C + (size_t)r0 * N, N, d_E ? d_E + (size_t)r0 * N : nullptr, N, reduce, compute, d_ws, b_ws, h->s_k);
      if (rc) return rc;
    }
    if ((e = cudaEventRecord(h->ev_done[b], h->s_k)) != cudaSuccess) return fail(e, "event record");
    if ((e = cudaStreamWaitEvent(h->s_out, h->ev_done[b], 0)) != cudaSuccess) return fail(e, "event wait");
    if (rows > 0) {
      if ((e = cudaMemcpyAsync(C + (size_t)r0 * N, d_C + (size_t)r0 * N, 4 * (size_t)rows * N, cudaMemcpyDeviceToHost, h->s_out)) != cudaSuccess) return fail(e, "d2h C");
      if (with_arg && (e = cudaMemcpyAsync(E + (size_t)r0 * N, d_E + (size_t)r0 * N, 4 * (size_t)rows * N, cudaMemcpyDeviceToHost, h->s_out)) != cudaSuccess) return fail(e, "d2h E");
    }
  }
  if ((e = cudaStreamSynchronize(h->s_out)) != cudaSuccess) return fail(e, "dgs_spmm_csr_resident_host(sync out)");
  return ok_or(cudaStreamSynchronize(h->s_k), "dgs_spmm_csr_resident_host(sync compute)");
}

int dgs_sddmm_csr_host(int M, int Kdim, int ncols, int64_t nnz, const int *rowptr, const int *col, const float *D1,
                       const float *D2, float *out) {
  if (M < 0 || Kdim < 0 || ncols < 0 || nnz < 0) return fail(cudaErrorInvalidValue, "dgs_sddmm_csr_host(sizes)");
  const size_t b_rowptr = up256(4 * ((size_t)M + 1)), b_col = up256(4 * (size_t)nnz);
  const size_t b_D1 = up256(4 * (size_t)M * Kdim), b_D2 = up256(4 * (size_t)ncols * Kdim), b_out = b_col;
  cudaError_t e = stage_reserve(b_rowptr + b_col + b_D1 + b_D2 + b_out);
  if (e != cudaSuccess) return fail(e, "dgs_sddmm_csr_host(alloc)");
  char *d = static_cast<char *>(g_stage.dev);
  cudaStream_t s = g_stage.stream;
  int *d_rowptr = (int *)d; d += b_rowptr;
  int *d_col = (int *)d; d += b_col;
  float *d_D1 = (float *)d; d += b_D1;
  float *d_D2 = (float *)d; d += b_D2;
  float *d_out = (float *)d;
  if ((e = cudaMemcpyAsync(d_rowptr, rowptr, 4 * ((size_t)M + 1), cudaMemcpyHostToDevice, s)) != cudaSuccess) return fail(e, "h2d rowptr");
  if ((e = cudaMemcpyAsync(d_col, col, 4 * (size_t)nnz, cudaMemcpyHostToDevice, s)) != cudaSuccess) return fail(e, "h2d col");
  if ((e = cudaMemcpyAsync(d_D1, D1, 4 * (size_t)M * Kdim, cudaMemcpyHostToDevice, s)) != cudaSuccess) return fail(e, "h2d D1");
  if ((e = cudaMemcpyAsync(d_D2, D2, 4 * (size_t)ncols * Kdim, cudaMemcpyHostToDevice, s)) != cudaSuccess) return fail(e, "h2d D2");
  int rc = dgs_sddmm_csr(M, Kdim, nnz, d_rowptr, d_col, d_D1, Kdim, d_D2, Kdim, nullptr, 0, d_out, s);
  if (rc) return rc;
  if ((e = cudaMemcpyAsync(out, d_out, 4 * (size_t)nnz, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return fail(e, "d2h out");
  return ok_or(cudaStreamSynchronize(s), "dgs_sddmm_csr_host(sync)");
}

// ---- peer memory (column-shard epilogue over NVLink) ---------------------------------------------
int dgs_ipc_export(const void *dptr, void *handle64, int64_t *offset) {
  typedef CUresult (*range_fn)(CUdeviceptr *, size_t *, CUdeviceptr);
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || fn == nullptr) return fail(e != cudaSuccess ? e : cudaErrorNotSupported, "dgs_ipc_export(entry point)");
  CUdeviceptr base = 0;
  size_t size = 0;
  if (((range_fn)fn)(&base, &size, (CUdeviceptr)dptr) != CUDA_SUCCESS) return fail(cudaErrorInvalidValue, "dgs_ipc_export(range)");
  cudaIpcMemHandle_t h;
  if ((e = cudaIpcGetMemHandle(&h, (void *)base)) != cudaSuccess) return fail(e, "dgs_ipc_export(handle)");
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  *offset = (int64_t)((CUdeviceptr)dptr - base);
  return 0;
}

int dgs_ipc_open(const void *handle64, void **base) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  return ok_or(cudaIpcOpenMemHandle(base, h, cudaIpcMemLazyEnablePeerAccess), "dgs_ipc_open");
}

int dgs_ipc_close(void *base) { return ok_or(cudaIpcCloseMemHandle(base), "dgs_ipc_close"); }

// ---- per-launch timing for bench.py's roofline leg -----------------------------------------------
int dgs_profile_enable(int on) { return dgs::profile_enable(on != 0); }
int dgs_profile_collect(int max_records, int *kernel_ids, float *ms) { return dgs::profile_collect(max_records, kernel_ids, ms); }

// ---- legacy dgSPARSE symbols (include/dgsparse.h) ------------------------------------------------

void gespmmCsrSpMM(const SpMatCsrDescr_t A, float *B, const int N, float *C, bool transpose_BC, gespmmAlg_t alg) {
  (void)alg;  // every reference algorithm computes the same C; one kernel serves them all
  if (A.nrow <= 0 || N <= 0) return;
  // Column-major operands (transpose_BC = false: ldB = ncol, ldC = nrow; src/ge-spmm/csrspmm_non_transpose.cu): the row-major
  // kernel between two tiled transposes — B^T in, C^T out through the legacy scratch — which costs 2 x 8 bytes per dense
  // element of HBM traffic instead of an uncoalesced 4-byte gather per nonzero and output element.  Measured against the
  // one-thread-per-element kernel it replaces (tools/exp_colmajor.py, profiles/r02_exp_colmajor.jsonl): p2p-Gnutella31 N = 32
  // 27.3 vs 29.3 us, arxiv-like N = 32 / 128 0.067 / 0.184 vs 1.11 / 1.92 ms, a quarter-size reddit-like matrix at N = 64
  // 0.42 vs 13.2 ms (the reference's best non-transposed algorithm: 34.6 us, 0.75 / 1.36 ms, 4.37 ms).  The simple kernel
  // remains for a descriptor without ncol (B cannot be sized) and behind spmm_colmajor = 0.
  bool colmajor_naive = false;
  if (!transpose_BC) colmajor_naive = dgs::option(dgs::OPT_SPMM_COLMAJOR) == 0 || A.ncol <= 0;
  if (colmajor_naive) {
    dim3 grid((A.nrow + 255) / 256, N);
    spmm_colmajor_kernel<<<grid, 256, 0, 0>>>(A.nrow, N, A.ncol, A.indptr, A.indices, A.data, B, C);
    legacy_report(ok_or(cudaGetLastError(), "gespmmCsrSpMM(colmajor)"), "gespmmCsrSpMM");
    return;
  }
  dgs::SpmmProblem p;
  p.M = A.nrow; p.K = A.ncol; p.N = N; p.nnz = A.nnz; p.rowptr = A.indptr; p.col = A.indices; p.val = A.data; p.B = B; p.ldb = N;
  p.n_dst = 1; p.dst[0] = C; p.ldc = N; p.reduce = dgs::R_SUM; p.compute = dgs::C_MUL;
  if (A.nnz < 0) {   // spmm_cuda(m, k, rowptr, ...): nnz lives on the device, the host only has a hint (see legacy_nnz_hint)
    int hint = 0, *report = nullptr;
    cudaError_t e = legacy_nnz_hint(A.indptr, A.nrow, &hint, &report);
    if (e != cudaSuccess) { legacy_report(fail(e, "gespmmCsrSpMM(nnz hint)"), "gespmmCsrSpMM"); return; }
    p.nnz = hint;
    if (hint > 0) { p.nnz_on_device = true; p.nnz_report = report; }
  }
  const size_t need = dgs::spmm_workspace_bytes(N, p.nnz, false);
  const size_t ws_bytes = (need + 255) / 256 * 256;
  const size_t bt_bytes = transpose_BC ? 0 : ((size_t)A.ncol * N * 4 + 255) / 256 * 256;
  const size_t ct_bytes = transpose_BC ? 0 : ((size_t)A.nrow * N * 4 + 255) / 256 * 256;
  void *ws = nullptr;
  cudaError_t e = legacy_scratch(ws_bytes + bt_bytes + ct_bytes, &ws);
  if (e != cudaSuccess) { legacy_report(fail(e, "gespmmCsrSpMM(scratch)"), "gespmmCsrSpMM"); return; }
  if (!transpose_BC) {
    float *Bt = reinterpret_cast<float *>(static_cast<char *>(ws) + ws_bytes);
    float *Ct = reinterpret_cast<float *>(static_cast<char *>(ws) + ws_bytes + bt_bytes);
    // column-major B[ncol, N] is a row-major [N, ncol] array: transpose it to the row-major [ncol, N] the kernel gathers from
    if ((e = transpose_tiles(B, N, A.ncol, Bt, nullptr)) != cudaSuccess) { legacy_report(fail(e, "gespmmCsrSpMM(B transpose)"), "gespmmCsrSpMM"); return; }
    p.B = Bt; p.dst[0] = Ct;
    int rc = ok_or(dgs::spmm_csr(p, ws, need, nullptr), "gespmmCsrSpMM");
    if (rc == 0) rc = ok_or(transpose_tiles(Ct, A.nrow, N, C, nullptr), "gespmmCsrSpMM(C transpose)");
    legacy_report(rc, "gespmmCsrSpMM");
    return;
  }
  legacy_report(ok_or(dgs::spmm_csr(p, ws, need, nullptr), "gespmmCsrSpMM"), "gespmmCsrSpMM");
}

// ---- the older SpMV/SpMM API (src/ge-spmm/gespmm_v2.h) ----------------------------------------------
void cuda_csr_spmm(int algo_code, int layout_code, int nr, int nc, int nv, int nnz, int *csrRowPtr, int *csrCol,
                   float *csrVal, float *vin, float *vout) {
  (void)algo_code;
  if (layout_code != 0 && layout_code != 1) { fprintf(stderr, "[dgsparse_b200] cuda_csr_spmm: wrong layout code %d\n", layout_code); return; }
  SpMatCsrDescr_t A = {nr, nc, nnz, csrRowPtr, csrCol, csrVal};
  gespmmCsrSpMM(A, vin, nv, vout, /*transpose_BC (row-major) =*/layout_code == 1, GESPMM_ALG_DEFAULT);
}

void cuda_csr_coo_spmm(SPMV_SPMM_ALG kAlg, DenseLayout layout, const int nr, const int nc, const int nnz, const int nv,
                       const int *rowPtr, const int *rowIdx, const int *colIdx, const float *values,
                       const float *dnInput, float *dnOutput) {
  (void)kAlg;
  int *rp = const_cast<int *>(rowPtr);
  if (rp == nullptr) {   // COO only: rebuild the row pointer from the sorted row indices in the legacy scratch
    if (rowIdx == nullptr) { legacy_report(fail(cudaErrorInvalidValue, "cuda_csr_coo_spmm(rowPtr and rowIdx NULL)"), "cuda_csr_coo_spmm"); return; }
    void *ws = nullptr;
    const size_t spmm_ws = dgs::spmm_workspace_bytes(nv, nnz, false);
    cudaError_t e = legacy_scratch(spmm_ws + 4 * ((size_t)nr + 1) + 512, &ws);
    if (e != cudaSuccess) { legacy_report(fail(e, "cuda_csr_coo_spmm(scratch)"), "cuda_csr_coo_spmm"); return; }
    rp = reinterpret_cast<int *>(static_cast<char *>(ws) + ((spmm_ws + 255) / 256 * 256));
    coo_to_rowptr_kernel<<<(nr + 1 + 255) / 256, 256, 0, 0>>>(rowIdx, nnz, nr, rp);
  }
  SpMatCsrDescr_t A = {nr, nc, nnz, rp, const_cast<int *>(colIdx), const_cast<float *>(values)};
  gespmmCsrSpMM(A, const_cast<float *>(dnInput), nv, dnOutput, layout == DENSE_ROW_MAJOR, GESPMM_ALG_DEFAULT);
}

void spmm_cuda(int m, int k, int *rowptr, int *colind, float *values, float *dense, float *out) {
  SpMatCsrDescr_t A = {m, 0, -1, rowptr, colind, values};
  gespmmCsrSpMM(A, dense, k, out, true, GESPMM_ALG_DEFAULT);
}

void spmm_cuda_no_edge_value(int m, int k, int *rowptr, int *colind, float *values, float *dense, float *out) {
  (void)values;
  spmm_cuda(m, k, rowptr, colind, nullptr, dense, out);
}

void sddmm_cuda_coo(int k, int nnz, int *rowind, int *colind, float *D1, float *D2, float *out) {
  legacy_report(dgs_sddmm_coo(k, nnz, rowind, colind, D1, k, D2, k, out, nullptr), "sddmm_cuda_coo");
}

void sddmm_cuda_csr(int m, int k, int nnz, int *rowptr, int *colind, float *D1, float *D2, float *out) {
  legacy_report(dgs_sddmm_csr(m, k, nnz, rowptr, colind, D1, k, D2, k, nullptr, 0, out, nullptr), "sddmm_cuda_csr");
}

void edge_softmax_cuda(int mrows, int head, int *rowptr, float *values, float *softmax) {
  legacy_report(dgs_edge_softmax(mrows, head, rowptr, values, softmax, nullptr), "edge_softmax_cuda");
}

}  // extern "C"
