// spmm_slab.cu — column-slab tiling of the SpMM for dense operands far larger than the L2 (products-like: B = 2.45 M rows).
//
// With B >> L2 every nonzero of the plain kernel pulls its 256 B row slice from HBM: 57 GB of DRAM reads for 3.5 GB of
// algorithmic bytes at products@128, 8.7 ms at the DRAM peak (narrower column panels were measured and are
// worse).  The gather only becomes an L2 hit if the rows it can touch fit the L2 — ~65 MB of 256 B rows (the capacity curve in
// profiles/r02_exp_l2_capacity.jsonl).  So A is cut by COLUMN RANGE into S slabs of K_s rows of B (K_s x 256 B ~ 0.45 L2):
//   1. slab_count_kernel:   one thread per (slab, row): how many nonzeros of the row fall in the slab (the row's columns are
//                           sorted, so that is two binary searches), in slab-major order;
//   2. an exclusive scan of those counts = the row pointers of S stacked CSR matrices that share one nnz stream;
//   3. slab_scatter_kernel: col (and val, pre-divided by the row degree for MEAN) copied into that order;
//   4. one pass of the ordinary row-segment kernel per slab over [rowptr_s[0], rowptr_s[M]) — its gathers now hit a B slab
//      that stays L2-resident — the first pass writing C, the later ones COMBINING the finished row with what C holds
//      (sum: add; max / min: compare).  A row's nonzeros are sorted by column, so the slabs are visited in the row's own
//      nnz order.
// Traffic: 3 GB for the partition + col / val once per column panel + a read-modify-write of the C rows a slab touches,
// against 57 GB.  sum, mean, max, min (no arg index), with or without edge values, local destination only.
#include <cstdint>
#include "options.h"
#include "spmm.h"
#include "spmm_rowseg.cuh"

namespace dgs {

cudaError_t exclusive_scan_i32(int *a, int64_t n, int *scratch, cudaStream_t s);   // csr2csc.cu
size_t exclusive_scan_scratch_ints(int64_t n);
int device_l2_bytes();

namespace {

// first position in [lo, hi) with col[pos] >= x
__device__ __forceinline__ int lower_bound_col(const int *__restrict__ col, int lo, int hi, int x) {
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(col + mid) < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// cnt[s * M + r] = nonzeros of row r with column in [s * Ks, (s + 1) * Ks)
__global__ void __launch_bounds__(256) slab_count_kernel(int M, int S, int Ks, const int *__restrict__ rowptr,
                                                         const int *__restrict__ col, int *__restrict__ cnt) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)S * M) return;
  const int s = (int)(t / M), r = (int)(t % M);
  const int lo = __ldg(rowptr + r), hi = __ldg(rowptr + r + 1);
  const int b = lower_bound_col(col, lo, hi, s * Ks);
  const int e = (s + 1 == S) ? hi : lower_bound_col(col, b, hi, (s + 1) * Ks);
  cnt[t] = e - b;
}

// col_p / val_p[rowptr_lin[s * M + r] ...] = the row's nonzeros of slab s;  MODE 0: no values, 1: copy, 2: val / deg, 3: 1 / deg
template <int MODE>
__global__ void __launch_bounds__(256) slab_scatter_kernel(int M, int S, int Ks, const int *__restrict__ rowptr,
                                                           const int *__restrict__ col, const float *__restrict__ val,
                                                           const int *__restrict__ rowptr_lin, int *__restrict__ col_p,
                                                           float *__restrict__ val_p) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)S * M) return;
  const int dst = __ldg(rowptr_lin + t), n = __ldg(rowptr_lin + t + 1) - dst;
  if (n == 0) return;
  const int s = (int)(t / M), r = (int)(t % M);
  const int lo = __ldg(rowptr + r), hi = __ldg(rowptr + r + 1);
  const int b = lower_bound_col(col, lo, hi, s * Ks);
  const float inv = (MODE >= 2) ? 1.0f / (float)(hi - lo) : 1.0f;
  for (int i = 0; i < n; i++) {
    col_p[dst + i] = __ldcs(col + b + i);
    if (MODE == 1) val_p[dst + i] = __ldcs(val + b + i);
    if (MODE == 2) val_p[dst + i] = __ldcs(val + b + i) * inv;
    if (MODE == 3) val_p[dst + i] = inv;
  }
}

inline size_t up256(size_t x) { return (x + 255) / 256 * 256; }

struct SlabLayout {
  int S = 0, Ks = 0;
  int64_t nnz_hint = 0;            // nonzeros per slab the grids are sized for
  size_t inner = 0, off_rowptr = 0, off_scan = 0, off_col = 0, off_val = 0, total = 0;
};

bool slab_layout(int M, int K, int N, int64_t nnz, bool need_val, SlabLayout *L) {
  int rows = option(OPT_SPMM_SLAB_ROWS);
  if (rows < 64) rows = (int)(0.45 * (double)device_l2_bytes() / 256.0);
  rows = (rows + 63) / 64 * 64;
  const int S = (K + rows - 1) / rows;
  if (S < 2 || S > 256 || (int64_t)S * M + 1 > 0x7fffffffLL) return false;
  L->S = S; L->Ks = rows;
  L->nnz_hint = (nnz + S - 1) / S;
  L->inner = up256(spmm_workspace_bytes(N, L->nnz_hint, false));
  const size_t lin = (size_t)S * M + 1;
  L->off_rowptr = L->inner;
  L->off_scan = L->off_rowptr + up256(lin * 4);
  L->off_col = L->off_scan + up256(exclusive_scan_scratch_ints((int64_t)lin) * 4);
  L->off_val = L->off_col + up256((size_t)nnz * 4);
  L->total = L->off_val + (need_val ? up256((size_t)nnz * 4) : 0);
  return true;
}

}  // namespace

// By size: the 64-column panel of B (K x 256 B) must be far beyond the L2 and the matrix large enough to pay for the
// partition.  Option spmm_slab: 0 never, 1 whenever the kernels allow it (tests).
bool spmm_slab_wanted(int M, int K, int N, int64_t nnz) {
  (void)M;
  const int mode = option(OPT_SPMM_SLAB);
  if (mode == 0 || N < 64 || nnz < 64) return false;
  if (mode == 1) return true;
  return (double)K * 256.0 >= 1.2 * (double)device_l2_bytes() && nnz >= (32 << 20);
}

// Is the column-slab path the right one for this problem, and does the caller's workspace hold it?
bool spmm_slab_eligible(const SpmmProblem &p, bool can_vec4, int comp, size_t workspace_bytes) {
  if (p.slab_pass || !can_vec4 || p.E != nullptr || p.n_dst != 1 || p.mcast) return false;
  if (comp != C_MUL && comp != C_COPY) return false;
  const int K = p.K > 0 ? p.K : p.M;
  if (!spmm_slab_wanted(p.M, K, p.N, p.nnz)) return false;
  SlabLayout L;
  if (!slab_layout(p.M, K, p.N, p.nnz, p.val != nullptr || p.reduce == R_MEAN, &L)) return false;
  return workspace_bytes >= L.total;
}

size_t spmm_slab_workspace_bytes(int M, int K, int N, int64_t nnz) {
  SlabLayout L;
  if (M <= 0 || K <= 0 || N <= 0 || nnz <= 0 || !slab_layout(M, K, N, nnz, true, &L)) return 0;
  return L.total + 256;
}

cudaError_t spmm_csr_slabbed(const SpmmProblem &p, void *workspace, size_t workspace_bytes, cudaStream_t stream) {
  const int K = p.K > 0 ? p.K : p.M;
  const bool mean = p.reduce == R_MEAN;
  const bool need_val = p.val != nullptr || mean;
  SlabLayout L;
  if (!slab_layout(p.M, K, p.N, p.nnz, need_val, &L) || workspace_bytes < L.total) return cudaErrorInvalidValue;
  char *w = static_cast<char *>(workspace);
  int *rowptr_lin = reinterpret_cast<int *>(w + L.off_rowptr);
  int *scan_tmp = reinterpret_cast<int *>(w + L.off_scan);
  int *col_p = reinterpret_cast<int *>(w + L.off_col);
  float *val_p = need_val ? reinterpret_cast<float *>(w + L.off_val) : nullptr;
  const int64_t cells = (int64_t)L.S * p.M;
  const int blocks = (int)((cells + 255) / 256);
  cudaError_t e;
  {
    ProfileScope prof(7, stream);
    slab_count_kernel<<<blocks, 256, 0, stream>>>(p.M, L.S, L.Ks, p.rowptr, p.col, rowptr_lin);
    if ((e = cudaMemsetAsync(rowptr_lin + cells, 0, sizeof(int), stream)) != cudaSuccess) return e;
    if ((e = exclusive_scan_i32(rowptr_lin, cells + 1, scan_tmp, stream)) != cudaSuccess) return e;
    if (!need_val) slab_scatter_kernel<0><<<blocks, 256, 0, stream>>>(p.M, L.S, L.Ks, p.rowptr, p.col, p.val, rowptr_lin, col_p, val_p);
    else if (!mean) slab_scatter_kernel<1><<<blocks, 256, 0, stream>>>(p.M, L.S, L.Ks, p.rowptr, p.col, p.val, rowptr_lin, col_p, val_p);
    else if (p.val) slab_scatter_kernel<2><<<blocks, 256, 0, stream>>>(p.M, L.S, L.Ks, p.rowptr, p.col, p.val, rowptr_lin, col_p, val_p);
    else slab_scatter_kernel<3><<<blocks, 256, 0, stream>>>(p.M, L.S, L.Ks, p.rowptr, p.col, p.val, rowptr_lin, col_p, val_p);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  for (int s = 0; s < L.S; s++) {
    SpmmProblem q = p;
    q.slab_pass = 1;
    q.accum = s > 0;
    q.rowptr_full = p.rowptr;
    q.rowptr = rowptr_lin + (size_t)s * p.M;     // M + 1 entries: the next slab's first entry closes this one
    q.col = col_p;
    q.val = val_p;
    q.nnz = L.nnz_hint;                          // the true extent [rowptr[0], rowptr[M]) is read on the device
    q.nnz_on_device = true;
    q.nnz_report = nullptr;
    if (mean) { q.reduce = R_SUM; q.compute = C_MUL; }   // 1 / deg is folded into val_p
    if ((e = spmm_csr(q, workspace, L.inner, stream)) != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace dgs
