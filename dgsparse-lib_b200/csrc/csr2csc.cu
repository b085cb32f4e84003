// csr2csc.cu — exact CSR -> CSC transpose for sm_100a (integer work, bit-exact, deterministic).
//
// Replaces the reference's cuSPARSE call (cusparseCsr2cscEx2 ALG1, include/cuda/csr2csc.cuh:8-26,
// launched from csr2csc_cuda, src/cuda/spmm_cuda.cu:384-414) and the float32-arange permutation
// trick of dgsparse/storage.py:159-174 (inexact above 2^24 nnz): we return the exact
// int32 permutation perm[q] = CSR position of the q-th CSC entry.
//
// The result is pinned by test/test_csr2csr.py:42-49 against scipy tocsc(): entries of a column
// keep their CSR order (rows ascending).  That is a STABLE sort of the nnz positions by column
// index, done here as an LSD radix sort (8-bit digits over ceil(log2(ncols)) bits) with
//   hist:    per-tile digit counts                       (coalesced streaming read of the keys)
//   scan:    exclusive scan of counts[digit][tile]       (three small kernels)
//   scatter: tile-stable ranking with __match_any_sync + per-warp digit counters in shared memory
// colptr comes from a direct column histogram + the same scan.  No atomics decide any output
// position, so every run gives identical bytes.  All traffic is streaming/HBM-bound:
// ~20 B per nnz per pass.
#include <algorithm>
#include "common.cuh"
#include "spmm.h"

namespace dgs {

constexpr int kRsThreads = 256;
constexpr int kRsItems = 16;
constexpr int kRsTile = kRsThreads * kRsItems;   // keys per block
constexpr int kRsWarps = kRsThreads / 32;
constexpr int kRadix = 256;

// ---- generic exclusive scan over int32 (n up to 2^31), in place ---------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ int block_exclusive_scan(int x, int *total, int *s_warp) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = x;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_warp[w] = inc;
  __syncthreads();
  if (w == 0) {
    int v = lane < (int)(blockDim.x >> 5) ? s_warp[lane] : 0;
    int vi = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, vi, o);
      if (lane >= o) vi += t;
    }
    s_warp[lane] = vi - v;            // exclusive warp offsets
    if (lane == 31) s_warp[32] = vi;  // block total
  }
  __syncthreads();
  const int res = inc - x + s_warp[w];
  *total = s_warp[32];
  __syncthreads();
  return res;
}

// phase 1: per-tile sums
__global__ void __launch_bounds__(kScanThreads) scan_tile_sums(const int *__restrict__ a, int64_t n, int *__restrict__ sums) {
  __shared__ int s_warp[33];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; i++) if (base + i < n) s += a[base + i];
  int total;
  block_exclusive_scan(s, &total, s_warp);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// phase 2: one block scans the tile sums in place (exclusive), looping with a carry
__global__ void __launch_bounds__(1024) scan_sums_inplace(int *__restrict__ sums, int n) {
  __shared__ int s_warp[33];
  int carry = 0;
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const int x = i < n ? sums[i] : 0;
    int total;
    const int ex = block_exclusive_scan(x, &total, s_warp);
    if (i < n) sums[i] = carry + ex;
    carry += total;
  }
}

// phase 3: per-tile exclusive scan + tile offset, in place
__global__ void __launch_bounds__(kScanThreads) scan_apply(int *__restrict__ a, int64_t n, const int *__restrict__ sums) {
  __shared__ int s_warp[33];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int v[kScanItems];
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; i++) { v[i] = (base + i < n) ? a[base + i] : 0; s += v[i]; }
  int total;
  int run = block_exclusive_scan(s, &total, s_warp) + sums[blockIdx.x];
#pragma unroll
  for (int i = 0; i < kScanItems; i++) {
    if (base + i < n) a[base + i] = run;
    run += v[i];
  }
}

static size_t scan_scratch_ints(int64_t n) { return (size_t)((n + kScanTile - 1) / kScanTile) + 1; }

static cudaError_t exclusive_scan_inplace(int *a, int64_t n, int *scratch, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  const int tiles = (int)((n + kScanTile - 1) / kScanTile);
  scan_tile_sums<<<tiles, kScanThreads, 0, s>>>(a, n, scratch);
  scan_sums_inplace<<<1, 1024, 0, s>>>(scratch, tiles);
  scan_apply<<<tiles, kScanThreads, 0, s>>>(a, n, scratch);
  return cudaGetLastError();
}

// ---- column histogram -> colptr -----------------------------------------------------------------
// counts[c] += 1 for every nnz; integer atomics commute, so the result is exact and deterministic.
__global__ void __launch_bounds__(256) col_histogram(const int *__restrict__ col, int nnz, int *__restrict__ counts) {
  const int stride = gridDim.x * blockDim.x;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < nnz; p += stride) atomicAdd(counts + __ldcs(col + p), 1);
}

// ---- radix passes -------------------------------------------------------------------------------
// element e of tile t:  warp w owns [w*32*ITEMS, (w+1)*32*ITEMS), item i of lane l is w*32*ITEMS + i*32 + l
__device__ __forceinline__ int tile_elem(int w, int i, int lane) { return w * 32 * kRsItems + i * 32 + lane; }

__global__ void __launch_bounds__(kRsThreads) radix_hist(const int *__restrict__ keys, int nnz, int shift,
                                                         int num_tiles, int *__restrict__ counts /*[256][num_tiles]*/) {
  __shared__ int s_hist[kRadix];
  s_hist[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kRsTile;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < kRsItems; i++) {
    const int64_t e = base + tile_elem(w, i, lane);
    if (e < nnz) atomicAdd(&s_hist[(__ldg(keys + e) >> shift) & (kRadix - 1)], 1);
  }
  __syncthreads();
  counts[(size_t)threadIdx.x * num_tiles + blockIdx.x] = s_hist[threadIdx.x];
}

// FIRST: values are the positions themselves (key array = col).  LAST: only values are written.
template <bool FIRST, bool LAST>
__global__ void __launch_bounds__(kRsThreads) radix_scatter(const int *__restrict__ keys_in, const int *__restrict__ vals_in,
                                                            int *__restrict__ keys_out, int *__restrict__ vals_out, int nnz,
                                                            int shift, int num_tiles, const int *__restrict__ offsets) {
  __shared__ int s_cnt[kRsWarps][kRadix];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kRsWarps * kRadix; i += kRsThreads) (&s_cnt[0][0])[i] = 0;
  __syncthreads();

  const int64_t base = (int64_t)blockIdx.x * kRsTile;
  int key[kRsItems], rank[kRsItems];
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int i = 0; i < kRsItems; i++) {
    const int64_t e = base + tile_elem(w, i, lane);
    const bool ok = e < nnz;
    key[i] = ok ? __ldg(keys_in + e) : 0x7fffffff;
    const int d = ok ? ((key[i] >> shift) & (kRadix - 1)) : kRadix;   // invalid lanes form their own group
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const int leader = __ffs(peers) - 1;
    int prev = 0;
    if (ok && lane == leader) {
      prev = s_cnt[w][d];
      s_cnt[w][d] = prev + __popc(peers);
    }
    prev = __shfl_sync(0xffffffffu, prev, leader);
    rank[i] = prev + __popc(peers & lt);
    __syncwarp();
  }
  __syncthreads();
  {  // digit threadIdx.x: exclusive scan over the warps, seeded with this tile's global offset
    const int d = threadIdx.x;
    int run = offsets[(size_t)d * num_tiles + blockIdx.x];
#pragma unroll
    for (int ww = 0; ww < kRsWarps; ww++) {
      const int t = s_cnt[ww][d];
      s_cnt[ww][d] = run;
      run += t;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kRsItems; i++) {
    const int64_t e = base + tile_elem(w, i, lane);
    if (e < nnz) {
      const int d = (key[i] >> shift) & (kRadix - 1);
      const int dst = s_cnt[w][d] + rank[i];
      if (!LAST) keys_out[dst] = key[i];
      vals_out[dst] = FIRST ? (int)e : __ldg(vals_in + e);
    }
  }
}

// row[q] = row owning CSR position perm[q]; val_t[q] = val[perm[q]]
__global__ void __launch_bounds__(256) finish_transpose(const int *__restrict__ perm, int nnz, const int *__restrict__ rowptr,
                                                        int M, const float *__restrict__ val, int *__restrict__ row,
                                                        float *__restrict__ val_t) {
  const int stride = gridDim.x * blockDim.x;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nnz; q += stride) {
    const int p = __ldcs(perm + q);
    if (row) row[q] = row_of_nnz(rowptr, M, p);
    if (val_t) val_t[q] = __ldg(val + p);
  }
}


static inline size_t up256(size_t x) { return (x + 255) / 256 * 256; }

static int key_bits(int ncols) {
  int b = 1;
  while (b < 31 && (1 << b) < ncols) b++;
  return b;
}

size_t csr2csc_workspace_bytes(int M, int ncols, int64_t nnz) {
  (void)M;
  if (nnz < 0) nnz = 0;
  const size_t tiles = (size_t)((nnz + kRsTile - 1) / kRsTile);
  const size_t counts = (size_t)kRadix * tiles;
  size_t b = 0;
  b += up256(4 * (size_t)nnz) * 4;                     // key/value ping-pong buffers
  b += up256(4 * counts);                              // per-tile digit counts
  size_t sc = scan_scratch_ints((int64_t)counts);
  const size_t sc2 = scan_scratch_ints((int64_t)ncols + 1);
  if (sc2 > sc) sc = sc2;
  b += up256(4 * sc);
  return b + 256;
}

cudaError_t csr2csc(int M, int ncols, int64_t nnz64, const int *rowptr, const int *col, const float *val, int *colptr,
                    int *row, float *val_t, int *perm, void *workspace, size_t workspace_bytes, cudaStream_t s) {
  if (M < 0 || ncols < 0 || nnz64 < 0 || nnz64 > 0x7fffffff) return cudaErrorInvalidValue;
  const int nnz = (int)nnz64;
  if (workspace == nullptr || workspace_bytes < csr2csc_workspace_bytes(M, ncols, nnz64)) return cudaErrorInvalidValue;
  const int sms = device_sm_count();
  cudaError_t e;

  char *w = static_cast<char *>(workspace);
  const size_t arr = up256(4 * (size_t)nnz);
  int *kA = reinterpret_cast<int *>(w);
  int *vA = reinterpret_cast<int *>(w + arr);
  int *kB = reinterpret_cast<int *>(w + 2 * arr);
  int *vB = reinterpret_cast<int *>(w + 3 * arr);
  const int tiles = (nnz + kRsTile - 1) / kRsTile;
  const size_t counts_n = (size_t)kRadix * tiles;
  int *counts = reinterpret_cast<int *>(w + 4 * arr);
  int *scratch = reinterpret_cast<int *>(w + 4 * arr + up256(4 * counts_n));

  // colptr: histogram at colptr[c] (slot ncols stays 0), then an exclusive scan over ncols + 1 entries
  if ((e = cudaMemsetAsync(colptr, 0, sizeof(int) * ((size_t)ncols + 1), s)) != cudaSuccess) return e;
  if (nnz == 0) return cudaSuccess;
  {
    const int blocks = (int)std::min<int64_t>(((int64_t)nnz + 255) / 256, (int64_t)sms * 16);
    col_histogram<<<blocks, 256, 0, s>>>(col, nnz, colptr);
  }
  if ((e = exclusive_scan_inplace(colptr, (int64_t)ncols + 1, scratch, s)) != cudaSuccess) return e;

  // stable sort of positions by column
  const int bits = key_bits(ncols);
  const int passes = (bits + 7) / 8;
  const int *kin = col;
  const int *vin = nullptr;
  int *kout = kA, *vout = vA;
  for (int pass = 0; pass < passes; pass++) {
    const int shift = pass * 8;
    const bool first = pass == 0, last = pass == passes - 1;
    if (last && perm != nullptr) vout = perm;   // the last pass writes straight into the caller's perm
    radix_hist<<<tiles, kRsThreads, 0, s>>>(kin, nnz, shift, tiles, counts);
    if ((e = exclusive_scan_inplace(counts, (int64_t)counts_n, scratch, s)) != cudaSuccess) return e;
    if (first && last) radix_scatter<true, true><<<tiles, kRsThreads, 0, s>>>(kin, vin, kout, vout, nnz, shift, tiles, counts);
    else if (first) radix_scatter<true, false><<<tiles, kRsThreads, 0, s>>>(kin, vin, kout, vout, nnz, shift, tiles, counts);
    else if (last) radix_scatter<false, true><<<tiles, kRsThreads, 0, s>>>(kin, vin, kout, vout, nnz, shift, tiles, counts);
    else radix_scatter<false, false><<<tiles, kRsThreads, 0, s>>>(kin, vin, kout, vout, nnz, shift, tiles, counts);
    kin = kout; vin = vout;
    if (kout == kA) { kout = kB; vout = vB; } else { kout = kA; vout = vA; }
  }
  const int *perm_out = vin;
  if (row != nullptr || (val_t != nullptr && val != nullptr)) {
    const int blocks = (int)std::min<int64_t>(((int64_t)nnz + 255) / 256, (int64_t)sms * 32);
    finish_transpose<<<blocks, 256, 0, s>>>(perm_out, nnz, rowptr, M, val, row, (val != nullptr) ? val_t : nullptr);
  }
  return cudaGetLastError();
}

}  // namespace dgs
