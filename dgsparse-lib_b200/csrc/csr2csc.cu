// csr2csc.cu — exact CSR -> CSC transpose for sm_100a (integer work, bit-exact, deterministic).
//
// Replaces the reference's cuSPARSE call (cusparseCsr2cscEx2 ALG1, include/cuda/csr2csc.cuh:8-26,
// launched from csr2csc_cuda, src/cuda/spmm_cuda.cu:384-414) and the float32-arange permutation
// trick of dgsparse/storage.py:159-174 (inexact above 2^24 nnz): we return the exact
// int32 permutation perm[q] = CSR position of the q-th CSC entry.
//
// The result is pinned by test/test_csr2csr.py:42-49 against scipy tocsc(): entries of a column
// keep their CSR order (rows ascending).  That is a STABLE sort of the nnz positions by column
// index, done here as an LSD radix sort over ceil(log2(ncols)) bits with digits of up to 9 bits
// (reddit / arxiv: 18 bits = 2 passes; products: 22 bits = 3), every element carrying its whole
// CSC record {column, CSR position | row, value} so that the last pass writes row / val_t / perm
// directly and nothing is gathered at random afterwards:
//   rows:    row index of every nnz position, one warp per CSR row              (coalesced stores)
//   hist:    per-tile digit counts                                              (streaming read of the keys)
//   scan:    exclusive scan of counts[digit][tile]                              (three small kernels)
//   scatter: tile-stable ranking (__match_any_sync + per-warp digit counters), then the tile is SORTED IN SHARED
//            MEMORY and written out as runs — consecutive threads store consecutive addresses of one bucket, 8 B per
//            element — instead of one random 4 B store per element and array (the first version: 3.7 ms per pass
//            and a 3.1 ms gather at the end on the reddit-like matrix, 13.4 ms in all)
//   colptr:  from the boundaries of the sorted column stream (no atomics)
// No atomics decide any output position, so every run gives identical bytes.  All traffic is streaming:
// 8 + 16 B per nnz in the first pass, 16 + 16 in the others, + 4 (hist) per pass.
#include <algorithm>
#include "common.cuh"
#include "spmm.h"

namespace dgs {

constexpr int kRsThreads = 256;
constexpr int kRsItems = 16;
constexpr int kRsTile = kRsThreads * kRsItems;   // elements per block
constexpr int kRsWarps = kRsThreads / 32;
constexpr int kMaxDigitBits = 9;
constexpr int kMaxRadix = 1 << kMaxDigitBits;

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

// ---- row index of every nnz position ------------------------------------------------------------
__global__ void __launch_bounds__(256) expand_rows(const int *__restrict__ rowptr, int M, int *__restrict__ row) {
  const int lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < M; r += nwarps) {
    const int s = __ldg(rowptr + r), e = __ldg(rowptr + r + 1);
    for (int p = s + lane; p < e; p += 32) row[p] = r;
  }
}

// ---- radix passes -------------------------------------------------------------------------------
// element e of tile t:  warp w owns [w*32*ITEMS, (w+1)*32*ITEMS), item i of lane l is w*32*ITEMS + i*32 + l
__device__ __forceinline__ int tile_elem(int w, int i, int lane) { return w * 32 * kRsItems + i * 32 + lane; }

// keys: FIRST pass = the col array (stride 1), later passes = .x of the {key, pos} records (stride 2)
__global__ void __launch_bounds__(kRsThreads) radix_hist(const int *__restrict__ keys, int key_stride, int nnz, int shift,
                                                         int radix, int num_tiles, int *__restrict__ counts /*[radix][num_tiles]*/) {
  __shared__ int s_hist[kMaxRadix];
  for (int d = threadIdx.x; d < radix; d += kRsThreads) s_hist[d] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kRsTile;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int k[kRsItems];
#pragma unroll
  for (int i = 0; i < kRsItems; i++) {
    const int64_t e = base + tile_elem(w, i, lane);
    k[i] = (e < nnz) ? __ldg(keys + e * key_stride) : -1;
  }
#pragma unroll
  for (int i = 0; i < kRsItems; i++)
    if (k[i] >= 0) atomicAdd(&s_hist[(k[i] >> shift) & (radix - 1)], 1);
  __syncthreads();
  for (int d = threadIdx.x; d < radix; d += kRsThreads) counts[(size_t)d * num_tiles + blockIdx.x] = s_hist[d];
}

struct ScatterArgs {
  int nnz, shift, radix, num_tiles;
  const int *offsets;       // scanned counts [radix][num_tiles]
  // FIRST pass inputs
  const int *col, *rowexp;
  const float *val;         // may be null
  // later passes: records in
  const int2 *a_in, *p_in;  // {key, pos}, {row, value bits}
  // not LAST: records out
  int2 *a_out, *p_out;
  // LAST: final arrays (any may be null)
  int *perm, *row, *skey;
  float *val_t;
};

constexpr size_t kScatterSmem = sizeof(int) * (kRsWarps * kMaxRadix + 2 * kMaxRadix) + sizeof(int2) * kRsTile;

template <bool FIRST, bool LAST>
__global__ void __launch_bounds__(kRsThreads) radix_scatter(const ScatterArgs a) {
  extern __shared__ __align__(16) unsigned char rs_smem[];
  int *s_cnt = reinterpret_cast<int *>(rs_smem);                  // [kRsWarps][radix]
  int *s_dstart = s_cnt + kRsWarps * kMaxRadix;                   // [radix] first local slot of a digit
  int *s_gbase = s_dstart + kMaxRadix;                            // [radix] global slot of local slot 0 of a digit
  int2 *s_stage = reinterpret_cast<int2 *>(s_gbase + kMaxRadix);  // [kRsTile]
  __shared__ int s_warp[33];
  const int R = a.radix, mask = R - 1;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kRsWarps * R; i += kRsThreads) s_cnt[(i / R) * kMaxRadix + (i % R)] = 0;
  __syncthreads();

  const int64_t base = (int64_t)blockIdx.x * kRsTile;
  const int n_valid = (int)min((int64_t)kRsTile, (int64_t)a.nnz - base);
  int key[kRsItems], pos[kRsItems], slot[kRsItems];
  int2 pay[kRsItems];   // {row, value bits}: requested now, consumed after the keys have been written out
#pragma unroll
  for (int i = 0; i < kRsItems; i++) {
    const int64_t e = base + tile_elem(w, i, lane);
    if (e < a.nnz) {
      if (FIRST) { key[i] = __ldcs(a.col + e); pos[i] = (int)e; }
      else { const int2 t = __ldcs(a.a_in + e); key[i] = t.x; pos[i] = t.y; }
    } else { key[i] = -1; pos[i] = 0; }
  }
#pragma unroll
  for (int i = 0; i < kRsItems; i++) {
    const int64_t e = base + tile_elem(w, i, lane);
    pay[i] = make_int2(0, 0);
    if (e < a.nnz) {
      if (FIRST) { pay[i].x = __ldcs(a.rowexp + e); pay[i].y = a.val ? __float_as_int(__ldcs(a.val + e)) : 0; }
      else pay[i] = __ldcs(a.p_in + e);
    }
  }
  // Tile-stable rank of every element inside its (warp, digit) group.  The lanes holding the same digit are found with
  // one ballot per digit bit (independent, pipelined: ~30 instructions per item) — __match_any_sync costs ~40 cycles per
  // DISTINCT value in the warp, i.e. ~1300 cycles per item with 9-bit digits, and was 35 % of this kernel.  The group
  // leader then bumps the (warp, digit) counter with one shared-memory atomic per item; the atomics of successive items
  // are ordered by __syncwarp, so earlier items (lower CSR positions) always get the lower slots.
  const unsigned lt = (1u << lane) - 1u;
  unsigned peers[kRsItems];
#pragma unroll
  for (int i = 0; i < kRsItems; i++) {
    const bool ok = key[i] >= 0;
    const int d = (key[i] >> a.shift) & mask;
    const unsigned valid = __ballot_sync(0xffffffffu, ok);
    unsigned p = ok ? valid : ~valid;
#pragma unroll
    for (int b = 0; b < kMaxDigitBits; b++) {
      const bool bit = (d >> b) & 1;
      const unsigned bal = __ballot_sync(0xffffffffu, bit);
      p &= bit ? bal : ~bal;
    }
    peers[i] = p;
  }
#pragma unroll
  for (int i = 0; i < kRsItems; i++) {
    const int leader = __ffs(peers[i]) - 1;
    int prev = 0;
    if (key[i] >= 0 && lane == leader) prev = atomicAdd(&s_cnt[w * kMaxRadix + ((key[i] >> a.shift) & mask)], __popc(peers[i]));
    slot[i] = prev;
    __syncwarp();
  }
#pragma unroll
  for (int i = 0; i < kRsItems; i++)
    slot[i] = __shfl_sync(0xffffffffu, slot[i], __ffs(peers[i]) - 1) + __popc(peers[i] & lt);
  __syncthreads();
  // per digit: exclusive scan over the warps; tile totals -> exclusive scan over the digits (2 per thread)
  int tot[2];
#pragma unroll
  for (int j = 0; j < 2; j++) {
    const int d = threadIdx.x * 2 + j;
    int run = 0;
    if (d < R) {
#pragma unroll
      for (int ww = 0; ww < kRsWarps; ww++) {
        const int t = s_cnt[ww * kMaxRadix + d];
        s_cnt[ww * kMaxRadix + d] = run;
        run += t;
      }
    }
    tot[j] = run;
  }
  int total;
  const int ex = block_exclusive_scan(tot[0] + tot[1], &total, s_warp);
#pragma unroll
  for (int j = 0; j < 2; j++) {
    const int d = threadIdx.x * 2 + j;
    if (d < R) {
      const int start = ex + (j ? tot[0] : 0);
      s_dstart[d] = start;
      s_gbase[d] = __ldg(a.offsets + (size_t)d * a.num_tiles + blockIdx.x) - start;
    }
  }
  __syncthreads();
  // sort the tile in shared memory, then write runs
#pragma unroll
  for (int i = 0; i < kRsItems; i++) {
    if (key[i] >= 0) {
      const int d = (key[i] >> a.shift) & mask;
      slot[i] += s_dstart[d] + s_cnt[w * kMaxRadix + d];
      s_stage[slot[i]] = make_int2(key[i], pos[i]);
    }
  }
  __syncthreads();
  int dst[kRsItems];
#pragma unroll
  for (int i = 0; i < kRsItems; i++) {
    const int idx = threadIdx.x + i * kRsThreads;
    dst[i] = -1;
    if (idx < n_valid) {
      const int2 t = s_stage[idx];
      dst[i] = s_gbase[(t.x >> a.shift) & mask] + idx;
      if (!LAST) a.a_out[dst[i]] = t;
      else {
        if (a.perm) a.perm[dst[i]] = t.y;
        if (a.skey) a.skey[dst[i]] = t.x;
      }
    }
  }
  __syncthreads();
  // the payload {row, value} takes the same route
#pragma unroll
  for (int i = 0; i < kRsItems; i++)
    if (key[i] >= 0) s_stage[slot[i]] = pay[i];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kRsItems; i++) {
    if (dst[i] >= 0) {
      const int2 t = s_stage[threadIdx.x + i * kRsThreads];
      if (!LAST) a.p_out[dst[i]] = t;
      else {
        if (a.row) a.row[dst[i]] = t.x;
        if (a.val_t) a.val_t[dst[i]] = __int_as_float(t.y);
      }
    }
  }
}

// colptr from the sorted column stream: position q opens every column in (skey[q-1], skey[q]]; the tail gets nnz
__global__ void __launch_bounds__(256) colptr_from_sorted(const int *__restrict__ skey, int nnz, int ncols, int *__restrict__ colptr) {
  const int stride = gridDim.x * blockDim.x;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q <= nnz; q += stride) {
    const int prev = q > 0 ? min(__ldg(skey + q - 1), ncols - 1) : -1;
    const int cur = q < nnz ? min(__ldg(skey + q), ncols - 1) : ncols;
    for (int c = prev + 1; c <= cur; c++) colptr[c] = q;
  }
}

static inline size_t up256(size_t x) { return (x + 255) / 256 * 256; }

static int key_bits(int ncols) {
  int b = 1;
  while (b < 31 && (1 << b) < ncols) b++;
  return b;
}

static void pass_plan(int ncols, int *passes, int *digit_bits) {
  const int bits = key_bits(ncols);
  *passes = (bits + kMaxDigitBits - 1) / kMaxDigitBits;
  *digit_bits = (bits + *passes - 1) / *passes;
}

size_t csr2csc_workspace_bytes(int M, int ncols, int64_t nnz) {
  (void)M;
  if (nnz < 0) nnz = 0;
  int passes, db;
  pass_plan(ncols > 0 ? ncols : 1, &passes, &db);
  const size_t tiles = (size_t)((nnz + kRsTile - 1) / kRsTile);
  const size_t counts = ((size_t)1 << db) * tiles;
  size_t b = 0;
  b += up256(8 * (size_t)nnz) * 4;                     // {key,pos} and {row,value} record buffers, ping + pong
  b += up256(4 * (size_t)nnz);                         // expanded rows; later the sorted column stream
  b += up256(4 * counts);                              // per-tile digit counts
  b += up256(4 * scan_scratch_ints((int64_t)counts));
  return b + 256;
}

cudaError_t csr2csc(int M, int ncols, int64_t nnz64, const int *rowptr, const int *col, const float *val, int *colptr,
                    int *row, float *val_t, int *perm, void *workspace, size_t workspace_bytes, cudaStream_t s) {
  if (M < 0 || ncols < 0 || nnz64 < 0 || nnz64 > 0x7fffffff) return cudaErrorInvalidValue;
  const int nnz = (int)nnz64;
  if (workspace == nullptr || workspace_bytes < csr2csc_workspace_bytes(M, ncols, nnz64)) return cudaErrorInvalidValue;
  const int sms = device_sm_count();
  cudaError_t e;
  if (nnz == 0 || ncols == 0) return cudaMemsetAsync(colptr, 0, sizeof(int) * ((size_t)ncols + 1), s);

  int passes, db;
  pass_plan(ncols, &passes, &db);
  const int radix = 1 << db;
  char *w = static_cast<char *>(workspace);
  const size_t rec = up256(8 * (size_t)nnz);
  int2 *aA = reinterpret_cast<int2 *>(w), *pA = reinterpret_cast<int2 *>(w + rec);
  int2 *aB = reinterpret_cast<int2 *>(w + 2 * rec), *pB = reinterpret_cast<int2 *>(w + 3 * rec);
  int *rowexp = reinterpret_cast<int *>(w + 4 * rec);          // dead after the first pass ...
  int *skey = rowexp;                                           // ... the last pass (never the first when passes > 1) reuses it
  const int tiles = (nnz + kRsTile - 1) / kRsTile;
  const size_t counts_n = (size_t)radix * tiles;
  int *counts = reinterpret_cast<int *>(w + 4 * rec + up256(4 * (size_t)nnz));
  int *scratch = reinterpret_cast<int *>(w + 4 * rec + up256(4 * (size_t)nnz) + up256(4 * counts_n));
  if (passes == 1) skey = reinterpret_cast<int *>(aA);          // single pass: rowexp is read while skey is written

  {
    const int blocks = (int)std::min<int64_t>(((int64_t)M * 32 + 255) / 256, (int64_t)sms * 16);
    expand_rows<<<blocks > 0 ? blocks : 1, 256, 0, s>>>(rowptr, M, rowexp);
  }
  ScatterArgs a;
  a.nnz = nnz; a.radix = radix; a.num_tiles = tiles; a.offsets = counts;
  a.col = col; a.rowexp = rowexp; a.val = val;
  a.perm = perm; a.row = row; a.skey = skey; a.val_t = (val != nullptr) ? val_t : nullptr;
  a.a_in = nullptr; a.p_in = nullptr; a.a_out = aA; a.p_out = pA;
  for (int pass = 0; pass < passes; pass++) {
    a.shift = pass * db;
    const bool first = pass == 0, last = pass == passes - 1;
    if (first) radix_hist<<<tiles, kRsThreads, 0, s>>>(col, 1, nnz, a.shift, radix, tiles, counts);
    else radix_hist<<<tiles, kRsThreads, 0, s>>>(reinterpret_cast<const int *>(a.a_in), 2, nnz, a.shift, radix, tiles, counts);
    if ((e = exclusive_scan_inplace(counts, (int64_t)counts_n, scratch, s)) != cudaSuccess) return e;
#define DGS_SCATTER(F_, L_)                                                                                               \
  do {                                                                                                                    \
    if ((e = cudaFuncSetAttribute(radix_scatter<F_, L_>, cudaFuncAttributeMaxDynamicSharedMemorySize,                     \
                                  (int)kScatterSmem)) != cudaSuccess) return e;                                          \
    radix_scatter<F_, L_><<<tiles, kRsThreads, kScatterSmem, s>>>(a);                                                     \
  } while (0)
    if (first && last) DGS_SCATTER(true, true);
    else if (first) DGS_SCATTER(true, false);
    else if (last) DGS_SCATTER(false, true);
    else DGS_SCATTER(false, false);
#undef DGS_SCATTER
    a.a_in = a.a_out; a.p_in = a.p_out;
    if (a.a_out == aA) { a.a_out = aB; a.p_out = pB; } else { a.a_out = aA; a.p_out = pA; }
  }
  {
    const int blocks = (int)std::min<int64_t>(((int64_t)nnz + 256) / 256, (int64_t)sms * 32);
    colptr_from_sorted<<<blocks, 256, 0, s>>>(skey, nnz, ncols, colptr);
  }
  return cudaGetLastError();
}

}  // namespace dgs
