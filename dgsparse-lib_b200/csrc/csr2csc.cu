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
//   scatter: tile-stable ranking (one ballot per digit bit + per-warp digit counters), then the tile is SORTED IN SHARED
//            MEMORY and written out as runs — consecutive threads store consecutive addresses of one bucket, 8 B per
//            element — instead of one random 4 B store per element and array (the first version: 3.7 ms per pass
//            and a 3.1 ms gather at the end on the reddit-like matrix, 13.4 ms in all)
//   colptr:  the last scatter pass lowers colptr[c] (preset to nnz) to the first slot it writes for column c with atomicMin —
//            a minimum is the same whatever the order — and a suffix minimum then gives the empty columns their successor's
//            start (this replaced a 4 B / nnz sorted-key stream written by the last pass and re-read by a colptr kernel:
//            reddit-like 3.31 -> 3.14 ms)
// No atomic decides where a record goes, so every run gives identical bytes.  All traffic is streaming:
// 12 + 16 B per nnz in the first pass, 16 + 16 between passes, 16 + 12 in the last, + 4 / 8 (hist) per pass.
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
constexpr int kScanItems = 8;   // load8 / store8 move them as two int4
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

// eight consecutive int32 of one thread (32-byte aligned: tiles start at multiples of kScanTile, the buffers at 256 bytes):
// two 16-byte accesses when all eight exist
__device__ __forceinline__ void load8(const int *__restrict__ a, int64_t base, int64_t n, int fill, int (&v)[kScanItems]) {
  if (base + kScanItems <= n) {
    const int4 lo = *reinterpret_cast<const int4 *>(a + base), hi = *reinterpret_cast<const int4 *>(a + base + 4);
    v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
  } else {
#pragma unroll
    for (int i = 0; i < kScanItems; i++) v[i] = (base + i < n) ? a[base + i] : fill;
  }
}
__device__ __forceinline__ void store8(int *__restrict__ a, int64_t base, int64_t n, const int (&v)[kScanItems]) {
  if (base + kScanItems <= n) {
    *reinterpret_cast<int4 *>(a + base) = make_int4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<int4 *>(a + base + 4) = make_int4(v[4], v[5], v[6], v[7]);
  } else {
#pragma unroll
    for (int i = 0; i < kScanItems; i++) if (base + i < n) a[base + i] = v[i];
  }
}

// phase 1: per-tile sums
__global__ void __launch_bounds__(kScanThreads) scan_tile_sums(const int *__restrict__ a, int64_t n, int *__restrict__ sums) {
  __shared__ int s_warp[33];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int v[kScanItems];
  load8(a, base, n, 0, v);
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; i++) s += v[i];
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
  load8(a, base, n, 0, v);
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; i++) s += v[i];
  int total;
  int run = block_exclusive_scan(s, &total, s_warp) + sums[blockIdx.x];
#pragma unroll
  for (int i = 0; i < kScanItems; i++) {
    const int t = v[i];
    v[i] = run;
    run += t;
  }
  store8(a, base, n, v);
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
// (also presets colptr[0 .. ncols] to nnz for the last scatter pass's atomicMin: one launch less on small matrices)
__global__ void __launch_bounds__(256) expand_rows(const int *__restrict__ rowptr, int M, int *__restrict__ row,
                                                   int *__restrict__ colptr, int ncols, int nnz) {
  const int lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c <= ncols; c += gridDim.x * blockDim.x) colptr[c] = nnz;
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
  // LAST: final arrays (perm / row / val_t may be null); colptr[c] is lowered to the first slot holding column c
  int *perm, *row, *colptr;
  float *val_t;
  int ncols;
};

// The scatter pass.  Tile-stable rank of every element inside its (warp, digit) group: the lanes holding the same digit are
// found with one ballot per digit bit; the group leader bumps the warp's digit counter; items go in order, so earlier items
// (lower CSR positions) always get the lower slots.  Then the tile is sorted in shared memory and written out as runs.
// What ncu said about the first version (reddit-like, 1.19 + 1.27 ms per pass, 127 registers = 2 CTAs / SM): ALU pipe 64 % busy,
// 193 instructions per element-row of a warp, issue slots 49 % busy — instruction bound with too few warps to fill the gaps.
// This layout runs 3 CTAs / SM (80 registers, 52 KB of shared memory) and executes 17 % fewer instructions: 0.96 + 1.09 ms.
//   * full tiles (all but the last) run a body without any validity test; loads go through per-thread base pointers with
//     immediate offsets
//   * ranking and counting fused per item (no peer masks kept); the ballots of item i + 1 are issued before the shared-memory
//     round trip of item i is waited for; 16-bit (warp, digit) counters updated by the group leader with a plain
//     read-modify-write (one leader per digit and item, items ordered by __syncwarp)
//   * the {row, value} payload is fetched only after the keys have gone out (its 32 registers are not live during ranking)
//   * the destination is recomputed in the payload phase from the digit each thread parked in shared memory (s_dig)
// Measured and slower: 4 CTAs / SM at 64 registers (32 - 56 bytes of spill: 1.17 + 1.37 ms); __match_any_sync for the peers
// (~40 cycles per DISTINCT value in the warp, ~1300 cycles per item with 9-bit digits); the peers through a per-warp
// {lane mask, count} table in shared memory (atomicOr, 8-byte read-back, leader write-back: a third of the ranking
// instructions but three dependent shared-memory round trips per item, 1.33 ms).
constexpr size_t kScatterSmem =
    sizeof(unsigned short) * (kRsWarps * kMaxRadix) + sizeof(int) * (2 * kMaxRadix) + sizeof(int2) * kRsTile + sizeof(unsigned short) * kRsTile;

// lanes of the warp whose digit equals this lane's: one ballot per digit bit
__device__ __forceinline__ unsigned digit_peers(int d, unsigned start) {
  unsigned p = start;
#pragma unroll
  for (int b = 0; b < kMaxDigitBits; b++) {
    const bool bit = (d >> b) & 1;
    const unsigned bal = __ballot_sync(0xffffffffu, bit);
    p &= bit ? bal : ~bal;
  }
  return p;
}

template <bool FIRST, bool LAST, bool FULL>
__device__ __forceinline__ void radix_scatter_body(const ScatterArgs &a, unsigned char *rs_smem, int *s_warp) {
  int2 *s_stage = reinterpret_cast<int2 *>(rs_smem);                                 // [kRsTile]
  int *s_dstart = reinterpret_cast<int *>(s_stage + kRsTile);                       // [radix] first local slot of a digit
  int *s_gbase = s_dstart + kMaxRadix;                                               // [radix] global slot of local slot 0
  unsigned short *s_cnt = reinterpret_cast<unsigned short *>(s_gbase + kMaxRadix);   // [kRsWarps][radix]
  unsigned short *s_dig = s_cnt + kRsWarps * kMaxRadix;                              // [kRsTile] digit of sorted slot idx
  const int R = a.radix, mask = R - 1, shift = a.shift;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kRsWarps * kMaxRadix / 2; i += kRsThreads) reinterpret_cast<unsigned *>(s_cnt)[i] = 0u;
  __syncthreads();

  const int64_t base = (int64_t)blockIdx.x * kRsTile;
  const int n_valid = FULL ? kRsTile : (int)((int64_t)a.nnz - base);
  const int first = w * 32 * kRsItems + lane;      // tile_elem(w, 0, lane); item i is first + 32 i
  const int64_t e0 = base + first;
  int key[kRsItems], pos[kRsItems], slot[kRsItems];
  {
    const int *pc = a.col + e0;
    const int2 *pa = a.a_in + e0;
#pragma unroll
    for (int i = 0; i < kRsItems; i++) {
      pos[i] = 0;
      if (FULL || first + 32 * i < n_valid) {
        if (FIRST) key[i] = __ldcs(pc + 32 * i);
        else { const int2 t = __ldcs(pa + 32 * i); key[i] = t.x; pos[i] = t.y; }
      } else key[i] = -1;
    }
  }
  const unsigned lt = (1u << lane) - 1u;
  unsigned short *cnt = s_cnt + w * kMaxRadix;
  auto peers_of = [&](int i) -> unsigned {
    if (FULL) return digit_peers((key[i] >> shift) & mask, 0xffffffffu);
    const bool ok = key[i] >= 0;
    const unsigned valid = __ballot_sync(0xffffffffu, ok);
    return digit_peers((key[i] >> shift) & mask, ok ? valid : ~valid);
  };
  unsigned p_next = peers_of(0);
#pragma unroll
  for (int i = 0; i < kRsItems; i++) {
    const unsigned p = p_next;
    const int d = (key[i] >> shift) & mask;
    const bool leader = (p & lt) == 0u && (FULL || key[i] >= 0);
    int prev = 0;
    if (leader) { prev = cnt[d]; cnt[d] = (unsigned short)(prev + __popc(p)); }
    if (i + 1 < kRsItems) p_next = peers_of(i + 1);
    __syncwarp();
    slot[i] = __shfl_sync(0xffffffffu, prev, __ffs(p) - 1) + __popc(p & lt);
  }
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
        s_cnt[ww * kMaxRadix + d] = (unsigned short)run;
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
    if (FULL || key[i] >= 0) {
      const int d = (key[i] >> shift) & mask;
      slot[i] += s_dstart[d] + cnt[d];
      s_stage[slot[i]] = make_int2(key[i], FIRST ? (int)e0 + 32 * i : pos[i]);
    } else slot[i] = -1;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kRsItems; i++) {
    const int idx = threadIdx.x + i * kRsThreads;
    if (FULL || idx < n_valid) {
      const int2 t = s_stage[idx];
      const int dg = (t.x >> shift) & mask;
      const int dst = s_gbase[dg] + idx;
      s_dig[idx] = (unsigned short)dg;   // read back by this same thread in the payload phase
      if (!LAST) a.a_out[dst] = t;
      else {
        if (a.perm) a.perm[dst] = t.y;
        // A column opens where the key changes.  Inside the tile the predecessor in shared memory is the predecessor in the
        // output (same key = same digit = same run); the first element of a run may continue the previous tile's column,
        // which then holds the smaller slot: atomicMin keeps the first one whatever the order.
        if (idx == 0 || s_stage[idx - 1].x != t.x) atomicMin(a.colptr + min(t.x, a.ncols - 1), dst);
      }
    }
  }
  // the payload {row, value} takes the same route
  int2 pay[kRsItems];
  {
    const int *pr = a.rowexp + e0;
    const float *pv = a.val ? a.val + e0 : nullptr;
    const int2 *pp = a.p_in + e0;
    const bool has_val = a.val != nullptr;
#pragma unroll
    for (int i = 0; i < kRsItems; i++) {
      pay[i] = make_int2(0, 0);
      if (FULL || slot[i] >= 0) {
        if (FIRST) { pay[i].x = __ldcs(pr + 32 * i); if (has_val) pay[i].y = __float_as_int(__ldcs(pv + 32 * i)); }
        else pay[i] = __ldcs(pp + 32 * i);
      }
    }
  }
  __syncthreads();   // every key has been read out of the stage
#pragma unroll
  for (int i = 0; i < kRsItems; i++)
    if (FULL || slot[i] >= 0) s_stage[slot[i]] = pay[i];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kRsItems; i++) {
    const int idx = threadIdx.x + i * kRsThreads;
    if (FULL || idx < n_valid) {
      const int2 t = s_stage[idx];
      const int dst = s_gbase[s_dig[idx]] + idx;
      if (!LAST) a.p_out[dst] = t;
      else {
        if (a.row) a.row[dst] = t.x;
        if (a.val_t) a.val_t[dst] = __int_as_float(t.y);
      }
    }
  }
}

template <bool FIRST, bool LAST>
__global__ void __launch_bounds__(kRsThreads, 3) radix_scatter(const ScatterArgs a) {
  extern __shared__ __align__(16) unsigned char rs_smem[];
  __shared__ int s_warp[33];
  if ((int64_t)(blockIdx.x + 1) * kRsTile <= (int64_t)a.nnz) radix_scatter_body<FIRST, LAST, true>(a, rs_smem, s_warp);
  else radix_scatter_body<FIRST, LAST, false>(a, rs_smem, s_warp);
}

// colptr: preset to nnz (expand_rows), lowered by the last scatter pass to the first slot of every non-empty column; an empty
// column then takes the value of the next non-empty one = an inclusive suffix minimum over colptr[0 .. ncols] (three small
// kernels).
// exclusive suffix minimum over the threads of a block (identity INT_MAX); *total = the block's minimum
__device__ __forceinline__ int block_exclusive_suffix_min(int x, int *total, int *s_warp) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  int inc = x;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_down_sync(0xffffffffu, inc, o);
    if (lane + o < 32) inc = min(inc, t);
  }
  int ex = __shfl_down_sync(0xffffffffu, inc, 1);
  if (lane == 31) ex = 0x7fffffff;
  if (lane == 0) s_warp[w] = inc;
  __syncthreads();
  int after = 0x7fffffff, all = 0x7fffffff;
  for (int ww = 0; ww < nw; ww++) {
    const int t = s_warp[ww];
    all = min(all, t);
    if (ww > w) after = min(after, t);
  }
  *total = all;
  __syncthreads();
  return min(ex, after);
}

__global__ void __launch_bounds__(kScanThreads) sufmin_tile_mins(const int *__restrict__ a, int64_t n, int *__restrict__ mins) {
  __shared__ int s_warp[32];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int m = 0x7fffffff;
#pragma unroll
  for (int i = 0; i < kScanItems; i++) if (base + i < n) m = min(m, a[base + i]);
  int total;
  block_exclusive_suffix_min(m, &total, s_warp);
  if (threadIdx.x == 0) mins[blockIdx.x] = total;
}

// one block: mins[t] <- minimum over the tiles after t, walking backwards with a carry
__global__ void __launch_bounds__(1024) sufmin_tiles_inplace(int *__restrict__ mins, int n) {
  __shared__ int s_warp[32];
  int carry = 0x7fffffff;
  for (int hi = n; hi > 0; hi -= 1024) {
    const int i = hi - 1024 + (int)threadIdx.x;
    const int x = i >= 0 ? mins[i] : 0x7fffffff;
    int total;
    const int ex = block_exclusive_suffix_min(x, &total, s_warp);
    if (i >= 0) mins[i] = min(carry, ex);
    carry = min(carry, total);
  }
}

__global__ void __launch_bounds__(kScanThreads) sufmin_apply(int *__restrict__ a, int64_t n, const int *__restrict__ mins) {
  __shared__ int s_warp[32];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int v[kScanItems];
  int m = 0x7fffffff;
#pragma unroll
  for (int i = 0; i < kScanItems; i++) { v[i] = (base + i < n) ? a[base + i] : 0x7fffffff; m = min(m, v[i]); }
  int total;
  int run = min(block_exclusive_suffix_min(m, &total, s_warp), mins[blockIdx.x]);
#pragma unroll
  for (int i = kScanItems - 1; i >= 0; i--) {
    run = min(run, v[i]);
    if (base + i < n) a[base + i] = run;
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
  b += up256(4 * (size_t)nnz);                         // expanded rows
  b += up256(4 * counts);                              // per-tile digit counts
  b += up256(4 * std::max(scan_scratch_ints((int64_t)counts), scan_scratch_ints((int64_t)ncols + 1)));   // scan / suffix-min tiles
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
  int *rowexp = reinterpret_cast<int *>(w + 4 * rec);
  const int tiles = (nnz + kRsTile - 1) / kRsTile;
  const size_t counts_n = (size_t)radix * tiles;
  int *counts = reinterpret_cast<int *>(w + 4 * rec + up256(4 * (size_t)nnz));
  int *scratch = reinterpret_cast<int *>(w + 4 * rec + up256(4 * (size_t)nnz) + up256(4 * counts_n));

  {
    const int blocks = (int)std::min<int64_t>(((int64_t)M * 32 + 255) / 256, (int64_t)sms * 16);
    expand_rows<<<blocks > 0 ? blocks : 1, 256, 0, s>>>(rowptr, M, rowexp, colptr, ncols, nnz);
  }
  ScatterArgs a;
  a.nnz = nnz; a.radix = radix; a.num_tiles = tiles; a.offsets = counts;
  a.col = col; a.rowexp = rowexp; a.val = val;
  a.perm = perm; a.row = row; a.colptr = colptr; a.ncols = ncols; a.val_t = (val != nullptr) ? val_t : nullptr;
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
    // the scan scratch is free again: counts_n >= ncols / tile-size is not guaranteed, so size it for ncols + 1 here
    const int64_t n = (int64_t)ncols + 1;
    const int t = (int)((n + kScanTile - 1) / kScanTile);
    sufmin_tile_mins<<<t, kScanThreads, 0, s>>>(colptr, n, scratch);
    sufmin_tiles_inplace<<<1, 1024, 0, s>>>(scratch, t);
    sufmin_apply<<<t, kScanThreads, 0, s>>>(colptr, n, scratch);
  }
  return cudaGetLastError();
}

}  // namespace dgs
