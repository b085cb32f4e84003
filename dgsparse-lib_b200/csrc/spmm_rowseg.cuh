// spmm_rowseg.cuh — CSR SpMM / generalized SpMM for sm_100a: "each lane group owns a row segment".
//
// Replaces (behaviour, not code) the reference kernels
//   csrspmm_seqreduce_rowbalance_kernel      include/cuda/spmm_cuda.cuh:10-55   (torch face, algorithm 0)
//   csrspmm_rowcaching_rowbalance_kernel     src/ge-spmm/csrspmm_rowcaching.cu:11-119 (C ABI, N > 32)
//   weighted/topo {Simple,Cache,CacheCoarsen}SPMMKernel   src/gspmm-fp/gspmm.cu:6-404
//
// Scheme
//   * The nnz stream [0, nnz) is cut into equal SEGMENTS of `chunk` nonzeros (a multiple of 32).  A
//     segment is owned by a GROUP of G lanes (G*VEC = the dense panel width handled per pass, e.g.
//     N=64 -> 16 lanes x float4, two independent groups per warp; N=128 -> one warp x float4).  Load
//     balance therefore never depends on the degree distribution (hub rows are simply split).
//   * The group stages 32 (col, val) pairs at a time through shared memory (coalesced streaming
//     loads, next batch prefetched into registers while the current one is consumed), then walks
//     them: U dense rows B[col, panel] are requested back-to-back as 128-bit loads (one 128 B..512 B
//     coalesced row slice per request) before any of them is consumed, so each group keeps U row
//     gathers in flight.  Each lane owns VEC fixed output columns: accumulation over a row is
//     serial in nnz order, exactly like the reference, and needs no cross-lane reduction.
//   * Rows that end inside the segment are written straight to C.  A row cut by a segment boundary
//     leaves a raw partial accumulator (head and/or tail slot of the segment) in a small workspace;
//     spmm_fixup_kernel folds the partials of such a row IN SEGMENT ORDER (deterministic, preserves
//     the first-extremum-wins arg rule) and also writes the 0 / -1 rows of empty rows.
//     (Measured alternative: folding inside this kernel — every segment bumps a counter of the cut row's first
//     segment after a __threadfence, the last arrival folds — is correct but slower everywhere: the fence at each
//     segment end waits for the segment's streaming stores, reddit@64 1.568 -> 1.596 ms for 0.006 ms saved in the
//     fix-up, p2p-Gnutella31 27 -> 36 us, arxiv-like N=256 0.236 -> 0.295 ms.)
#pragma once
#include <type_traits>
#include "common.cuh"

namespace dgs {

struct SpmmArgs {
  int M, N, nnz;
  const int *rowptr;
  const int *col;
  const float *val;   // may be null iff COMP == C_COPY
  const float *B;
  int64_t ldb;
  int64_t ldc;        // leading dimension of every destination in dst[]
  int *E;             // arg column index (MAX/MIN), null otherwise
  int64_t lde;
  int chunk;          // nnz per segment, multiple of 32
  int num_chunks;
  int mean;           // divide the finished row by its nnz count
  float *part_val;    // [num_chunks][2][N] raw partial accumulators: slot 0 = head, 1 = tail
  int *part_arg;      // same shape, only with ARG
  int *tail_row;      // [num_chunks] row whose tail partial lives in slot 1, or -1
  int n_dst;          // 1 (local C) or the number of column-shard peers
  int mcast;          // dst[0] is an NVLS multicast address: one multimem.st reaches every rank's C (n_dst == 1)
  float *dst[kMaxDst];
  const int *mask;    // COMP == C_MASK only: arg index tensor E of the forward, [*, ldm]
  int64_t ldm;
  int nnz_dev;        // the true nnz is rowptr[M] ON THE DEVICE; nnz / chunk / num_chunks above come from a host-side hint
  int *nnz_report;    // (with nnz_dev) mapped host word that receives the true nnz: the next call's hint
  int *hub_flag;      // mapped host word set to 1 when a row longer than kRowParLimit is seen (null: nobody asks)
  int hub_limit;
};

// Segment layout actually used by a launch.  Normally the host's (a.nnz, a.chunk, a.num_chunks).  The legacy entry points
// (spmm_cuda(m, k, rowptr, ...): no nnz argument) must not block on a device->host copy of rowptr[M], so there the host
// sizes the grid and the workspace from the nnz it saw on the previous call with this rowptr (a hint) and every kernel
// re-derives the layout from the true nnz: same segments when the hint was right, fewer when nnz shrank, LONGER segments
// (never more than the a.num_chunks the workspace was sized for) when it grew.
struct SegLayout { int nnz, chunk, num_chunks; };
__device__ __forceinline__ SegLayout seg_layout(const SpmmArgs &a) {
  SegLayout s = {a.nnz, a.chunk, a.num_chunks};
  if (a.nnz_dev) {
    s.nnz = __ldg(a.rowptr + a.M);
    const int need = (int)(((int64_t)s.nnz + a.num_chunks - 1) / a.num_chunks);
    if (need > s.chunk) s.chunk = (need + 31) / 32 * 32;
    s.num_chunks = (int)(((int64_t)s.nnz + s.chunk - 1) / s.chunk);
  }
  return s;
}

// Threads per CTA of the SpMM kernels.  Lane groups never talk to each other, so this only sets the granularity in which an SM's
// slots are handed out; -DDGS_SPMM_THREADS=64 builds the variant library tools/build_variant.py A/Bs against this one.
#ifndef DGS_SPMM_THREADS
#define DGS_SPMM_THREADS 256
#endif
constexpr int kSpmmThreads = DGS_SPMM_THREADS;
constexpr int kSpmmCtaScale = 256 / kSpmmThreads;   // CTAs per SM asked for scale with the CTA size
constexpr int kBatch = 32;

template <int RED, bool ARG, int VEC>
__device__ __forceinline__ void reduce_step(float (&acc)[VEC], int (&arg)[VEC], const float (&x)[VEC], int c) {
#pragma unroll
  for (int v = 0; v < VEC; v++) {
    if (RED == R_MAX) {          // include/cuda/spmm_cuda.cuh:38-42 + MAX macro include/gspmm.h:17
      if (ARG) { if (acc[v] < x[v]) arg[v] = c; }
      acc[v] = (acc[v] < x[v]) ? x[v] : acc[v];
    } else if (RED == R_MIN) {   // MIN macro include/gspmm.h:16
      if (ARG) { if (acc[v] > x[v]) arg[v] = c; }
      acc[v] = (acc[v] < x[v]) ? acc[v] : x[v];
    } else {
      acc[v] += x[v];
    }
  }
}

// Shared-memory staging slot: (col, val bits) of one nonzero, read back with one LDS.64 broadcast.
// Row stride 34 entries (272 B): rows of the groups of one warp start 4 banks apart -> conflict-free
// 64-bit broadcasts, and every row stays 16-byte aligned.
constexpr int kStageStride = kBatch + 2;

// Register budget: the sum / max / min flavours of the 8- and 16-lane vec4 geometries compile to 79 - 80 registers (3 CTAs per
// SM); left alone, ptxas gives their arg-tracking flavours 96 - 100, i.e. 2 CTAs per SM, although they fit 80 with 8 - 24 bytes
// of spill.  Ask for 3 CTAs per SM there (the masked flavour would spill 200 bytes, the 4-lane geometry 40: both keep the
// compiler's own choice).  The second argument must be 0 ("unspecified") and NOT 1 for the others: 1 tells ptxas the kernel
// may take the whole register file, it then gave the 8-lane sum kernel 124 registers, and N = 32 ran 12 - 20 % slower at
// 2 CTAs per SM (reddit-like 0.93 against 0.84 ms, products-like 2.81 against 2.31 ms; tools/exp_ab_r1.py).
template <int VEC, int G, int RED, int COMP, bool ARG, int U>
__global__ void __launch_bounds__(kSpmmThreads, (VEC == 4 && G >= 8 && COMP != C_MASK) ? 3 * kSpmmCtaScale : 0) spmm_rowseg_kernel(const SpmmArgs a) {
  constexpr int GPB = kSpmmThreads / G;  // groups (segments) per block
  constexpr int PER = kBatch / G;        // staged entries per lane per batch
  constexpr bool HAS_VAL = (COMP != C_COPY);
  constexpr bool MASKED = (COMP == C_MASK);
  constexpr int MV = MASKED ? VEC : 1;
  static_assert(kBatch % G == 0 && kBatch % U == 0, "bad tiling");

  __shared__ __align__(16) int2 s_cv[HAS_VAL ? GPB : 1][kStageStride];
  __shared__ __align__(16) int s_c[HAS_VAL ? 1 : GPB][kStageStride];

  const int grp = threadIdx.x / G;
  const int gl = threadIdx.x % G;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << (G & 31)) - 1u) << (((threadIdx.x & 31) / G) * G));
  // Programmatic dependent launch: the fix-up grid may start now; its first half (the zero rows of empty rows) touches
  // nothing this kernel reads or writes, its second half waits for this grid (griddepcontrol.wait).
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int chunk_id = blockIdx.x * GPB + grp;
  const SegLayout sl = seg_layout(a);
  if (chunk_id >= sl.num_chunks) return;

  const int lo = chunk_id * sl.chunk;
  const int hi = (sl.nnz - lo <= sl.chunk) ? sl.nnz : lo + sl.chunk;
  const int colbase = blockIdx.y * (G * VEC) + gl * VEC;
  const bool active = colbase < a.N;
  // lanes beyond a ragged N recompute panel 0 (always in bounds) and never store: no predication in the hot loop
  const int ldcol = active ? colbase : 0;
  const char *__restrict__ Bp = reinterpret_cast<const char *>(a.B + ldcol);
  const char *__restrict__ Mp = MASKED ? reinterpret_cast<const char *>(a.mask + ldcol) : nullptr;
  const unsigned ldb_bytes = (unsigned)(a.ldb * 4);   // one IMAD.WIDE.U32 per gather address
  const unsigned ldm_bytes = MASKED ? (unsigned)(a.ldm * 4) : 0u;
  const int *__restrict__ rowptr = a.rowptr;

  int r = row_of_nnz(rowptr, a.M, lo);
  int row_start = __ldg(rowptr + r);
  int row_end = __ldg(rowptr + r + 1);

  float acc[VEC];
  int arg[VEC];
#pragma unroll
  for (int v = 0; v < VEC; v++) { acc[v] = reduce_identity<RED>(); arg[v] = -1; }

  auto store_partial = [&](int slot) {
    const size_t off = ((size_t)chunk_id * 2 + slot) * a.N + colbase;
    if (VEC == 4) {
      *reinterpret_cast<float4 *>(a.part_val + off) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      if (ARG) *reinterpret_cast<int4 *>(a.part_arg + off) = make_int4(arg[0], arg[1], arg[2], arg[3]);
    } else {
#pragma unroll
      for (int v = 0; v < VEC; v++) { a.part_val[off + v] = acc[v]; if (ARG) a.part_arg[off + v] = arg[v]; }
    }
  };

  // row r has been consumed up to row_end (<= hi)
  auto finish_row = [&]() {
    if (active) {
      if (row_start >= lo) {  // the whole row lies inside this segment: final result
        float o[VEC];
        const float deg = (float)(row_end - row_start);
#pragma unroll
        for (int v = 0; v < VEC; v++) o[v] = a.mean ? acc[v] / deg : acc[v];
        const size_t off = (size_t)r * a.ldc + colbase;
        if (a.mcast) st_vec_multimem<VEC>(a.dst[0] + off, o);                   // NVSwitch multicast: all ranks at once
        else {
          st_vec_cs<VEC>(a.dst[0] + off, o);
          for (int d = 1; d < a.n_dst; d++) st_vec_cs<VEC>(a.dst[d] + off, o);   // NVLink peers
        }
        if (ARG) st_vec_cs<VEC>(a.E + (size_t)r * a.lde + colbase, arg);
      } else {
        store_partial(0);     // the row began in an earlier segment: head partial
      }
    }
#pragma unroll
    for (int v = 0; v < VEC; v++) { acc[v] = reduce_identity<RED>(); arg[v] = -1; }
  };

  auto advance_to = [&](int pos) {  // next non-empty row; pos is the first nnz not yet consumed
    r += 1;
    row_start = row_end;
    row_end = __ldg(rowptr + r + 1);
    if (row_end <= pos) {           // run of empty rows: jump (galloping search from here)
      r = row_of_nnz_from(rowptr, a.M, pos, r);
      row_start = __ldg(rowptr + r);
      row_end = __ldg(rowptr + r + 1);
    }
  };

  int creg[PER];
  float vreg[PER];
  auto prefetch = [&](int base) {
#pragma unroll
    for (int k = 0; k < PER; k++) {
      const int idx = base + gl + k * G;
      const bool ok = idx < hi;
      creg[k] = ok ? __ldcs(a.col + idx) : 0;   // col 0 keeps the speculative B load in bounds
      if (HAS_VAL) vreg[k] = (MASKED && a.val == nullptr) ? 1.0f : (ok ? __ldcs(a.val + idx) : 0.0f);
    }
  };
  prefetch(lo);

  // acc (+)= compute(val, B-row) for one nonzero
  auto accumulate = [&](int c, float ev, const float (&bv)[VEC], const int (&mv)[MV]) {
    float x[VEC];
#pragma unroll
    for (int v = 0; v < VEC; v++) {
      x[v] = compute_op<COMP>(ev, bv[v]);
      if (MASKED) x[v] = (mv[v] == r) ? x[v] : 0.0f;   // include/cuda/spmm_cuda.cuh:400-433 (intended semantics)
    }
    reduce_step<RED, ARG, VEC>(acc, arg, x, c);
  };

  auto gather = [&](int c, float (&bv)[VEC], int (&mv)[MV]) {
    ld_vec<VEC>(bv, reinterpret_cast<const float *>(row_addr(Bp, (unsigned)c, ldb_bytes)));
    if (MASKED) ld_ivec<MV>(mv, reinterpret_cast<const int *>(row_addr(Mp, (unsigned)c, ldm_bytes)));
  };

  for (int base = lo; base < hi; base += kBatch) {
#pragma unroll
    for (int k = 0; k < PER; k++) {
      if (HAS_VAL) s_cv[grp][gl + k * G] = make_int2(creg[k], __float_as_int(vreg[k]));
      else s_c[grp][gl + k * G] = creg[k];
    }
    __syncwarp(gmask);
    prefetch(base + kBatch);
    if (hi - base >= kBatch) {
#pragma unroll 1
      for (int j0 = 0; j0 < kBatch; j0 += U) {
        // U row gathers in flight before the first one is consumed
        float b[U][VEC];
        int mk[U][MV];
        int cc[U];
        float ev[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
          if (HAS_VAL) { const int2 t = s_cv[grp][j0 + u]; cc[u] = t.x; ev[u] = __int_as_float(t.y); }
          else { cc[u] = s_c[grp][j0 + u]; ev[u] = 1.0f; }
          gather(cc[u], b[u], mk[u]);
        }
        const int p0 = base + j0;
        if (p0 + U <= row_end) {
          // fast path: the U nonzeros all belong to the current row
#pragma unroll
          for (int u = 0; u < U; u++) accumulate(cc[u], ev[u], b[u], mk[u]);
        } else {
#pragma unroll
          for (int u = 0; u < U; u++) {
            if (p0 + u >= row_end) { finish_row(); advance_to(p0 + u); }
            accumulate(cc[u], ev[u], b[u], mk[u]);
          }
        }
      }
    } else {
      // ragged last batch of the last segment
#pragma unroll 1
      for (int j = 0; j < hi - base; j++) {
        float b[VEC];
        int mk[MV];
        int c;
        float ev = 1.0f;
        if (HAS_VAL) { const int2 t = s_cv[grp][j]; c = t.x; ev = __int_as_float(t.y); }
        else c = s_c[grp][j];
        gather(c, b, mk);
        if (base + j >= row_end) { finish_row(); advance_to(base + j); }
        accumulate(c, ev, b, mk);
      }
    }
    __syncwarp(gmask);
  }

  int tail = -1;
  if (row_end == hi) {
    finish_row();
  } else if (row_start >= lo) {   // row continues into the next segment: tail partial, this group owns the row
    if (active) store_partial(1);
    tail = r;
  } else {                        // row spans the whole segment
    if (active) store_partial(0);
  }
  if (gl == 0 && blockIdx.y == 0) a.tail_row[chunk_id] = tail;
}

// Folds the partials of rows cut by segment boundaries (one thread per (segment, column)), in segment
// order, and writes empty rows (0 and E = -1: include/cuda/spmm_cuda.cuh:49-53, src/gspmm-fp/gspmm.cu:222).
// FV = 4: one thread folds four adjacent columns (16-byte loads of the partials, 16-byte stores to every
// destination — with NVLink peers as destinations the store count is what matters); FV = 1: any N / alignment.
template <int RED, bool ARG, int FV>
__global__ void __launch_bounds__(256) spmm_fixup_kernel(const SpmmArgs a) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // ---- part 1 (independent of the SpMM grid, may overlap it under programmatic dependent launch) ----
  // empty rows: one warp scans 32 rows, then writes the zero rows cooperatively
  {
  const int lane = threadIdx.x & 31;
  const int64_t warp = t >> 5;
  const int64_t row0 = warp * 32;
  if (row0 < a.M) {
    const int row = (int)row0 + lane;
    bool empty = false, hub = false;
    if (row < a.M) {
      const int deg = __ldg(a.rowptr + row + 1) - __ldg(a.rowptr + row);
      empty = deg == 0;
      hub = deg > a.hub_limit;
    }
    if (a.hub_flag && __any_sync(0xffffffffu, hub) && lane == 0) *a.hub_flag = 1;   // not a matrix for the row-parallel kernel
    unsigned m = __ballot_sync(0xffffffffu, empty);
    while (m) {
      const int rr = (int)row0 + (__ffs(m) - 1);
      m &= m - 1;
      for (int c = lane * FV; c < a.N; c += 32 * FV) {   // FV = 4: 16-byte stores (N % 4 == 0, aligned rows)
        float z[FV];
        int m1[FV];
#pragma unroll
        for (int v = 0; v < FV; v++) { z[v] = 0.0f; m1[v] = -1; }
        if (a.mcast) st_vec_multimem<FV>(a.dst[0] + (size_t)rr * a.ldc + c, z);
        else for (int d = 0; d < a.n_dst; d++) st_vec_cs<FV>(a.dst[d] + (size_t)rr * a.ldc + c, z);
        if (ARG) st_vec_cs<FV>(a.E + (size_t)rr * a.lde + c, m1);
      }
    }
  }
  }
  // ---- part 2: the partials of the SpMM grid must be complete and visible ----
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const SegLayout sl = seg_layout(a);
  if (a.nnz_dev && t == 0 && a.nnz_report) *a.nnz_report = sl.nnz;   // the next call's hint (mapped host memory)
  const int nq = a.N / FV;                      // column groups per row
  const int64_t n_fold = (int64_t)sl.num_chunks * nq;
  if (t < n_fold) {
    const int g = (int)(t / nq);
    const int c = (int)(t % nq) * FV;
    const int r = a.tail_row[g];
    if (r >= 0) {
      const int start = __ldg(a.rowptr + r), end = __ldg(a.rowptr + r + 1);
      const int g_last = (end - 1) / sl.chunk;
      float acc[FV];
      int arg[FV];
      ld_vec<FV>(acc, a.part_val + ((size_t)g * 2 + 1) * a.N + c);
      if (ARG) ld_ivec<FV>(arg, a.part_arg + ((size_t)g * 2 + 1) * a.N + c);
      // A hub row spans many segments (arxiv-like, 64-nnz segments: 200): fetch the partials kFoldU at a time (independent
      // loads, 8 or 16 in flight), fold them strictly in segment order.  One at a time this loop was a 50 us dependent-load chain.
      constexpr int kFoldU = ARG ? 8 : 16;
      for (int g0 = g + 1; g0 <= g_last; g0 += kFoldU) {
        float x[kFoldU][FV];
        int xa[kFoldU][FV];
#pragma unroll
        for (int u = 0; u < kFoldU; u++) {
          const int gg = min(g0 + u, g_last);
          ld_vec<FV>(x[u], a.part_val + ((size_t)gg * 2) * a.N + c);
          if (ARG) ld_ivec<FV>(xa[u], a.part_arg + ((size_t)gg * 2) * a.N + c);
        }
#pragma unroll
        for (int u = 0; u < kFoldU; u++) {
          if (g0 + u <= g_last) {
#pragma unroll
            for (int v = 0; v < FV; v++) {
              if (RED == R_MAX) {
                if (ARG) { if (acc[v] < x[u][v]) arg[v] = xa[u][v]; }
                acc[v] = (acc[v] < x[u][v]) ? x[u][v] : acc[v];
              } else if (RED == R_MIN) {
                if (ARG) { if (acc[v] > x[u][v]) arg[v] = xa[u][v]; }
                acc[v] = (acc[v] < x[u][v]) ? acc[v] : x[u][v];
              } else {
                acc[v] += x[u][v];
              }
            }
          }
        }
      }
      if (a.mean) {
        const float deg = (float)(end - start);
#pragma unroll
        for (int v = 0; v < FV; v++) acc[v] = acc[v] / deg;
      }
      const size_t off = (size_t)r * a.ldc + c;
      if (a.mcast) st_vec_multimem<FV>(a.dst[0] + off, acc);
      else for (int d = 0; d < a.n_dst; d++) st_vec_cs<FV>(a.dst[d] + off, acc);
      if (ARG) st_vec_cs<FV>(a.E + (size_t)r * a.lde + c, arg);
    }
  }
}

// host-side handle of one (VEC, G, RED, COMP, ARG) instance (instantiated in spmm_inst.cu): its launcher and how many of its
// CTAs fit on one SM (the register count differs between flavours: 79 - 80 registers = 3 CTAs for sum / max / min, 100 - 128 =
// 2 for the arg-tracking and masked kernels and for the 4-lane geometry).  The dispatcher sizes the segments from it so
// that the grid is a whole number of waves.
using SpmmLaunchFn = cudaError_t (*)(const SpmmArgs &, dim3 grid, cudaStream_t);
struct SpmmKernel {
  SpmmLaunchFn launch = nullptr;
  int (*blocks_per_sm)() = nullptr;
  explicit operator bool() const { return launch != nullptr; }
};

template <int VEC, int G, int RED, int COMP, bool ARG>
cudaError_t launch_spmm_rowseg(const SpmmArgs &a, dim3 grid, cudaStream_t s) {
  constexpr int U = (VEC == 4) ? 8 : 8;
  spmm_rowseg_kernel<VEC, G, RED, COMP, ARG, U><<<grid, kSpmmThreads, 0, s>>>(a);
  return cudaGetLastError();
}

template <int VEC, int G, int RED, int COMP, bool ARG>
int occupancy_spmm_rowseg() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 2;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, spmm_rowseg_kernel<VEC, G, RED, COMP, ARG, 8>, kSpmmThreads, 0) != cudaSuccess || n < 1) {
      cudaGetLastError();
      n = 2;
    }
    cached[dev] = n;
  }
  return cached[dev];
}

template <int VEC, int G, int RED, int COMP, bool ARG>
SpmmKernel spmm_rowseg_handle() {
  SpmmKernel k;
  k.launch = &launch_spmm_rowseg<VEC, G, RED, COMP, ARG>;
  k.blocks_per_sm = &occupancy_spmm_rowseg<VEC, G, RED, COMP, ARG>;
  return k;
}

}  // namespace dgs
