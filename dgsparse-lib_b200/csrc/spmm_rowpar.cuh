// spmm_rowpar.cuh — single-launch row-parallel SpMM for the latency regime (10^4 .. 10^6 nonzeros, short rows).
//
// The row-segment scheme (spmm_rowseg.cuh) costs two launches and a binary search per segment, which is most of a
// 15 - 20 us call on matrices like example/data/p2p-Gnutella31.mtx (148 k nnz, 74 % empty rows, longest row 78) — the one
// published configuration of the reference (example/README.md:47-60), where its one-kernel thread-per-element path
// (include/cuda/spmm_cuda.cuh:10-55) was ahead.  Here a lane group owns a ROW: no search, no partials, no fix-up, the
// result (and the 0 / -1 of an empty row) is written directly; same arithmetic and nnz order as the segment kernel.
// A row is walked serially by its group, so this kernel is only chosen when the matrix is known to have no long rows
// (spmm.cu, graph notes); it is correct for any matrix and reports rows longer than kRowParLimit through a.hub_flag so that
// a stale note heals itself.
#pragma once
#include "spmm_rowseg.cuh"

namespace dgs {

constexpr int kRowParLimit = 192;   // longest row (nnz) the row-parallel kernel is chosen for
constexpr int kRowParU = 4;         // gathers in flight per lane and sub-round (a round covers G nonzeros)

template <int VEC, int G, int RED, int COMP, bool ARG>
__global__ void __launch_bounds__(kSpmmThreads) spmm_rowpar_kernel(const SpmmArgs a) {
  constexpr int GPB = kSpmmThreads / G, U = (G < 8) ? G : 8;
  static_assert(G % U == 0, "a round is a whole number of sub-rounds");
  constexpr bool HAS_VAL = (COMP != C_COPY);
  const int grp = threadIdx.x / G, gl = threadIdx.x % G;
  const int r = blockIdx.x * GPB + grp;
  if (r >= a.M) return;
  const int colbase = blockIdx.y * (G * VEC) + gl * VEC;
  const bool active = colbase < a.N;     // lanes beyond a ragged N stay for the shuffles, gather panel 0 and never store
  const int s = __ldg(a.rowptr + r), e = __ldg(a.rowptr + r + 1);
  if (e - s > kRowParLimit && gl == 0 && blockIdx.y == 0 && a.hub_flag) *a.hub_flag = 1;
  const char *__restrict__ Bp = reinterpret_cast<const char *>(a.B + (active ? colbase : 0));
  const unsigned ldb_bytes = (unsigned)(a.ldb * 4);

  float acc[VEC];
  int arg[VEC];
#pragma unroll
  for (int v = 0; v < VEC; v++) { acc[v] = reduce_identity<RED>(); arg[v] = -1; }
  // G nonzeros per round: lane gl fetches (col, val) of nonzero p + gl (one coalesced request per group instead of G
  // identical ones), the group shares them with shuffles and keeps up to U row gathers in flight per lane
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << (G & 31)) - 1u) << (((threadIdx.x & 31) / G) * G));
  for (int p = s; p < e; p += G) {
    const int q = min(p + gl, e - 1);      // past the row end: re-read its last nonzero (in bounds), never accumulated
    const int myc = __ldg(a.col + q);
    const float myv = HAS_VAL ? __ldg(a.val + q) : 1.0f;
    const int n = min(G, e - p);
#pragma unroll
    for (int u0 = 0; u0 < G; u0 += U) {
      if (u0 < n) {                        // uniform within the group
        int cc[U];
        float ev[U], b[U][VEC];
#pragma unroll
        for (int u = 0; u < U; u++) {
          cc[u] = __shfl_sync(gmask, myc, u0 + u, G);
          ev[u] = __shfl_sync(gmask, myv, u0 + u, G);
          ld_vec<VEC>(b[u], reinterpret_cast<const float *>(row_addr(Bp, (unsigned)cc[u], ldb_bytes)));
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
          if (u0 + u < n) {
            float x[VEC];
#pragma unroll
            for (int v = 0; v < VEC; v++) x[v] = compute_op<COMP>(ev[u], b[u][v]);
            reduce_step<RED, ARG, VEC>(acc, arg, x, cc[u]);
          }
        }
      }
    }
  }
  float o[VEC];
  if (e == s) {                          // empty row: 0 and E = -1 (include/cuda/spmm_cuda.cuh:49-53)
#pragma unroll
    for (int v = 0; v < VEC; v++) { o[v] = 0.0f; arg[v] = -1; }
  } else {
    const float deg = (float)(e - s);
#pragma unroll
    for (int v = 0; v < VEC; v++) o[v] = a.mean ? acc[v] / deg : acc[v];
  }
  if (!active) return;
  const size_t off = (size_t)r * a.ldc + colbase;
  if (a.mcast) st_vec_multimem<VEC>(a.dst[0] + off, o);
  else {
    st_vec_cs<VEC>(a.dst[0] + off, o);
    for (int d = 1; d < a.n_dst; d++) st_vec_cs<VEC>(a.dst[d] + off, o);
  }
  if (ARG) st_vec_cs<VEC>(a.E + (size_t)r * a.lde + colbase, arg);
}

template <int VEC, int G, int RED, int COMP, bool ARG>
cudaError_t launch_spmm_rowpar(const SpmmArgs &a, dim3 grid, cudaStream_t s) {
  spmm_rowpar_kernel<VEC, G, RED, COMP, ARG><<<grid, kSpmmThreads, 0, s>>>(a);
  return cudaGetLastError();
}

}  // namespace dgs
