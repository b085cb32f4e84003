// DEAD END (round 2) — not compiled into the product; kept as the source of the experiment in profiles/r02_exp_panels_*.jsonl.
// See pick_panel() in dgsparse-lib_b200/csrc/spmm.cu for the measurements and why it loses.
// spmm_narrow.cuh — the row-segment SpMM for dense operands that do NOT fit the L2: narrow column panels.
//
// Problem (ogbn-products-like, feat 128: B = 2.45 M rows x 512 B = 1.25 GB): with 64-column panels every nonzero pulls
// its 256 B row slice from HBM — 61 GB of DRAM traffic for 3.5 GB of algorithmic bytes (profiles/r01_ncu_full_products128.txt).
// Here the feature axis is cut into panels of W = 4*G columns (G = 2 -> 8 columns = ONE 32 B sector per gathered row,
// G = 4 -> 16 columns) sized by the dispatcher so that K x W x 4 B stays L2-resident, and the panels run one after the
// other (blockIdx.y is the slow grid axis).  Each pass re-streams col / val (evict-first) and gathers B through an
// L2::evict_last policy with no L1 allocation; C is written once, one sector per row and pass.  DRAM traffic becomes
// passes x (col + val) + B + C instead of nnz x row bytes.
//
// Same scheme and same arithmetic as spmm_rowseg_kernel (segments of the nnz stream, a lane group per segment, serial
// accumulation in nnz order, head / tail partials folded by spmm_fixup_kernel), with the staging changed for tiny groups:
// a group of 2 or 4 lanes has nothing to share through shared memory, so every lane reads the group's (col, val) stream
// itself, 8 nonzeros per 2 x 16-byte load (the lanes of a group read the same address = one transaction; a warp
// instruction covers one sector-half of 32 / G segments, the second load of the pair the other half), one block of 8
// prefetched in registers while the current one is consumed.
#pragma once
#include "spmm_rowseg.cuh"

namespace dgs {

__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}

__device__ __forceinline__ void ld_gather4(float (&d)[4], const void *p, uint64_t pol) {
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3]) : "l"(p), "l"(pol));
}

__device__ __forceinline__ int4 ld_stream_i4(const int *p) { return __ldcs(reinterpret_cast<const int4 *>(p)); }

constexpr int kNarrowU = 8;   // nonzeros per block: one 32 B sector of col, one of val, 8 gathers in flight per lane

template <int G, int RED, int COMP, bool ARG>
__global__ void __launch_bounds__(kSpmmThreads) spmm_narrow_kernel(const SpmmArgs a) {
  constexpr int VEC = 4, U = kNarrowU;
  constexpr int GPB = kSpmmThreads / G;
  constexpr bool HAS_VAL = (COMP != C_COPY);
  static_assert(G == 2 || G == 4, "narrow panels: 8 or 16 columns");

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int grp = threadIdx.x / G, gl = threadIdx.x % G;
  const int chunk_id = blockIdx.x * GPB + grp;
  if (chunk_id >= a.num_chunks) return;
  const int lo = chunk_id * a.chunk;
  const int hi = (a.nnz - lo <= a.chunk) ? a.nnz : lo + a.chunk;
  const int colbase = blockIdx.y * (G * VEC) + gl * VEC;
  const bool active = colbase < a.N;
  const int ldcol = active ? colbase : 0;
  const char *__restrict__ Bp = reinterpret_cast<const char *>(a.B + ldcol);
  const unsigned ldb_bytes = (unsigned)(a.ldb * 4);
  const int *__restrict__ rowptr = a.rowptr;
  const uint64_t pol = l2_policy_evict_last();

  int r = row_of_nnz(rowptr, a.M, lo);
  int row_start = __ldg(rowptr + r);
  int row_end = __ldg(rowptr + r + 1);

  float acc[VEC];
  int arg[VEC];
#pragma unroll
  for (int v = 0; v < VEC; v++) { acc[v] = reduce_identity<RED>(); arg[v] = -1; }

  auto store_partial = [&](int slot) {
    const size_t off = ((size_t)chunk_id * 2 + slot) * a.N + colbase;
    *reinterpret_cast<float4 *>(a.part_val + off) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    if (ARG) *reinterpret_cast<int4 *>(a.part_arg + off) = make_int4(arg[0], arg[1], arg[2], arg[3]);
  };
  auto finish_row = [&]() {
    if (active) {
      if (row_start >= lo) {
        float o[VEC];
        const float deg = (float)(row_end - row_start);
#pragma unroll
        for (int v = 0; v < VEC; v++) o[v] = a.mean ? acc[v] / deg : acc[v];
        const size_t off = (size_t)r * a.ldc + colbase;
        if (a.mcast) st_vec_multimem<VEC>(a.dst[0] + off, o);
        else {
          st_vec_cs<VEC>(a.dst[0] + off, o);
          for (int d = 1; d < a.n_dst; d++) st_vec_cs<VEC>(a.dst[d] + off, o);
        }
        if (ARG) st_vec_cs<VEC>(a.E + (size_t)r * a.lde + colbase, arg);
      } else {
        store_partial(0);
      }
    }
#pragma unroll
    for (int v = 0; v < VEC; v++) { acc[v] = reduce_identity<RED>(); arg[v] = -1; }
  };
  auto advance_to = [&](int pos) {
    r += 1;
    row_start = row_end;
    row_end = __ldg(rowptr + r + 1);
    if (row_end <= pos) {
      r = row_of_nnz_from(rowptr, a.M, pos, r);
      row_start = __ldg(rowptr + r);
      row_end = __ldg(rowptr + r + 1);
    }
  };
  auto accumulate = [&](int c, float ev, const float (&bv)[VEC]) {
    float x[VEC];
#pragma unroll
    for (int v = 0; v < VEC; v++) x[v] = compute_op<COMP>(ev, bv[v]);
    reduce_step<RED, ARG, VEC>(acc, arg, x, c);
  };

  // (col, val) of the next block of 8, in registers while the current block is consumed
  int4 nc0 = make_int4(0, 0, 0, 0), nc1 = nc0, nv0 = nc0, nv1 = nc0;
  auto prefetch = [&](int base) {
    if (base + U <= hi) {
      nc0 = ld_stream_i4(a.col + base); nc1 = ld_stream_i4(a.col + base + 4);
      if (HAS_VAL) { nv0 = ld_stream_i4(reinterpret_cast<const int *>(a.val) + base); nv1 = ld_stream_i4(reinterpret_cast<const int *>(a.val) + base + 4); }
    }
  };
  prefetch(lo);
  int base = lo;
  for (; base + U <= hi; base += U) {
    const int cc[U] = {nc0.x, nc0.y, nc0.z, nc0.w, nc1.x, nc1.y, nc1.z, nc1.w};
    const int vb[U] = {nv0.x, nv0.y, nv0.z, nv0.w, nv1.x, nv1.y, nv1.z, nv1.w};
    prefetch(base + U);
    float b[U][VEC];
#pragma unroll
    for (int u = 0; u < U; u++) ld_gather4(b[u], row_addr(Bp, (unsigned)cc[u], ldb_bytes), pol);
    if (base + U <= row_end) {
#pragma unroll
      for (int u = 0; u < U; u++) accumulate(cc[u], HAS_VAL ? __int_as_float(vb[u]) : 1.0f, b[u]);
    } else {
#pragma unroll
      for (int u = 0; u < U; u++) {
        if (base + u >= row_end) { finish_row(); advance_to(base + u); }
        accumulate(cc[u], HAS_VAL ? __int_as_float(vb[u]) : 1.0f, b[u]);
      }
    }
  }
  // ragged end of the last segment (nnz % 8 nonzeros)
#pragma unroll 1
  for (; base < hi; base++) {
    const int c = __ldcs(a.col + base);
    const float ev = HAS_VAL ? __ldcs(a.val + base) : 1.0f;
    float b[VEC];
    ld_gather4(b, row_addr(Bp, (unsigned)c, ldb_bytes), pol);
    if (base >= row_end) { finish_row(); advance_to(base); }
    accumulate(c, ev, b);
  }

  int tail = -1;
  if (row_end == hi) {
    finish_row();
  } else if (row_start >= lo) {
    if (active) store_partial(1);
    tail = r;
  } else {
    if (active) store_partial(0);
  }
  if (gl == 0 && blockIdx.y == 0) a.tail_row[chunk_id] = tail;
}

template <int G, int RED, int COMP, bool ARG>
cudaError_t launch_spmm_narrow(const SpmmArgs &a, dim3 grid, cudaStream_t s) {
  spmm_narrow_kernel<G, RED, COMP, ARG><<<grid, kSpmmThreads, 0, s>>>(a);
  return cudaGetLastError();
}

}  // namespace dgs
