// sddmm.cu — SDDMM for sm_100a: out[e] = dot(D1[row(e), :], D2[col(e), :]) over the nonzeros of a
// CSR or COO pattern, plus the MEAN-scaled and arg-masked flavours the SpMM backward needs.
//
// Replaces (behaviour, not code) the reference kernels
//   sddmm_csr_ebalance_{vec4,vec2,scalar}   src/sddmm/csrsddmm_ebalance.cuh:5-206   (C ABI, CSR)
//   sddmm_coo_ebalance_{vec4,vec2,scalar}   src/sddmm/coosddmm_ebalance.cuh:5-212   (C ABI, COO)
//   sddmmCSR{1,2}Scale<REDUCE>              include/cuda/sddmm_cuda.cuh:222-401     (torch face, MEAN)
//   sddmmCSR1Scale_with_mask                include/cuda/sddmm_cuda.cuh:403-507     (max/min backward)
// and fixes the reference's missing K%32 residue: any K works.
//
// Scheme: edge-balanced like the SpMM.  The edge stream is cut into equal segments; a group of G
// lanes (G*VEC covers one pass over K, e.g. K=256 -> a full warp, 2 float4 per lane) owns a segment.
// Column (and COO row) indices are staged through shared memory 32 at a time; CSR rows are recovered
// by walking rowptr from one binary search per segment (the reference does 4 searches per 4 edges).
// U=8 edges are processed together: their D2 rows are requested back-to-back, the D1 row is shared
// when the 8 edges lie in one CSR row, and the 8 partial dots are reduced with a transposing shuffle
// tree (7 + log2(G/8) shuffles for 8 edges instead of 8*log2(G)); results are stored coalesced.
#include "common.cuh"
#include "spmm.h"

namespace dgs {

constexpr int kSdThreads = 256;
constexpr int kSdBatch = 32;
constexpr int kSdU = 8;

struct SddmmArgs {
  int M, K, nnz;
  const int *rowptr, *row, *col;
  const float *D1, *D2;
  int64_t ld1, ld2;
  const int *E;
  float *out;
  int chunk, num_chunks;
};

// Reduce NV per-lane partials across the G lanes of a group.  On return the lane whose low
// log2(G/NV) bits are zero holds, in v[0], the full sum of edge `edge_slot(gl)`.
template <int G, int NV>
__device__ __forceinline__ void group_multi_reduce(float (&v)[NV], int gl, unsigned gmask) {
  int off = G / 2;
#pragma unroll
  for (int n = NV; n > 1; n >>= 1, off >>= 1) {
    const bool upper = (gl & off) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; i++) {
      const float send = upper ? v[i] : v[i + n / 2];
      const float keep = upper ? v[i + n / 2] : v[i];
      v[i] = keep + __shfl_xor_sync(gmask, send, off, 32);
    }
  }
#pragma unroll
  for (; off >= 1; off >>= 1) v[0] += __shfl_xor_sync(gmask, v[0], off, 32);
}

template <int G, int NV> __device__ __forceinline__ int edge_slot(int gl) {
  int slot = 0, off = G / 2;
#pragma unroll
  for (int n = NV; n > 1; n >>= 1, off >>= 1) slot += (gl & off) ? n / 2 : 0;
  return slot;
}

template <int VEC, int G, bool COO, bool MEAN, bool MASK>
__global__ void __launch_bounds__(kSdThreads) sddmm_kernel(const SddmmArgs a) {
  constexpr int GPB = kSdThreads / G;
  constexpr int PER = kSdBatch / G;
  constexpr int NV = (G < kSdU) ? G : kSdU;       // edges reduced together
  __shared__ int s_col[GPB][kSdBatch + 1];
  __shared__ int s_row[COO ? GPB : 1][kSdBatch + 1];

  const int grp = threadIdx.x / G;
  const int gl = threadIdx.x % G;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << (G & 31)) - 1u) << (((threadIdx.x & 31) / G) * G));
  const int chunk_id = blockIdx.x * GPB + grp;
  if (chunk_id >= a.num_chunks) return;
  const int lo = chunk_id * a.chunk;
  const int hi = (a.nnz - lo <= a.chunk) ? a.nnz : lo + a.chunk;

  int r = 0, row_start = 0, row_end = 0;
  if (!COO) {
    r = row_of_nnz(a.rowptr, a.M, lo);
    row_start = __ldg(a.rowptr + r);
    row_end = __ldg(a.rowptr + r + 1);
  }
  auto advance_to = [&](int pos) {
    r += 1;
    row_start = row_end;
    row_end = __ldg(a.rowptr + r + 1);
    if (row_end <= pos) {
      r = row_of_nnz(a.rowptr, a.M, pos);
      row_start = __ldg(a.rowptr + r);
      row_end = __ldg(a.rowptr + r + 1);
    }
  };

  const int myslot = edge_slot<G, NV>(gl);
  const bool writer = (gl & (G / NV - 1)) == 0;

  int creg[PER], rreg[PER];
  auto prefetch = [&](int base) {
#pragma unroll
    for (int k = 0; k < PER; k++) {
      const int idx = base + gl + k * G;
      const bool ok = idx < hi;
      creg[k] = ok ? __ldcs(a.col + idx) : 0;
      if (COO) rreg[k] = ok ? __ldcs(a.row + idx) : 0;
    }
  };
  prefetch(lo);

  for (int base = lo; base < hi; base += kSdBatch) {
#pragma unroll
    for (int k = 0; k < PER; k++) {
      s_col[grp][gl + k * G] = creg[k];
      if (COO) s_row[grp][gl + k * G] = rreg[k];
    }
    __syncwarp(gmask);
    prefetch(base + kSdBatch);
    const int n = min(kSdBatch, hi - base);
#pragma unroll 1
    for (int j0 = 0; j0 < n; j0 += NV) {
      float p[NV];
      int rows[NV], cols[NV], degs[NV];
#pragma unroll
      for (int u = 0; u < NV; u++) {
        p[u] = 0.0f;
        const int pos = base + j0 + u;
        const bool ok = pos < hi;           // beyond the segment: dummy edge (row 0, col 0), never stored
        cols[u] = s_col[grp][j0 + u];
        degs[u] = 1;
        if (COO) {
          rows[u] = ok ? s_row[grp][j0 + u] : 0;
        } else {
          if (ok && pos >= row_end) advance_to(pos);
          rows[u] = ok ? r : 0;
          if (MEAN) degs[u] = row_end - row_start;
        }
      }
      const bool one_row = rows[0] == rows[NV - 1] && !COO;   // CSR rows are monotone within a segment
      for (int k = gl * VEC; k < a.K; k += G * VEC) {
        float b[NV][VEC];
#pragma unroll
        for (int u = 0; u < NV; u++) ld_vec<VEC>(b[u], a.D2 + (size_t)cols[u] * a.ld2 + k);
        if (one_row && !MASK) {
          float d[VEC];
          ld_vec<VEC>(d, a.D1 + (size_t)rows[0] * a.ld1 + k);
#pragma unroll
          for (int u = 0; u < NV; u++)
#pragma unroll
            for (int v = 0; v < VEC; v++) p[u] += d[v] * b[u][v];
        } else {
#pragma unroll
          for (int u = 0; u < NV; u++) {
            float d[VEC];
            ld_vec<VEC>(d, a.D1 + (size_t)rows[u] * a.ld1 + k);
            if (MASK) {   // only feature positions whose arg index points at this edge's column
              int e[VEC];
              ld_ivec<VEC>(e, a.E + (size_t)rows[u] * a.K + k);
#pragma unroll
              for (int v = 0; v < VEC; v++) p[u] += (e[v] == cols[u]) ? d[v] * b[u][v] : 0.0f;
            } else {
#pragma unroll
              for (int v = 0; v < VEC; v++) p[u] += d[v] * b[u][v];
            }
          }
        }
      }
      group_multi_reduce<G, NV>(p, gl, gmask);
      if (writer) {
        const int pos = base + j0 + myslot;
        if (pos < hi) {
          float res = p[0];
          if (MEAN) {
            int deg = degs[0];
#pragma unroll
            for (int u = 1; u < NV; u++) deg = (myslot == u) ? degs[u] : deg;
            if (deg > 0) res /= (float)deg;   // include/cuda/sddmm_cuda.cuh:259-265
          }
          a.out[pos] = res;
        }
      }
    }
    __syncwarp(gmask);
  }
}

template <int VEC, int G>
static cudaError_t launch_g(const SddmmArgs &a, bool coo, bool mean, bool mask, cudaStream_t s) {
  const int gpb = kSdThreads / G;
  const int grid = (a.num_chunks + gpb - 1) / gpb;
  if (coo) sddmm_kernel<VEC, G, true, false, false><<<grid, kSdThreads, 0, s>>>(a);
  else if (mask) sddmm_kernel<VEC, G, false, false, true><<<grid, kSdThreads, 0, s>>>(a);
  else if (mean) sddmm_kernel<VEC, G, false, true, false><<<grid, kSdThreads, 0, s>>>(a);
  else sddmm_kernel<VEC, G, false, false, false><<<grid, kSdThreads, 0, s>>>(a);
  return cudaGetLastError();
}

template <int VEC> static cudaError_t launch_v(int G, const SddmmArgs &a, bool coo, bool mean, bool mask, cudaStream_t s) {
  switch (G) {
  case 4: return launch_g<VEC, 4>(a, coo, mean, mask, s);
  case 8: return launch_g<VEC, 8>(a, coo, mean, mask, s);
  case 16: return launch_g<VEC, 16>(a, coo, mean, mask, s);
  default: return launch_g<VEC, 32>(a, coo, mean, mask, s);
  }
}

cudaError_t sddmm(const SddmmProblem &p, cudaStream_t stream) {
  if (p.nnz <= 0 || p.K < 0) return cudaSuccess;
  const bool coo = p.rowptr == nullptr;
  if (coo && p.row == nullptr) return cudaErrorInvalidValue;
  const bool mask = p.E != nullptr;
  auto al16 = [](const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  const bool vec4 = (p.K % 4 == 0) && (p.ld1 % 4 == 0) && (p.ld2 % 4 == 0) && al16(p.D1) && al16(p.D2) &&
                    (!mask || al16(p.E));
  int G = 4;
  const int lanes = vec4 ? (p.K + 3) / 4 : p.K;
  while (G < lanes && G < 32) G <<= 1;

  SddmmArgs a;
  a.M = p.M; a.K = p.K; a.nnz = (int)p.nnz;
  a.rowptr = p.rowptr; a.row = p.row; a.col = p.col;
  a.D1 = p.D1; a.D2 = p.D2; a.ld1 = p.ld1; a.ld2 = p.ld2;
  a.E = p.E; a.out = p.out;
  const int64_t resident_groups = (int64_t)device_sm_count() * 4 * (kSdThreads / G);
  int64_t chunk = (p.nnz + resident_groups * 8 - 1) / (resident_groups * 8);
  if (chunk < 32) chunk = 32;
  if (chunk > 4096) chunk = 4096;
  a.chunk = (int)((chunk + kSdBatch - 1) / kSdBatch * kSdBatch);
  a.num_chunks = (int)((p.nnz + a.chunk - 1) / a.chunk);
  ProfileScope prof(3, stream);
  return vec4 ? launch_v<4>(G, a, coo, p.mean != 0 && !coo, mask && !coo, stream)
              : launch_v<1>(G, a, coo, p.mean != 0 && !coo, mask && !coo, stream);
}

}  // namespace dgs
