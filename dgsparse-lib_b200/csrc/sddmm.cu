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
#include <cstdint>
#include <cstdlib>
#include "common.cuh"
#include "spmm.h"
#include "options.h"

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

// row_of_nnz (common.cuh), by the G lanes of a lane group TOGETHER (all of them call it with the same p; gl = lane index in the group, gmask
// = the group's lanes of the warp): a G-ary search — every round the lanes probe G evenly spaced row pointers and one ballot
// picks the sub-range — i.e. log_G(M) dependent loads instead of log_2(M): 4 instead of 18 for a full warp on 169 k rows.
// The start-of-chunk search is pure latency in front of a chunk's first gather, paid once per chunk.
template <int G>
__device__ __forceinline__ int row_of_nnz_group(const int *__restrict__ rowptr, int M, int p, int gl, unsigned gmask) {
  constexpr int LG = (G == 32) ? 5 : (G == 16) ? 4 : (G == 8) ? 3 : 2;
  static_assert(G == 32 || G == 16 || G == 8 || G == 4, "lane groups of 4, 8, 16 or 32");
  const int shift = (G == 32) ? 0 : (int)((threadIdx.x & 31u) - (unsigned)gl);
  constexpr unsigned low = (G == 32) ? 0xffffffffu : ((1u << (G & 31)) - 1u);
  int lo = 0, hi = M;   // smallest idx with rowptr[idx] > p lies in [lo, hi]; rowptr[hi] > p holds throughout (rowptr[M] = nnz > p)
  while (hi - lo >= G) {
    const long long span = hi - lo;
    const int q = lo + (int)((span * (gl + 1)) >> LG);          // probe G - 1 is hi itself
    const unsigned b = (__ballot_sync(gmask, __ldg(rowptr + q) > p) >> shift) & low;
    const int f = b ? __ffs(b) - 1 : G - 1;                      // first probe beyond p (b != 0 for a valid CSR: the last probe is hi)
    const int q_prev = lo + (int)((span * f) >> LG);             // probe f - 1 (f = 0: lo, not a probe)
    hi = lo + (int)((span * (f + 1)) >> LG);
    lo = (f == 0) ? lo : q_prev + 1;
  }
  const int idx = lo + gl;                                       // at most G candidates lo .. hi left
  const unsigned b = (__ballot_sync(gmask, idx >= hi || __ldg(rowptr + idx) > p) >> shift) & low;
  return max(0, min(M - 1, lo + __ffs(b) - 2));   // the clamp only matters for an inconsistent CSR (nnz beyond rowptr[M])
}

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

// THREADS: lane groups never talk to each other, so the CTA size only sets the granularity in which an SM's slots are handed
// out and taken back (see the ring kernel's one-warp CTAs below)
template <int VEC, int G, bool COO, bool MEAN, bool MASK, int THREADS>
__global__ void __launch_bounds__(THREADS) sddmm_kernel(const SddmmArgs a) {
  constexpr int GPB = THREADS / G;
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
    r = row_of_nnz_group<G>(a.rowptr, a.M, lo, gl, gmask);
    row_start = __ldg(a.rowptr + r);
    row_end = __ldg(a.rowptr + r + 1);
  }
  auto advance_to = [&](int pos) {
    r += 1;
    row_start = row_end;
    row_end = __ldg(a.rowptr + r + 1);
    if (row_end <= pos) {
      r = row_of_nnz_from(a.rowptr, a.M, pos, r);
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
        const bool ok = pos < hi;           // beyond the segment: dummy edge (current row, col 0), never stored
        cols[u] = s_col[grp][j0 + u];
        degs[u] = 1;
        if (COO) {
          rows[u] = ok ? s_row[grp][j0 + u] : 0;
        } else {
          if (ok && pos >= row_end) advance_to(pos);
          rows[u] = r;                      // padding keeps the current row: rows[] stays monotone for one_row below
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


// ---- shared-memory ring variant (rows of 256 B .. 4 KB) ---------------------------------------------------------------
// A random 1 KB-row gather is latency bound: the register-staged kernel above holds its loads in 112 registers, sits
// at 20 % warp occupancy and leaves every warp on long-scoreboard.  Here the operand rows go global -> shared with
// cp.async (LDGSTS, 16 B per lane, no data registers): a warp keeps STAGES-1 batches of NB edges (NB D2 rows + the
// distinct D1 rows among them) in flight in its own shared-memory ring while it consumes the oldest batch.
// Index math is done one edge per LANE for 32 edges at a time (column load, CSR row search in a rowptr window,
// source pointers, D1-row de-duplication by ballot) and broadcast with shuffles, so the per-edge instruction stream
// is shuffle + address add + cp.async and the consumer is LDS + FMA with the edges of a batch independent (ILP).
// (Measured dead ends, profiles/README.md: cp.async.bulk/UBLKCP, one 1 KB bulk copy per row, ran 2.3x slower — the
// TMA unit serialises small random copies; the first cp.async ring did its index math serially per edge and was
// instruction-latency bound, time inversely proportional to the warp count.)
// D1 rows are streamed once -> L2 evict_first policy; D2 rows are the reused operand and keep the default policy
// (a fractional evict_last policy on D2, 0.3 .. 0.9 of L2, was measured and changed nothing: 0.198 - 0.205 ms;
// column-range passes that keep one slice of D2 L2-resident per pass, with the pass's edges compacted by shuffles,
// were measured too: 0.208 ms for 1 pass, 0.277 / 0.348 / 0.413 ms for 2 / 3 / 4 — the kernel is bound by per-warp
// issue latency, not by the 909 MB of DRAM traffic, so re-walking the edge list costs more than the misses it saves;
// FEATURE-axis passes — K/p columns per pass with the full leading dimension, so that a narrower D2 slice stays L2-resident
// — lose as well: 0.197 ms for 1 pass, 0.242 for 2 x 128 columns, 0.450 for 4 x 64, tools/exp_sddmm_fsplit.py).
constexpr int kRgNB = 4;          // edges per batch: one per 8-lane group of the copying warp
constexpr int kRgBPS = 32 / kRgNB; // batches per 32-edge superbatch
constexpr int kRgMetaBytes = 2 * kRgBPS * 4 + 2 * 32 * 4 + 2 * 32 * 4;   // two superbatches of packed slots + degrees + row indices

__device__ __forceinline__ uint32_t sd_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sd_cp16(uint32_t dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void sd_cp16_hint(uint32_t dst, const void *src, uint64_t pol) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "l"(pol) : "memory");
}
__device__ __forceinline__ const char *sd_shfl_ptr(const char *p, int src_lane) {
  unsigned long long v = reinterpret_cast<unsigned long long>(p);
  unsigned lo = __shfl_sync(0xffffffffu, (unsigned)v, src_lane), hi = __shfl_sync(0xffffffffu, (unsigned)(v >> 32), src_lane);
  return reinterpret_cast<const char *>(((unsigned long long)hi << 32) | lo);
}

struct SddmmRingArgs {
  SddmmArgs a;
  int wpc;             // warps per CTA
  int nd;              // D1 row slots per stage (2 or 4); a batch's further distinct D1 rows are read from global by the consumer
  uint32_t slot_bytes; // K*4 rounded up to 128
  uint32_t warp_bytes; // shared memory per warp
};

// KCH = ceil(K / 128): 16-byte chunks per lane and row in the consumer; FULLK: K == 128 * KCH (no column predicates)
template <int KCH, int STAGES, bool FULLK, bool COO, bool MEAN>
__global__ void __launch_bounds__(512, 1) sddmm_ring_kernel(const SddmmRingArgs g) {
  extern __shared__ __align__(128) uint8_t sd_smem[];
  constexpr int NB = kRgNB, BPS = kRgBPS;
  constexpr int SEGS = KCH * 4;   // 128-byte segments per row: one per copy instruction of an 8-lane group
  const SddmmArgs &a = g.a;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = lane >> 3, ql = lane & 7;
  const uint32_t rowbytes = (uint32_t)a.K * 4u;
  // per-warp carve-up: [STAGES][NB + nd] row slots | packed D1 slots per batch [2][BPS] | degree per edge [2][32] | row per edge [2][32]
  uint8_t *wbase = sd_smem + (size_t)warp * g.warp_bytes;
  const uint32_t rows_u32 = sd_smem_u32(wbase);
  constexpr bool OVF = KCH <= 4;   // K > 512: always four D1 slots (the overflow path costs the 8-chunk consumer its registers)
  const int nd = OVF ? g.nd : NB;
  const uint32_t stage_bytes = (uint32_t)(NB + nd) * g.slot_bytes;
  int *s_slots = reinterpret_cast<int *>(wbase + (size_t)STAGES * stage_bytes);
  int *s_deg = s_slots + 2 * BPS;
  int *s_rowidx = s_deg + 2 * 32;
  uint64_t pol_stream;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));

  const int chunk_id = blockIdx.x * g.wpc + warp;
  if (chunk_id >= a.num_chunks) return;
  const int lo = chunk_id * a.chunk;
  const int hi = (a.nnz - lo <= a.chunk) ? a.nnz : lo + a.chunk;
  const int nb_total = (hi - lo + NB - 1) / NB;

  int r_hint = COO ? 0 : row_of_nnz_group<32>(a.rowptr, a.M, lo, lane, 0xffffffffu);   // row of the chunk's first edge (uniform)
  bool in_k[KCH];
#pragma unroll
  for (int j = 0; j < KCH; j++) in_k[j] = FULLK || (lane + 32 * j) * 4 < a.K;

  // ---- per-lane state of the current superbatch: lane u <-> edge min(lo + 32*sb + u, hi - 1) ----
  // (edges past the segment end are clamped to its last edge: they are fetched and multiplied like the others and
  //  simply not stored, which keeps every batch full and the inner loops free of validity branches)
  const char *my_src2 = nullptr, *my_src1 = nullptr;   // base of this edge's D2 / D1 row
  unsigned new_mask = 0;                               // bit u: edge u opens a new D1 slot in its batch (uniform)

  auto load_super = [&](int sb) {
    const int pos = min(lo + sb * 32 + lane, hi - 1);
    const int c = __ldcs(a.col + pos);
    int rr, deg = 1;
    if (COO) {
      rr = __ldcs(a.row + pos);
    } else {
      // first idx in (r_hint, M] with rowptr[idx] > pos; try a 32-row window first, else the whole tail
      int l = r_hint + 1, h = min(a.M, r_hint + 32);
      if (__ldg(a.rowptr + h) <= pos) { l = h + 1; h = a.M; }
      while (l < h) {
        const int mid = (l + h) >> 1;
        if (__ldg(a.rowptr + mid) > pos) h = mid; else l = mid + 1;
      }
      rr = l - 1;
      if (MEAN) deg = __ldg(a.rowptr + rr + 1) - __ldg(a.rowptr + rr);
    }
    my_src2 = reinterpret_cast<const char *>(a.D2 + (size_t)c * a.ld2);
    my_src1 = reinterpret_cast<const char *>(a.D1 + (size_t)rr * a.ld1);
    const int prev = __shfl_up_sync(0xffffffffu, rr, 1);
    new_mask = __ballot_sync(0xffffffffu, (lane % NB) == 0 || rr != prev);
    if ((lane % NB) == 0) {   // D1 slot of each edge of the batch, 4 bits each: running count of "new" flags
      const unsigned gb = (new_mask >> lane) & 15u;
      const int s1 = (gb >> 1) & 1, s2 = s1 + ((gb >> 2) & 1), s3 = s2 + ((gb >> 3) & 1);
      s_slots[(sb & 1) * BPS + lane / NB] = (s1 << 4) | (s2 << 8) | (s3 << 12);
    }
    if (MEAN) s_deg[(sb & 1) * 32 + lane] = deg;
    if (nd < NB) s_rowidx[(sb & 1) * 32 + lane] = rr;   // for the consumer's global read of a batch's third / fourth D1 row
    if (!COO) r_hint = __shfl_sync(0xffffffffu, rr, 31);   // rows are monotone: the next superbatch searches from here
  };

  auto issue = [&](int b) {   // always commits exactly one cp.async group (possibly empty)
    if (b < nb_total) {
      if (b % BPS == 0) load_super(b / BPS);
      const int u0 = (b % BPS) * NB;
      const unsigned gb = (new_mask >> u0) & 15u;
      // 8-lane group q fetches edge u0 + q: its D2 row, and its D1 row when that edge opens a new slot
      const char *p2 = sd_shfl_ptr(my_src2, u0 + q) + ql * 16;
      const char *p1 = sd_shfl_ptr(my_src1, u0 + q) + ql * 16;
      const uint32_t sbase = rows_u32 + (uint32_t)(b % STAGES) * stage_bytes + (uint32_t)ql * 16u;
      const uint32_t dst2 = sbase + (uint32_t)q * g.slot_bytes;
#pragma unroll
      for (int i = 0; i < SEGS; i++)
        if (FULLK || i * 128u + ql * 16u < rowbytes) sd_cp16(dst2 + 128u * i, p2 + 128 * i);
      const int slot1 = __popc(gb & ((2u << q) - 1u)) - 1;   // D1 slot of edge q: new flags among edges 0 .. q, minus one
      if (((gb >> q) & 1u) && slot1 < nd) {
        const uint32_t dst1 = sbase + (uint32_t)(NB + slot1) * g.slot_bytes;
#pragma unroll
        for (int i = 0; i < SEGS; i++)
          if (FULLK || i * 128u + ql * 16u < rowbytes) sd_cp16_hint(dst1 + 128u * i, p1 + 128 * i, pol_stream);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  const int myslot = edge_slot<32, NB>(lane);
  const bool writer = (lane & (32 / NB - 1)) == 0;

#pragma unroll
  for (int b = 0; b < STAGES - 1; b++) issue(b);
  for (int b = 0; b < nb_total; b++) {
    issue(b + STAGES - 1);
    asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 1) : "memory");   // this lane's copies of batch b landed
    __syncwarp();                                                            // ... and every other lane's
    const uint8_t *srows = wbase + (size_t)(b % STAGES) * stage_bytes;
    const int mi = ((b / BPS) & 1) * BPS + (b % BPS);
    const int packed = s_slots[mi];
    float4 y[NB][KCH];
#pragma unroll
    for (int e = 0; e < NB; e++) {
      const float4 *yr = reinterpret_cast<const float4 *>(srows + (size_t)e * g.slot_bytes);
#pragma unroll
      for (int j = 0; j < KCH; j++)
        if (in_k[j]) y[e][j] = yr[lane + 32 * j];
    }
    float p[NB];
    float4 x[KCH];
#pragma unroll
    for (int e = 0; e < NB; e++) {
      const int d1 = (packed >> (4 * e)) & 15;
      if (e == 0 || d1 != ((packed >> (4 * (e - 1))) & 15)) {   // uniform: the D1 row changes with this edge
        if (!OVF || d1 < nd) {
          const float4 *xr = reinterpret_cast<const float4 *>(srows + (size_t)(NB + d1) * g.slot_bytes);
#pragma unroll
          for (int j = 0; j < KCH; j++)
            if (in_k[j]) x[j] = xr[lane + 32 * j];
        } else {   // rare (nd = 2: three or four distinct rows among four consecutive edges): no ring slot, read the row from global
          const int rr = s_rowidx[((b / BPS) & 1) * 32 + (b % BPS) * NB + e];
          const float4 *xr = reinterpret_cast<const float4 *>(a.D1 + (size_t)rr * a.ld1);
#pragma unroll
          for (int j = 0; j < KCH; j++)
            if (in_k[j]) x[j] = __ldg(xr + lane + 32 * j);
        }
      }
      float acc = 0.0f;
#pragma unroll
      for (int j = 0; j < KCH; j++) {
        if (in_k[j]) {
          acc += x[j].x * y[e][j].x;
          acc += x[j].y * y[e][j].y;
          acc += x[j].z * y[e][j].z;
          acc += x[j].w * y[e][j].w;
        }
      }
      p[e] = acc;
    }
    group_multi_reduce<32, NB>(p, lane, 0xffffffffu);
    const int pos = lo + b * NB + myslot;
    if (writer && pos < hi) {
      float res = p[0];
      if (MEAN) {
        const int deg = s_deg[((b / BPS) & 1) * 32 + (b % BPS) * NB + myslot];
        if (deg > 0) res /= (float)deg;
      }
      __stcs(a.out + pos, res);
    }
    __syncwarp();   // every lane is done with this stage's rows before it is refilled
  }
}

constexpr int kRingSmemOptIn = 227 * 1024;   // the most a CTA may ask for on sm_100; rings use <= ~200 KB

// occ_out != nullptr: no launch, *occ_out = CTAs of this geometry (g.wpc warps, smem bytes) one SM really holds — the
// registers of the instantiation, the ring's shared memory and the thread limit together — the last answer cached per device.
template <int KCH, int STAGES>
static cudaError_t launch_ring(const SddmmRingArgs &g, int grid, size_t smem, bool coo, bool mean, cudaStream_t s, int *occ_out) {
  cudaError_t e;
  const bool fullk = g.a.K == 128 * KCH;
  // the opt-in to > 48 KB of dynamic shared memory is per kernel and per device, not per launch: set it once (a
  // cudaFuncSetAttribute on every call was ~2 us of host time in front of a 15 us kernel)
#define DGS_RING(FULL_, COO_, MEAN_)                                                                                       \
  do {                                                                                                                     \
    static bool optin[64] = {false};                                                                                       \
    static int occ_wpc[64] = {0}, occ_n[64] = {0};                                                                         \
    static size_t occ_smem[64] = {0};                                                                                      \
    int dev_ = 0;                                                                                                          \
    if ((e = cudaGetDevice(&dev_)) != cudaSuccess) return e;                                                               \
    if (dev_ < 0 || dev_ >= 64 || !optin[dev_]) {                                                                          \
      if ((e = cudaFuncSetAttribute(sddmm_ring_kernel<KCH, STAGES, FULL_, COO_, MEAN_>,                                    \
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, kRingSmemOptIn)) != cudaSuccess) return e; \
      if (dev_ >= 0 && dev_ < 64) optin[dev_] = true;                                                                      \
    }                                                                                                                      \
    if (occ_out != nullptr) {                                                                                              \
      const bool slot_ = dev_ >= 0 && dev_ < 64;                                                                           \
      int n_ = -1;                                                                                                         \
      if (slot_ && occ_wpc[dev_] == g.wpc && occ_smem[dev_] == smem) n_ = occ_n[dev_];                                     \
      if (n_ < 0) {                                                                                                        \
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n_, sddmm_ring_kernel<KCH, STAGES, FULL_, COO_, MEAN_>,    \
                                                               g.wpc * 32, smem)) != cudaSuccess) return e;                \
        if (slot_) { occ_wpc[dev_] = 0; occ_smem[dev_] = smem; occ_n[dev_] = n_; occ_wpc[dev_] = g.wpc; }                  \
      }                                                                                                                    \
      *occ_out = n_;                                                                                                       \
      return cudaSuccess;                                                                                                  \
    }                                                                                                                      \
    sddmm_ring_kernel<KCH, STAGES, FULL_, COO_, MEAN_><<<grid, g.wpc * 32, smem, s>>>(g);                                  \
  } while (0)
  if (fullk) {
    if (coo) DGS_RING(true, true, false);
    else if (mean) DGS_RING(true, false, true);
    else DGS_RING(true, false, false);
  } else {
    if (coo) DGS_RING(false, true, false);
    else if (mean) DGS_RING(false, false, true);
    else DGS_RING(false, false, false);
  }
#undef DGS_RING
  return cudaGetLastError();
}

template <int VEC, int G, int THREADS>
static cudaError_t launch_gt(const SddmmArgs &a, bool coo, bool mean, bool mask, cudaStream_t s) {
  const int gpb = THREADS / G;
  const int grid = (a.num_chunks + gpb - 1) / gpb;
  if (coo) sddmm_kernel<VEC, G, true, false, false, THREADS><<<grid, THREADS, 0, s>>>(a);
  else if (mask) sddmm_kernel<VEC, G, false, false, true, THREADS><<<grid, THREADS, 0, s>>>(a);
  else if (mean) sddmm_kernel<VEC, G, false, true, false, THREADS><<<grid, THREADS, 0, s>>>(a);
  else sddmm_kernel<VEC, G, false, false, false, THREADS><<<grid, THREADS, 0, s>>>(a);
  return cudaGetLastError();
}

template <int VEC, int G>
static cudaError_t launch_g(const SddmmArgs &a, bool coo, bool mean, bool mask, cudaStream_t s) {
  // 64-thread CTAs: 3 - 16 % faster than 256 wherever a row is shorter than 256 B, level above (tools/exp_sddmm_threads.py,
  // profiles/r02_exp_sddmm_threads.jsonl: arxiv-like K = 16 / 32 / 48 34.4 / 43.2 / 84.7 -> 29.0 / 39.2 / 77.9 us)
  if (option(OPT_SDDMM_THREADS) == 256) return launch_gt<VEC, G, kSdThreads>(a, coo, mean, mask, s);
  return launch_gt<VEC, G, 64>(a, coo, mean, mask, s);
}

template <int VEC> static cudaError_t launch_v(int G, const SddmmArgs &a, bool coo, bool mean, bool mask, cudaStream_t s) {
  switch (G) {
  case 4: return launch_g<VEC, 4>(a, coo, mean, mask, s);
  case 8: return launch_g<VEC, 8>(a, coo, mean, mask, s);
  case 16: return launch_g<VEC, 16>(a, coo, mean, mask, s);
  default: return launch_g<VEC, 32>(a, coo, mean, mask, s);
  }
}

static thread_local int g_last_geo[3] = {0, 0, 0};
void sddmm_last_geometry(int *wpc, int *ctas_per_sm, int *chunk) {
  if (wpc) *wpc = g_last_geo[0];
  if (ctas_per_sm) *ctas_per_sm = g_last_geo[1];
  if (chunk) *chunk = g_last_geo[2];
}

cudaError_t sddmm(const SddmmProblem &p, cudaStream_t stream) {
  if (p.nnz <= 0 || p.K < 0) return cudaSuccess;
  const bool coo = p.rowptr == nullptr;
  if (coo && p.row == nullptr) return cudaErrorInvalidValue;
  const bool mask = p.E != nullptr;
  auto al16 = [](const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  const bool vec4 = (p.K % 4 == 0) && (p.ld1 % 4 == 0) && (p.ld2 % 4 == 0) && al16(p.D1) && al16(p.D2) &&
                    (!mask || al16(p.E));
  // rows of 256 B .. 4 KB made of aligned 16-byte chunks: shared-memory ring kernel
  // latency regime (< 1 M edges) at K = 64: one 16-lane pass of the register kernel beats the ring's set-up
  // (ca-CondMat 19.1 vs 26.8 us, p2p-Gnutella31 14.7 vs 16.7 us; at K >= 128 the ring is level or ahead)
  const bool small_k64 = p.K == 64 && p.nnz < (1 << 20) && option(OPT_SDDMM_NO_RING) != 0;   // sddmm_no_ring = 0 forces the ring
  if (vec4 && !mask && p.K >= 64 && p.K <= 1024 && !small_k64 && option(OPT_SDDMM_NO_RING) != 1) {
    SddmmRingArgs g;
    SddmmArgs &a = g.a;
    a.M = p.M; a.K = p.K; a.nnz = (int)p.nnz;
    a.rowptr = p.rowptr; a.row = p.row; a.col = p.col;
    a.D1 = p.D1; a.D2 = p.D2; a.ld1 = p.ld1; a.ld2 = p.ld2;
    a.E = nullptr; a.out = p.out;
    g.slot_bytes = ((uint32_t)p.K * 4u + 127u) & ~127u;
    // ring geometry: 4 edges per batch, STAGES batches per warp ring.  The warps of the ring kernel never talk to each other,
    // so the CTA is only the unit in which an SM's registers and shared memory are handed out and taken back: with ONE warp
    // per CTA a finished warp's ring goes to the next chunk at once instead of waiting for the slowest warp of its CTA, and
    // the SM holds as many rings as fit (13 at K = 256 instead of 12, 24 at K = 128 instead of 22, 28 at K = 64).  Measured
    // (tools/exp_sddmm_ring.py, profiles/r02_exp_sddmm_ring.jsonl; 1 .. 16 warps per CTA, the residency from the occupancy
    // API): arxiv-like K = 64 / 128 / 256 / 512 0.084 / 0.098 / 0.186 / 0.386 -> 0.077 / 0.090 / 0.170 / 0.373 ms, ca-CondMat
    // K = 256 32.4 -> 26.8 us (the last line that trailed the reference's kernel, 29.1 us).
    int STAGES = 2;
    if (option(OPT_SDDMM_STAGES) == 3) STAGES = 3;
    // D1 slots per stage: four consecutive edges of a CSR rarely touch more than two rows, so two slots (and a global read by
    // the consumer for a batch's third / fourth row) leave room for a third more rings per SM (16 instead of 12 at K = 256,
    // 8 instead of 6 at K = 512).  Measured (tools/exp_sddmm_ring.py --d1, profiles/r02_exp_sddmm_ring_d1.jsonl): level on
    // the arxiv-like graph (the kernel is not short of rings there), 10 - 18 % faster on the two fixtures at K >= 256, where
    // more resident warps mean fewer waves (p2p-Gnutella31 K = 256 / 512 24.3 / 50.0 -> 21.0 / 43.3 us).  COO rows need not be
    // sorted and matrices of one-edge rows would take the global read on every batch: both keep four slots.
    g.nd = (!coo && p.nnz >= 2 * (int64_t)p.M) ? 2 : kRgNB;
    if (option(OPT_SDDMM_D1SLOTS) == 2 || option(OPT_SDDMM_D1SLOTS) == 4) g.nd = option(OPT_SDDMM_D1SLOTS);
    if (p.K > 512) g.nd = kRgNB;   // the 8-chunk instantiations are compiled without the overflow path
    g.warp_bytes = (uint32_t)STAGES * (uint32_t)(kRgNB + g.nd) * g.slot_bytes + (uint32_t)kRgMetaBytes;
    g.warp_bytes = (g.warp_bytes + 127u) & ~127u;
    int wpc = g.warp_bytes <= 220u * 1024u ? 1 : 0;
    if (option(OPT_SDDMM_WPC) >= 1 && option(OPT_SDDMM_WPC) <= 16 && (size_t)option(OPT_SDDMM_WPC) * g.warp_bytes <= 220u * 1024u)
      wpc = option(OPT_SDDMM_WPC);
    if (wpc >= 1) {
      g.wpc = wpc;
      const size_t smem = (size_t)wpc * g.warp_bytes;
      const bool mean = p.mean != 0 && !coo;
      // one ring geometry, two uses: the occupancy query (occ != nullptr) and the launch
      auto ring = [&](int grid, int *occ) -> cudaError_t {
#define DGS_RING_GEO(KCH_)                                                                         \
  do {                                                                                             \
    if (STAGES == 3) return launch_ring<KCH_, 3>(g, grid, smem, coo, mean, stream, occ);           \
    return launch_ring<KCH_, 2>(g, grid, smem, coo, mean, stream, occ);                            \
  } while (0)
        if (p.K <= 128) DGS_RING_GEO(1);
        if (p.K <= 256) DGS_RING_GEO(2);
        if (p.K <= 512) DGS_RING_GEO(4);
        DGS_RING_GEO(8);
#undef DGS_RING_GEO
      };
      // CTAs one SM really holds (registers of the instantiation, ring bytes, thread limit): the wave arithmetic below is
      // only as good as this number
      int ctas_per_sm = 1;
      cudaError_t eo = ring(0, &ctas_per_sm);
      if (eo != cudaSuccess) return eo;
      if (ctas_per_sm < 1) ctas_per_sm = 1;
      const int64_t resident_warps = (int64_t)device_sm_count() * wpc * ctas_per_sm;
      // Edges per warp.  Large inputs: ~6 chunks per resident warp for dynamic balance.  When that would cut chunks shorter
      // than ~112 edges (arxiv-like: 1.17 M edges on 1 184 .. 4 292 resident warps) the per-chunk costs — the row search, the
      // ring's fill and drain — show: take chunks of at most ~112 edges instead, sized so that the grid is a whole number of
      // waves of resident warps (tools/exp_sddmm_ring.py --chunks / --d1, profiles/r02_exp_sddmm_ring_chunks.jsonl,
      // r02_exp_sddmm_ring_d1.jsonl: arxiv-like K = 64 48 -> 96 edges 0.0777 -> 0.072 ms, K = 128 56 -> 92 0.0906 -> 0.075 ms,
      // K = 256 / 512 flat between 64 and 220; 256 and more lose again, and a ragged last wave costs ~7 %).
      int64_t chunk = (p.nnz + resident_warps * 6 - 1) / (resident_warps * 6);
      if (chunk < 112) {
        int64_t waves = (p.nnz + resident_warps * 112 - 1) / (resident_warps * 112);
        if (waves < 1) waves = 1;
        chunk = (p.nnz + resident_warps * waves - 1) / (resident_warps * waves);
      }
      if (chunk > 8192) chunk = 8192;
      chunk = (chunk + kRgNB - 1) / kRgNB * kRgNB;
      // Latency regime (a few waves of warps at most): a warp walks its edges 4 at a time, so the call takes
      // ceil(warps / resident warps) x chunk batches — choose the chunk that wastes no wave.  Measured on the reference's
      // fixtures (us, chunk 64 -> chosen): p2p-Gnutella31 K=256 33.4 -> 25.6 (96), K=512 57.8 -> 54.3; ca-CondMat K=128
      // 26.9 -> 20.9 (96), K=256 33.0 -> 31.9 (128), K=512 61.6 -> 57.5 (128); a chunk that leaves a sliver of a second wave
      // is the worst case (ca-CondMat K=256 at 96: 43.4).
      if (p.nnz <= resident_warps * 64 * 4) {
        int64_t best = 64, best_cost = INT64_MAX;
        for (int64_t c = 32; c <= 512; c += 32) {
          const int64_t nw = (p.nnz + c - 1) / c, waves = (nw + resident_warps - 1) / resident_warps, cost = waves * c;
          if (cost < best_cost || (cost == best_cost && c > best)) { best = c; best_cost = cost; }
        }
        chunk = best;
      }
      if (option(OPT_SDDMM_CHUNK) >= 32) chunk = option(OPT_SDDMM_CHUNK);
      a.chunk = (int)chunk;
      a.num_chunks = (int)((p.nnz + a.chunk - 1) / a.chunk);
      const int grid = (a.num_chunks + wpc - 1) / wpc;
      g_last_geo[0] = wpc; g_last_geo[1] = ctas_per_sm; g_last_geo[2] = a.chunk;
      ProfileScope prof(3, stream);
      return ring(grid, nullptr);
    }
  }
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
  g_last_geo[0] = 0; g_last_geo[1] = 0; g_last_geo[2] = a.chunk;
  ProfileScope prof(3, stream);
  return vec4 ? launch_v<4>(G, a, coo, p.mean != 0 && !coo, mask && !coo, stream)
              : launch_v<1>(G, a, coo, p.mean != 0 && !coo, mask && !coo, stream);
}

}  // namespace dgs
