// spconv.cu — sparse 3-D convolution, fused gather-GEMM-scatter (forward and the dX backward), for sm_100a.
//
//   out[omap[p], :] += in[imap[p], :] @ W[k]      for every kernel offset k, pair p in [kpos[k], kpos[k+1])
//
// Replaces spconv_fwd_fused / spconv_bwd_fused (dX half) of the reference, src/cuda/spconv_cuda.cu:18-253, and its
// _fgms_fusion_{fp32,tf32,fp16_tc4}* kernels, include/cuda/spconv.cuh (tf32 model kernel :1100-1248).
//
// B200 design (tensor path): the pair list of every kernel offset is cut into tiles of 128 pairs (the
// reference's q = 128 quantisation of kpos -> qkpos, test/test_spconv.py:5-14, so one tile never mixes two
// offsets).  A CTA of 4 warps takes a run of consecutive tiles.  Per tile it
//   1. gathers the 128 input rows named by imap with 16-byte loads (8 lanes cover one 128-byte row segment),
//      rounds them to tf32 (cvt.rna) or packs them to bf16, and writes them into shared memory in the UMMA
//      canonical K-major SWIZZLE_128B layout (row r, 16-byte chunk c -> r*128 + ((c ^ (r & 7)) << 4));
//   2. keeps W[k]^T (pre-transposed / pre-converted once per call by spconv_prep_weights_kernel, so that it is
//      K-major too) resident in shared memory while consecutive tiles share the same offset k;
//   3. one elected thread issues tcgen05.mma (kind::tf32 or kind::f16, M = 128, N = c_out tile, K = 8 or 16 per
//      instruction) with the fp32 accumulator in TMEM, and commits to an mbarrier;
//   4. every warp reads its 32 accumulator rows back with tcgen05.ld (thread t = pair t of the tile) and
//      scatters them with vectorised red.global.add.v4.f32 into out[omap[p], :].
// Several CTAs are resident per SM (48 KB smem, <= 128 TMEM columns each) so one CTA's gather overlaps another's
// MMA and scatter.  The exact-fp32 path (arch80 = false in the reference) is a SIMT kernel further down.
//
// Fixes of the reference: the output is zeroed by the call, every dtype path writes fp32 into the fp32
// output, `separate_mid` runs the centre offset as identity-mapped tiles of the same kernel instead of a cuBLAS call.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdlib>
#include "common.cuh"
#include "options.h"
#include "spmm.h"
#include "spconv.h"

namespace dgs {
namespace {

constexpr int kTileM = 128;          // pairs per tile == UMMA M == the reference's q
constexpr int kAtomBytes = 128;      // one swizzle row: 32 tf32 or 64 bf16 channels
constexpr int kMaxAtomsPerGroup = 2; // K staged per MMA group: 64 fp32 / 128 bf16 channels
constexpr int kMaxNT = 128;          // widest c_out tile (TMEM columns)

struct SpconvArgs {
  const int *kpos, *qkpos, *imap, *omap;
  const float *in;     // [in_rows, ld_in]
  const uint8_t *Wt;   // [k_vol][n_rows_pad / NT][n_atoms][NT][128 B swizzled] in the MMA dtype: shared-memory images (prep kernel)
  float *out;          // [out_rows, ld_out]
  int64_t ld_in, ld_out;
  int k_vol, c_in, c_out;     // GEMM K and N of this launch (swapped for the dX backward)
  int n_atoms;                // K atoms: ceil(c_in / elements per atom)
  int n_rows_pad;             // rows of Wt per offset (gridDim.y * NT)
  int NT;                     // c_out tile, multiple of 16, <= 128
  int tmem_cols;              // power of two >= max(32, NT)
  int n_map_tiles;            // tiles driven by kpos/qkpos/imap/omap
  int n_id_tiles;             // identity-mapped tiles of the centre offset (separate_mid)
  int id_rows, mid_k;         // rows covered by the identity tiles, centre offset index
  int tiles_per_cta;
  // pipelined kernel: gather source in the MMA dtype (fp32 input itself for tf32, a bf16 copy for bf16)
  const uint8_t *feat;
  int64_t feat_pitch;        // bytes between rows
  int feat_valid_bytes;      // bytes of a row that hold channels (the rest of the last atom is zero-filled)
  int n_stages;              // shared-memory A stages
  int f16;                   // 16-bit operand format of the kind::f16 path: 0 = bf16, 1 = fp16 (IEEE half)
};

// ---- PTX wrappers -------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done;
  const uint32_t a = smem_u32(bar);
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
  } while (!done);
}
// TMA bulk copy global -> shared (cp.async.bulk, the copy engine writes shared memory and completes `bytes` on the mbarrier).
// Used for the prepared weight panels, which are stored as the exact shared-memory image (spconv_prep_weights_kernel).
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// one thread: the panel of `atoms` atoms x NT rows x 128 B, one bulk copy per atom (<= 32 KB each)
__device__ __forceinline__ void load_w_panel(uint32_t sW, const uint8_t *src, int atoms, int NT, uint64_t *bar) {
  const uint32_t per_atom = (uint32_t)NT * 128u;
  mbar_expect_tx(bar, per_atom * (uint32_t)atoms);
  for (int at = 0; at < atoms; at++) bulk_g2s(sW + (uint32_t)at * per_atom, src + (size_t)at * per_atom, per_atom, bar);
}
// a wait that cannot hang the GPU: a copy that never completes (a bad address would fault, not stall) traps after ~1 s
__device__ __forceinline__ void mbar_wait_bounded(uint64_t *bar, uint32_t parity) {
  uint32_t done;
  const uint32_t a = smem_u32(bar);
  for (uint32_t spins = 0;; spins++) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
    if (done) return;
    if (spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout): start address
// >> 4 in [0,14), LBO >> 4 in [16,30) (unused for swizzled K-major), SBO >> 4 in [32,46) = 1024 B between 8-row
// groups, version 1 in [46,48), layout type 2 (SWIZZLE_128B) in [61,64).
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): D fp32 (1 << 4), A/B format in [7,10)/[10,13)
// (kind::tf32: 2 = tf32; kind::f16: 0 = fp16, 1 = bf16), both K-major, N >> 3 in [17,23), M >> 4 in [24,29).
template <int KIND> __device__ __forceinline__ uint32_t umma_idesc(int n, int f16) {
  const uint32_t fmt = KIND == 0 ? 2u : (f16 ? 0u : 1u);
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}

template <int KIND> __device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (KIND == 0)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 16 consecutive fp32 accumulator columns of this thread's TMEM lane.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
#pragma unroll
  for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void red_add_v4(float *p, float a, float b, float c, float d) {
  asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t *>(&h);
}
// two fp32 -> one 32-bit word of 16-bit operands, round to nearest: fp16 (the reference's half kernels,
// include/cuda/spconv.cuh:1408-1552: 10 mantissa bits) or bf16 (7 bits, fp32 range)
__device__ __forceinline__ uint32_t pack_16(float lo, float hi, int f16) {
  if (f16) { __half2 h = __floats2half2_rn(lo, hi); return *reinterpret_cast<uint32_t *>(&h); }
  return pack_bf16(lo, hi);
}

template <int KIND> struct Kind;
template <> struct Kind<0> { static constexpr int kElemBytes = 4, kElemsPerAtom = 32, kUmmaK = 8; };   // tf32
template <> struct Kind<1> { static constexpr int kElemBytes = 2, kElemsPerAtom = 64, kUmmaK = 16; };  // bf16

// One 16-byte shared-memory chunk of a gathered row: channels [ch, ch + 16 / elem bytes) of `row` (zeros when the
// row is padding or the channels lie beyond c_in).
template <int KIND> __device__ __forceinline__ uint4 gather_chunk(const float *in, int64_t ld_in, int row, int ch, int c_in, int f16) {
  uint4 o = make_uint4(0u, 0u, 0u, 0u);
  if (row < 0) return o;
  const float *src = in + (int64_t)row * ld_in + ch;
  if (KIND == 0) {
    if (ch < c_in) {  // c_in % 4 == 0: a chunk is all-or-nothing
      const float4 v = __ldg(reinterpret_cast<const float4 *>(src));
      o = make_uint4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
    }
  } else {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (ch < c_in) a = __ldg(reinterpret_cast<const float4 *>(src));
    if (ch + 4 < c_in) b = __ldg(reinterpret_cast<const float4 *>(src + 4));
    o = make_uint4(pack_16(a.x, a.y, f16), pack_16(a.z, a.w, f16), pack_16(b.x, b.y, f16), pack_16(b.z, b.w, f16));
  }
  return o;
}

// ---- the fused tile kernel ------------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(128) spconv_fgms_tc_kernel(const SpconvArgs a) {
  using KD = Kind<KIND>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t s_bar, s_wbar;
  __shared__ uint32_t s_tmem;
  __shared__ int s_inrow[kTileM];

  const int tid = threadIdx.x, warp = tid >> 5;
  // 1024-byte aligned operand buffers (SWIZZLE_128B atoms are 8 rows x 128 B)
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int ga_max = a.n_atoms < kMaxAtomsPerGroup ? a.n_atoms : kMaxAtomsPerGroup;
  const uint32_t sA = base;                                       // [ga][128 rows][128 B]
  const uint32_t sW = base + (uint32_t)ga_max * kTileM * kAtomBytes;  // [ga][NT rows][128 B]
  const bool w_resident = a.n_atoms <= kMaxAtomsPerGroup;         // whole K fits one group: keep W[k] across tiles

  if (warp == 0) tmem_alloc(&s_tmem, (uint32_t)a.tmem_cols);
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    mbar_init(&s_wbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t idesc = umma_idesc<KIND>(a.NT, a.f16);
  const int n0 = blockIdx.y * a.NT;
  uint32_t phase = 0, wphase = 0;
  int resident_k = -1;

  const int n_tiles = a.n_map_tiles + a.n_id_tiles;
  const int t_begin = blockIdx.x * a.tiles_per_cta;
  const int t_end = min(n_tiles, t_begin + a.tiles_per_cta);

  for (int tile = t_begin; tile < t_end; tile++) {
    // ---- which pairs: offset k, this thread's (input row, output row) ----
    int k, in_row = -1, out_row = -1;
    if (tile < a.n_map_tiles) {
      const int q0 = tile * kTileM;
      k = upper_bound_i32(a.qkpos, a.k_vol + 1, q0) - 1;
      const int p = q0 - __ldg(a.qkpos + k) + __ldg(a.kpos + k) + tid;
      if (p < __ldg(a.kpos + k + 1)) { in_row = __ldg(a.imap + p); out_row = __ldg(a.omap + p); }
    } else {
      k = a.mid_k;
      const int r = (tile - a.n_map_tiles) * kTileM + tid;
      if (r < a.id_rows) in_row = out_row = r;
    }
    s_inrow[tid] = in_row;
    __syncthreads();

    for (int a0 = 0; a0 < a.n_atoms; a0 += kMaxAtomsPerGroup) {
      const int ga = min(kMaxAtomsPerGroup, a.n_atoms - a0);
      // ---- gather A: item = (row, atom, 16 B chunk); 8 consecutive lanes = one 128 B row segment ----
      const int per_thread = ga * 8;  // 128 rows * ga * 8 chunks / 128 threads
      for (int it0 = 0; it0 < per_thread; it0 += 8) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const int idx = (it0 + u) * 128 + tid;
          const int c = idx & 7, at = (idx >> 3) % ga, r = idx / (8 * ga);
          const int ch = (a0 + at) * KD::kElemsPerAtom + c * (16 / KD::kElemBytes);
          v[u] = gather_chunk<KIND>(a.in, a.ld_in, s_inrow[r], ch, a.c_in, a.f16);
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const int idx = (it0 + u) * 128 + tid;
          const int c = idx & 7, at = (idx >> 3) % ga, r = idx / (8 * ga);
          const uint32_t dst = sA + (uint32_t)at * (kTileM * kAtomBytes) + (uint32_t)r * kAtomBytes + (uint32_t)((c ^ (r & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(v[u].x), "r"(v[u].y), "r"(v[u].z), "r"(v[u].w) : "memory");
        }
      }
      // ---- W[k]^T rows [n0, n0 + NT), atoms [a0, a0 + ga): straight copy of the prepared K-major panel ----
      const bool load_w = !(w_resident && resident_k == k);
      if (load_w) {
        // the prepared panel IS the shared-memory image (swizzled): one TMA bulk copy per atom, issued by the thread that
        // issues the MMAs (the MMAs that read the previous panel completed at the end of the previous group)
        if (tid == 0)
          load_w_panel(sW, a.Wt + (((size_t)k * gridDim.y + blockIdx.y) * a.n_atoms + a0) * ((size_t)a.NT * kAtomBytes), ga, a.NT, &s_wbar);
        resident_k = k;
      }
      fence_proxy_async_smem();   // generic-proxy smem writes (the gathered A) -> visible to the tensor core (async proxy)
      __syncthreads();
      if (tid == 0) {
        if (load_w) { mbar_wait_bounded(&s_wbar, wphase); wphase ^= 1u; }
        tc_fence_after();
        for (int at = 0; at < ga; at++) {
          const int ch0 = (a0 + at) * KD::kElemsPerAtom;
          const int ksteps = (min(KD::kElemsPerAtom, a.c_in - ch0) + KD::kUmmaK - 1) / KD::kUmmaK;
          for (int ks = 0; ks < ksteps; ks++) {
            // advancing K inside the 128 B swizzle row = advancing the start address by 32 B per step
            const uint64_t ad = umma_desc_k128(sA + (uint32_t)at * (kTileM * kAtomBytes) + (uint32_t)ks * 32);
            const uint64_t bd = umma_desc_k128(sW + (uint32_t)at * (a.NT * kAtomBytes) + (uint32_t)ks * 32);
            umma<KIND>(tmem, ad, bd, idesc, (a0 + at + ks) > 0 ? 1u : 0u);
          }
        }
        umma_commit(&s_bar);      // implies tcgen05.fence::before_thread_sync
      }
      mbar_wait(&s_bar, phase);   // MMAs of this group done: smem reusable, accumulator readable
      phase ^= 1u;
    }
    tc_fence_after();

    // ---- scatter: thread t owns accumulator row t (TMEM lane 32*warp + lane) ----
    float *orow = out_row >= 0 ? a.out + (int64_t)out_row * a.ld_out : nullptr;
    for (int c0 = 0; c0 < a.NT; c0 += 16) {
      float v[16];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
      if (orow != nullptr) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const int col = n0 + c0 + j;
          if (col + 3 < a.c_out) {
            red_add_v4(orow + col, v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
#pragma unroll
            for (int e = 0; e < 4; e++)
              if (col + e < a.c_out) atomicAdd(orow + col + e, v[j + e]);
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();              // accumulator drained and s_inrow consumed before the next tile overwrites them
    tc_fence_after();
  }

  if (warp == 0) tmem_dealloc(tmem, (uint32_t)a.tmem_cols);
}

// ---- the pipelined (warp-specialised, persistent) tile kernel -----------------------------------------------------------
// Same math as spconv_fgms_tc_kernel, organised as three concurrent roles per CTA (one CTA per SM, a contiguous run of
// tiles each) so that the gather of tile i+2, the MMA of tile i+1 and the scatter of tile i overlap:
//   warps 0-3  gather   cp.async (LDGSTS) 16 B per lane straight into the swizzled K-major stage ring — no data
//                       registers, DEPTH tiles in flight per thread; rows beyond the pair list / channels beyond c_in
//                       are zero-filled by cp.async's src-size operand.  For tf32 the fp32 input is copied verbatim
//                       (the tensor core reads the top 19 bits); for bf16 the features are converted once per call.
//                       After wait_group: fence.proxy.async, then arrive on full[stage].
//   warp 8     MMA      one lane: wait full[stage] and tmem_empty[buf], issue the tcgen05.mma chain into one of two
//                       TMEM accumulators, commit to empty[stage] (stage reusable) and tmem_full[buf].
//   warps 4-7  scatter  wait tmem_full[buf], tcgen05.ld the 32 rows of their TMEM lane quadrant, red.global.add.v4
//                       into out[omap[p], :], arrive on tmem_empty[buf].
// W[k]^T stays resident in shared memory; when the kernel offset changes inside a CTA's run the gather warps drain the
// outstanding MMAs (wait on the empty barriers) and reload it.
constexpr int kPipeThreads = 288;
constexpr int kPipeMaxStages = 4;
constexpr int kStgCols = 32;   // accumulator columns staged per pass of the scatter

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void *src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

// Walks consecutive tiles: kernel offset k and the pair range of each tile.  One binary search at the first tile, then
// k only moves forward (a couple of dependent loads per offset change instead of a search per tile).
struct TileCursor {
  int k, q_lo, q_hi, kp, kpe;   // qkpos[k], qkpos[k+1], kpos[k], kpos[k+1]
  int pbase, pend;
  bool identity;
  __device__ __forceinline__ void init(const SpconvArgs &a, int tile) {
    if (tile < a.n_map_tiles) {
      k = upper_bound_i32(a.qkpos, a.k_vol + 1, tile * kTileM) - 1;
      q_lo = __ldg(a.qkpos + k); q_hi = __ldg(a.qkpos + k + 1);
      kp = __ldg(a.kpos + k); kpe = __ldg(a.kpos + k + 1);
    } else {
      k = a.k_vol; q_lo = q_hi = kp = kpe = 0;
    }
  }
  __device__ __forceinline__ void seek(const SpconvArgs &a, int tile) {   // tile must not decrease between calls
    if (tile < a.n_map_tiles) {
      const int q0 = tile * kTileM;
      while (q0 >= q_hi) {
        k++;
        q_lo = q_hi; q_hi = __ldg(a.qkpos + k + 1);
        kp = kpe; kpe = __ldg(a.kpos + k + 1);
      }
      pbase = q0 - q_lo + kp; pend = kpe; identity = false;
    } else {
      pbase = (tile - a.n_map_tiles) * kTileM; pend = a.id_rows; identity = true;
    }
  }
  __device__ __forceinline__ int offset(const SpconvArgs &a) const { return identity ? a.mid_k : k; }
};

template <int KIND, int DEPTH>
__global__ void __launch_bounds__(kPipeThreads, 1) spconv_fgms_pipe_kernel(const SpconvArgs a) {
  using KD = Kind<KIND>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t s_full[kPipeMaxStages], s_empty[kPipeMaxStages], s_tfull[2], s_tempty[2], s_wbar;
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = a.n_stages;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_bytes = (uint32_t)a.n_atoms * kTileM * kAtomBytes;
  const uint32_t sW = base + (uint32_t)S * stage_bytes;   // [n_atoms][NT rows][128 B]
  const uint32_t sStage = sW + (uint32_t)a.n_atoms * a.NT * kAtomBytes;   // [4 warps][32 rows][32 + 4] fp32 scatter staging
  const int n0 = blockIdx.y * a.NT;

  if (tid == 0) {
    for (int s = 0; s < S; s++) { mbar_init(&s_full[s], 128); mbar_init(&s_empty[s], 1); }
    for (int t = 0; t < 2; t++) { mbar_init(&s_tfull[t], 1); mbar_init(&s_tempty[t], 128); }
    mbar_init(&s_wbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) tmem_alloc(&s_tmem, (uint32_t)(2 * a.tmem_cols));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  const int n_tiles = a.n_map_tiles + a.n_id_tiles;
  const int t_begin = blockIdx.x * a.tiles_per_cta;
  const int n_my = max(0, min(n_tiles, t_begin + a.tiles_per_cta) - t_begin);

  if (warp < 4) {
    // ================= gather =================
    const int cpr = a.n_atoms * 8;                 // 16-byte chunks per row
    int resident_k = -1;
    uint32_t wphase = 0;      // parity of the weight-panel barrier
    bool w_pending = false;   // a panel copy is in flight: it must have landed before the next full[] arrival
    // full[] tells the MMA warp "this tile's A stage AND the resident W panel are in shared memory"
    auto arrive_full = [&](int stage) {
      if (w_pending) { mbar_wait_bounded(&s_wbar, wphase); wphase ^= 1u; w_pending = false; }
      mbar_arrive(&s_full[stage]);
    };
    int na = 0;   // next tile whose full[] arrival this thread still owes (its copies may still be in flight)
    TileCursor cur;
    int next_in = -1;   // imap entry of this lane's row of the NEXT tile, loaded one tile ahead
    if (n_my > 0) {
      cur.init(a, t_begin);
      cur.seek(a, t_begin);
      const int p = cur.pbase + warp * 32 + lane;
      if (p < cur.pend) next_in = cur.identity ? p : __ldg(a.imap + p);
    }
    for (int i = 0; i < n_my; i++) {
      const int s = i % S;
      const int my_in = next_in;                   // row 32*warp + lane of this tile gathers input row my_in
      const int tile_k = cur.offset(a);
      if (i + 1 < n_my) {                          // prefetch the map entry of the next tile behind this tile's copies
        cur.seek(a, t_begin + i + 1);
        const int p = cur.pbase + warp * 32 + lane;
        next_in = -1;
        if (p < cur.pend) next_in = cur.identity ? p : __ldg(a.imap + p);
      }
      if (tile_k != resident_k) {
        // The resident W changes: hand over every tile gathered so far, wait until all their MMAs have completed
        // (they read the resident W), then load W[k]^T rows [n0, n0 + NT).
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        fence_proxy_async_smem();
        for (; na < i; na++) arrive_full(na % S);
        for (int j = max(0, i - S); j < i; j++) mbar_wait(&s_empty[j % S], (uint32_t)(j / S) & 1u);
        // The prepared panel IS the shared-memory image (swizzled): thread 0 hands it to the copy engine (one TMA bulk copy
        // per atom) and everybody goes on gathering this tile; the copy is only waited for at the tile's full[] arrival.
        if (tid == 0)
          load_w_panel(sW, a.Wt + ((size_t)tile_k * gridDim.y + blockIdx.y) * ((size_t)a.n_atoms * a.NT * kAtomBytes), a.n_atoms, a.NT,
                       &s_wbar);
        w_pending = true;
        resident_k = tile_k;
      } else if (i >= S) {
        mbar_wait(&s_empty[s], (uint32_t)(i / S - 1) & 1u);   // the MMA that read this stage (tile i - S) is done
      }
      const uint32_t sA = base + (uint32_t)s * stage_bytes;
      // this warp's 32 rows x cpr chunks, consecutive lanes on consecutive chunks of a row (coalesced 16 B each)
      int rl = lane / cpr, cc = lane % cpr;
      const int drl = 32 / cpr, dcc = 32 % cpr;
      for (int it = 0; it < cpr; it++) {
        const int in_row = __shfl_sync(0xffffffffu, my_in, rl);
        const int r = warp * 32 + rl, at = cc >> 3, c = cc & 7;
        const bool ok = in_row >= 0 && cc * 16 < a.feat_valid_bytes;
        const uint32_t dst = sA + (uint32_t)at * (kTileM * kAtomBytes) + (uint32_t)r * kAtomBytes + (uint32_t)((c ^ (r & 7)) << 4);
        cp_async16_zfill(dst, ok ? a.feat + (size_t)in_row * a.feat_pitch + cc * 16 : a.feat, ok ? 16u : 0u);
        rl += drl; cc += dcc;
        if (cc >= cpr) { cc -= cpr; rl += 1; }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (i - na >= DEPTH) {   // the oldest owed tile has at most DEPTH younger groups behind it: wait for it only
        asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH) : "memory");
        fence_proxy_async_smem();
        arrive_full(na % S);
        na++;
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    fence_proxy_async_smem();
    for (; na < n_my; na++) arrive_full(na % S);
  } else if (warp == 8) {
    // ================= MMA issue =================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc<KIND>(a.NT, a.f16);
      for (int i = 0; i < n_my; i++) {
        const int s = i % S, t = i & 1;
        mbar_wait(&s_full[s], (uint32_t)(i / S) & 1u);
        if (i >= 2) mbar_wait(&s_tempty[t], (uint32_t)(i / 2 - 1) & 1u);
        tc_fence_after();
        const uint32_t sA = base + (uint32_t)s * stage_bytes;
        const uint32_t d_tmem = tmem + (uint32_t)(t * a.tmem_cols);
        for (int at = 0; at < a.n_atoms; at++) {
          const int ch0 = at * KD::kElemsPerAtom;
          const int ksteps = (min(KD::kElemsPerAtom, a.c_in - ch0) + KD::kUmmaK - 1) / KD::kUmmaK;
          for (int ks = 0; ks < ksteps; ks++) {
            const uint64_t ad = umma_desc_k128(sA + (uint32_t)at * (kTileM * kAtomBytes) + (uint32_t)ks * 32);
            const uint64_t bd = umma_desc_k128(sW + (uint32_t)at * (a.NT * kAtomBytes) + (uint32_t)ks * 32);
            umma<KIND>(d_tmem, ad, bd, idesc, (at + ks) > 0 ? 1u : 0u);
          }
        }
        umma_commit(&s_empty[s]);
        umma_commit(&s_tfull[t]);
      }
    }
  } else {
    // ================= scatter =================
    const int q = warp - 4;   // == warp % 4: the TMEM lane quadrant this warp may read
    constexpr int pitch = kStgCols + 4;           // floats per staged row (+16 B: conflict-free 128-bit rows)
    float *stg = reinterpret_cast<float *>(smem_raw + (sStage - smem_u32(smem_raw))) + (size_t)q * 32 * pitch;
    TileCursor cur;
    int next_out = -1;        // omap entry of this lane's row of the NEXT tile, loaded one tile ahead
    if (n_my > 0) {
      cur.init(a, t_begin);
      cur.seek(a, t_begin);
      const int p = cur.pbase + q * 32 + lane;
      if (p < cur.pend) next_out = cur.identity ? p : __ldg(a.omap + p);
    }
    for (int i = 0; i < n_my; i++) {
      const int t = i & 1;
      const int out_row = next_out;
      if (i + 1 < n_my) {
        cur.seek(a, t_begin + i + 1);
        const int p = cur.pbase + q * 32 + lane;
        next_out = -1;
        if (p < cur.pend) next_out = cur.identity ? p : __ldg(a.omap + p);
      }
      mbar_wait(&s_tfull[t], (uint32_t)(i / 2) & 1u);
      tc_fence_after();
      // TMEM -> registers (thread = row) -> this warp's staging tile in shared memory, so that the scatter below
      // issues whole contiguous row segments per instruction (full 32 B sectors at the L2 atomic units instead of
      // 32 lanes x 16 B of 32 different rows)
      const uint32_t taddr = tmem + (uint32_t)(t * a.tmem_cols) + ((uint32_t)(q * 32) << 16);
      for (int cb = 0; cb < a.NT; cb += kStgCols) {   // column blocks of <= 32: bounds the staging tile
        const int cw = min(kStgCols, a.NT - cb), cpn = cw / 4;
        for (int c0 = 0; c0 < cw; c0 += 16) {
          float v[16];
          tmem_ld16(taddr + (uint32_t)(cb + c0), v);
          float4 *dst = reinterpret_cast<float4 *>(stg + (size_t)lane * pitch + c0);
#pragma unroll
          for (int j = 0; j < 4; j++) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        if (cb + kStgCols >= a.NT) {   // accumulator fully read: the MMA warp may reuse this TMEM buffer
          tc_fence_before();
          mbar_arrive(&s_tempty[t]);
        }
        __syncwarp();
        int rl = lane / cpn, cc = lane % cpn;
        const int drl = 32 / cpn, dcc = 32 % cpn;
        for (int it = 0; it < cpn; it++) {
          const int orow = __shfl_sync(0xffffffffu, out_row, rl);
          const int col = n0 + cb + cc * 4;
          if (orow >= 0 && col < a.c_out) {
            const float4 x = *reinterpret_cast<const float4 *>(stg + (size_t)rl * pitch + cc * 4);
            red_add_v4(a.out + (int64_t)orow * a.ld_out + col, x.x, x.y, x.z, x.w);
          }
          rl += drl; cc += dcc;
          if (cc >= cpn) { cc -= cpn; rl += 1; }
        }
        __syncwarp();              // staging tile free for the next block / accumulator
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, (uint32_t)(2 * a.tmem_cols));
}

// fp32 features -> bf16 / fp16 rows of k_pad elements (zero padded), once per call, for the pipelined 16-bit gather
__global__ void __launch_bounds__(256) spconv_to_bf16_kernel(const float *__restrict__ in, int64_t ld_in, int rows, int c_in,
                                                             int k_pad, __nv_bfloat16 *__restrict__ out, int f16) {
  const int chunks = k_pad / 8;
  const int64_t total = (int64_t)rows * chunks;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / chunks), ch = (int)(i % chunks) * 8;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; e++) v[e] = ch + e < c_in ? __ldg(in + (int64_t)r * ld_in + ch + e) : 0.0f;
    uint4 o = make_uint4(pack_16(v[0], v[1], f16), pack_16(v[2], v[3], f16), pack_16(v[4], v[5], f16), pack_16(v[6], v[7], f16));
    *reinterpret_cast<uint4 *>(out + (int64_t)r * k_pad + ch) = o;
  }
}

// ---- W preparation: W[k][c * sc + n * sn] converted to the MMA dtype, zero padded, laid out as the SHARED-MEMORY IMAGE the
// MMA kernels use for a (kernel offset, NT-row tile) panel: [k][n / NT][atom][n % NT][128 B], the eight 16-byte chunks of a
// row XOR-swizzled with the row (SWIZZLE_128B, K-major).  A panel is then one contiguous block per atom and the kernels
// fetch it with TMA bulk copies instead of staging it through registers.
template <int KIND>
__global__ void __launch_bounds__(256) spconv_prep_weights_kernel(const float *__restrict__ W, uint8_t *__restrict__ Wt, int k_vol,
                                                                  int kdim, int ndim, int k_pad, int n_rows_pad, int NT, int64_t sc,
                                                                  int64_t sn, int64_t sk, int f16) {
  constexpr int ES = (KIND == 0) ? 4 : 2;
  const int64_t total = (int64_t)k_vol * n_rows_pad * k_pad;
  const int n_atoms = k_pad * ES / kAtomBytes, n_tiles = n_rows_pad / NT;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % k_pad);
    const int n = (int)((i / k_pad) % n_rows_pad);
    const int k = (int)(i / ((int64_t)k_pad * n_rows_pad));
    float v = 0.0f;
    if (c < kdim && n < ndim) v = __ldg(W + (int64_t)k * sk + (int64_t)c * sc + (int64_t)n * sn);
    const int byte = c * ES, at = byte / kAtomBytes, c16 = (byte % kAtomBytes) >> 4, within = byte & 15;
    const int nt = n / NT, nl = n % NT;
    const size_t off = ((((size_t)k * n_tiles + nt) * n_atoms + at) * NT + nl) * kAtomBytes + (size_t)(((c16 ^ (nl & 7)) << 4) + within);
    if (KIND == 0) *reinterpret_cast<uint32_t *>(Wt + off) = to_tf32(v);
    else if (f16) *reinterpret_cast<__half *>(Wt + off) = __float2half_rn(v);
    else *reinterpret_cast<__nv_bfloat16 *>(Wt + off) = __float2bfloat16_rn(v);
  }
}

// ---- exact fp32 SIMT path (the reference's arch80 = false kernels, _fgms_fusion_fp32*) --------------------------------
// One warp per group of 4 pairs of the same offset; lane l owns output channels l, l + 32, ... (<= 8 per lane, i.e.
// c_out <= 256 per pass); W rows are read once per 4 pairs.  Products are accumulated ci-ascending like cpu_compute.
constexpr int kSimtPairs = 4;
constexpr int kSimtCols = 8;
__global__ void __launch_bounds__(256) spconv_fgms_simt_kernel(const SpconvArgs a, const float *__restrict__ W, int64_t w_sc,
                                                               int64_t w_sn, int64_t w_sk) {
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int groups_per_tile = kTileM / kSimtPairs;
  const int n_tiles = a.n_map_tiles + a.n_id_tiles;
  const int tile = warp_global / groups_per_tile, g = warp_global % groups_per_tile;
  if (tile >= n_tiles) return;
  int k, in_row[kSimtPairs], out_row[kSimtPairs];
  if (tile < a.n_map_tiles) {
    const int q0 = tile * kTileM;
    k = upper_bound_i32(a.qkpos, a.k_vol + 1, q0) - 1;
    const int p0 = q0 - __ldg(a.qkpos + k) + __ldg(a.kpos + k) + g * kSimtPairs, pe = __ldg(a.kpos + k + 1);
#pragma unroll
    for (int i = 0; i < kSimtPairs; i++) {
      in_row[i] = p0 + i < pe ? __ldg(a.imap + p0 + i) : -1;
      out_row[i] = p0 + i < pe ? __ldg(a.omap + p0 + i) : -1;
    }
  } else {
    k = a.mid_k;
    const int r0 = (tile - a.n_map_tiles) * kTileM + g * kSimtPairs;
#pragma unroll
    for (int i = 0; i < kSimtPairs; i++) in_row[i] = out_row[i] = r0 + i < a.id_rows ? r0 + i : -1;
  }
  if (in_row[0] < 0) return;
  const float *Wk = W + (int64_t)k * w_sk;
  for (int nb = 0; nb < a.c_out; nb += 32 * kSimtCols) {
    float acc[kSimtPairs][kSimtCols];
#pragma unroll
    for (int i = 0; i < kSimtPairs; i++)
#pragma unroll
      for (int j = 0; j < kSimtCols; j++) acc[i][j] = 0.0f;
    for (int ci = 0; ci < a.c_in; ci++) {
      float x[kSimtPairs], w[kSimtCols];
#pragma unroll
      for (int i = 0; i < kSimtPairs; i++) x[i] = in_row[i] >= 0 ? __ldg(a.in + (int64_t)in_row[i] * a.ld_in + ci) : 0.0f;
#pragma unroll
      for (int j = 0; j < kSimtCols; j++) {
        const int co = nb + j * 32 + lane;
        w[j] = co < a.c_out ? __ldg(Wk + (int64_t)ci * w_sc + (int64_t)co * w_sn) : 0.0f;
      }
#pragma unroll
      for (int i = 0; i < kSimtPairs; i++)
#pragma unroll
        for (int j = 0; j < kSimtCols; j++) acc[i][j] = fmaf(x[i], w[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < kSimtPairs; i++) {
      if (out_row[i] < 0) continue;
#pragma unroll
      for (int j = 0; j < kSimtCols; j++) {
        const int co = nb + j * 32 + lane;
        if (co < a.c_out) atomicAdd(a.out + (int64_t)out_row[i] * a.ld_out + co, acc[i][j]);
      }
    }
  }
}

// ---- exact fp32, tiled: the path arch80 = false takes whenever rows are 16-byte aligned ---------------------------------
// The warp-per-4-pairs kernel above re-reads W[k] for every 4 pairs and runs at ~3 TFLOP/s (0.82 ms on the MinkUNet 64 -> 64
// layer, where the reference's _fgms_fusion_fp32 takes 0.10 ms on the same B200: profiles/r02_spconv_vs_reference.jsonl).
// This one is a plain register-blocked SGEMM on gathered tiles: a CTA of 256 threads owns one tile of 128 pairs (one offset)
// x 64 (or 32) output channels and walks c_in in chunks of 32: the gathered rows go to shared memory TRANSPOSED (channel-major, so
// that a thread reads its 8 pairs of one channel with two 16-byte loads), the W[k] chunk row-major; thread (ty, tx)
// accumulates an 8 x 4 block with fp32 FMAs, c_in ascending (the order of cpu_compute), and the block is scattered with
// red.global.add.v4.f32.  32 FMAs per 3 shared-memory loads; several CTAs per SM hide the gather.
// RT = pairs per thread: 8 -> 16 x 16 threads own 128 pairs x 64 channels; 4 -> 32 x 8 threads 128 x 32; 2 -> 64 x 4 threads 128 x 16 (layers
// with c_out <= 32 / 16, where the 64-wide block multiplied zeros for half of its columns: 32 -> 32 forward 56 us against 45 us for the
// reference's kernel).
constexpr int kFtK = 32, kFtLdA = kTileM + 4;
template <int RT>
__global__ void __launch_bounds__(256) spconv_fgms_fp32_tiled_kernel(const SpconvArgs a, const float *__restrict__ W, int64_t w_sc,
                                                                     int64_t w_sn, int64_t w_sk, int out_vec4) {
  __shared__ __align__(16) float sA[kFtK][kFtLdA];
  constexpr int TXN = 2 * RT, kFtN = TXN * 4;      // RT = 8 | 4 | 2 pairs per thread -> 16 | 8 | 4 thread columns of 4 channels = 64 | 32 | 16
  __shared__ __align__(16) float sW[kFtK][kFtN];
  __shared__ int s_in[kTileM], s_out[kTileM];
  const int tid = threadIdx.x, tx = tid % TXN, ty = tid / TXN;
  const int tile = blockIdx.x, co0 = blockIdx.y * kFtN;
  int k;
  if (tile < a.n_map_tiles) {
    const int q0 = tile * kTileM;
    k = upper_bound_i32(a.qkpos, a.k_vol + 1, q0) - 1;
    const int p0 = q0 - __ldg(a.qkpos + k) + __ldg(a.kpos + k), pe = __ldg(a.kpos + k + 1);
    if (tid < kTileM) {
      const bool ok = p0 + tid < pe;
      s_in[tid] = ok ? __ldg(a.imap + p0 + tid) : -1;
      s_out[tid] = ok ? __ldg(a.omap + p0 + tid) : -1;
    }
  } else {
    k = a.mid_k;
    const int r0 = (tile - a.n_map_tiles) * kTileM;
    if (tid < kTileM) s_in[tid] = s_out[tid] = r0 + tid < a.id_rows ? r0 + tid : -1;
  }
  __syncthreads();
  const float *Wk = W + (int64_t)k * w_sk;
  float acc[RT][4];
#pragma unroll
  for (int i = 0; i < RT; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.0f;

  const int g_row = tid & (kTileM - 1), g_half = tid >> 7;   // gather: thread -> (pair, 16 of the chunk's 32 channels)
  const int my_in = s_in[g_row];
  for (int c0 = 0; c0 < a.c_in; c0 += kFtK) {
    // A chunk, transposed into shared memory (rows of 128 consecutive pairs: conflict-free stores)
#pragma unroll
    for (int v = 0; v < 4; v++) {
      const int ch = c0 + g_half * 16 + v * 4;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (my_in >= 0 && ch < a.c_in) x = __ldg(reinterpret_cast<const float4 *>(a.in + (int64_t)my_in * a.ld_in + ch));   // c_in % 4 == 0
      const int cl = g_half * 16 + v * 4;
      sA[cl + 0][g_row] = x.x; sA[cl + 1][g_row] = x.y; sA[cl + 2][g_row] = x.z; sA[cl + 3][g_row] = x.w;
    }
    // W chunk [32][64] (any strides: the dX pass reads W transposed)
#pragma unroll
    for (int v = 0; v < (kFtK * kFtN) / 256; v++) {
      const int e = v * 256 + tid, ci = e / kFtN, co = e % kFtN;
      float w = 0.0f;
      if (c0 + ci < a.c_in && co0 + co < a.c_out) w = __ldg(Wk + (int64_t)(c0 + ci) * w_sc + (int64_t)(co0 + co) * w_sn);
      sW[ci][co] = w;
    }
    __syncthreads();
#pragma unroll 8
    for (int ci = 0; ci < kFtK; ci++) {
      float av[RT];
      if (RT == 2) {
        const float2 t = *reinterpret_cast<const float2 *>(&sA[ci][ty * RT]);
        av[0] = t.x; av[1] = t.y;
      } else {
#pragma unroll
        for (int h = 0; h < RT / 4; h++) {
          const float4 t = *reinterpret_cast<const float4 *>(&sA[ci][ty * RT + 4 * h]);
          av[4 * h] = t.x; av[4 * h + 1] = t.y; av[4 * h + 2] = t.z; av[4 * h + 3] = t.w;
        }
      }
      const float4 w = *reinterpret_cast<const float4 *>(&sW[ci][tx * 4]);
      const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < RT; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const int co = co0 + tx * 4;
  if (co >= a.c_out) return;
#pragma unroll
  for (int i = 0; i < RT; i++) {
    const int orow = s_out[ty * RT + i];
    if (orow < 0) continue;
    float *dst = a.out + (int64_t)orow * a.ld_out + co;
    if (out_vec4) red_add_v4(dst, acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    else {
#pragma unroll
      for (int j = 0; j < 4; j++)
        if (co + j < a.c_out) atomicAdd(dst + j, acc[i][j]);
    }
  }
}

// ---- kernel gradient: dW[k][ci][co] = sum over the pairs p of offset k of in[imap[p]][ci] * dout[omap[p]][co] ----------
// (_fgms_fusion_tf32_I_transpose of the reference, src/cuda/spconv_cuda.cu:241-247.)  A CTA of 16 x 16 threads owns a
// 64 x 64 block of dW and a run of pair tiles; per tile it stages the gathered 128 x 64 panels of `in` and `dout` in
// shared memory, every thread accumulates a 4 x 4 register block over the 128 pairs, and the block is flushed with
// atomics only when the offset changes (or the run ends).  Exact fp32 FMA.
constexpr int kWgBlk = 64;
__global__ void __launch_bounds__(256) spconv_wgrad_kernel(const SpconvArgs a, const float *__restrict__ dout, int64_t ld_dout,
                                                           float *__restrict__ dW) {
  extern __shared__ uint8_t smem_raw[];
  float *sX = reinterpret_cast<float *>(smem_raw);          // [128][64]
  float *sG = sX + kTileM * kWgBlk;                         // [128][64]
  __shared__ int s_in[kTileM], s_out[kTileM];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int ci0 = blockIdx.y * kWgBlk, co0 = blockIdx.z * kWgBlk;
  const int n_tiles = a.n_map_tiles + a.n_id_tiles;
  const int t_begin = blockIdx.x * a.tiles_per_cta, t_end = min(n_tiles, t_begin + a.tiles_per_cta);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.0f;
  int cur_k = -1;

  auto flush = [&](int k) {
    float *wk = dW + (int64_t)k * a.c_in * a.c_out;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int ci = ci0 + ty * 4 + i;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int co = co0 + tx * 4 + j;
        if (ci < a.c_in && co < a.c_out && acc[i][j] != 0.0f) atomicAdd(wk + (int64_t)ci * a.c_out + co, acc[i][j]);
        acc[i][j] = 0.0f;
      }
    }
  };

  for (int tile = t_begin; tile < t_end; tile++) {
    int k;
    if (tile < a.n_map_tiles) {
      const int q0 = tile * kTileM;
      k = upper_bound_i32(a.qkpos, a.k_vol + 1, q0) - 1;
      if (tid < kTileM) {
        const int p = q0 - __ldg(a.qkpos + k) + __ldg(a.kpos + k) + tid;
        const bool ok = p < __ldg(a.kpos + k + 1);
        s_in[tid] = ok ? __ldg(a.imap + p) : -1;
        s_out[tid] = ok ? __ldg(a.omap + p) : -1;
      }
    } else {
      k = a.mid_k;
      if (tid < kTileM) {
        const int r = (tile - a.n_map_tiles) * kTileM + tid;
        s_in[tid] = s_out[tid] = r < a.id_rows ? r : -1;
      }
    }
    if (k != cur_k) {
      if (cur_k >= 0) flush(cur_k);
      cur_k = k;
    }
    __syncthreads();
    // stage the two panels: 16 lanes x float (coalesced 64-float rows), zero for padding rows / channels
    for (int idx = tid; idx < kTileM * kWgBlk; idx += 256) {
      const int r = idx >> 6, c = idx & 63;
      const int ir = s_in[r], orow = s_out[r];
      sX[idx] = (ir >= 0 && ci0 + c < a.c_in) ? __ldg(a.in + (int64_t)ir * a.ld_in + ci0 + c) : 0.0f;
      sG[idx] = (orow >= 0 && co0 + c < a.c_out) ? __ldg(dout + (int64_t)orow * ld_dout + co0 + c) : 0.0f;
    }
    __syncthreads();
#pragma unroll 4
    for (int p = 0; p < kTileM; p++) {
      const float4 x = *reinterpret_cast<const float4 *>(sX + p * kWgBlk + ty * 4);
      const float4 g = *reinterpret_cast<const float4 *>(sG + p * kWgBlk + tx * 4);
      const float xs[4] = {x.x, x.y, x.z, x.w}, gs[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(xs[i], gs[j], acc[i][j]);
    }
    __syncthreads();
  }
  if (cur_k >= 0) flush(cur_k);
}

// ---- kernel gradient on the tensor cores (tf32) ----------------------------------------------------------------------
// dW[k] = X_k^T . G_k where X_k / G_k are the gathered input / output-gradient rows of the pairs of offset k: a GEMM whose
// contraction axis is the PAIR axis.  The gathered 128-pair tiles are written to shared memory exactly as in the forward
// ([32-channel block][128 pairs][128 B], but with the SWIZZLE_128B_BASE32B pattern) and handed to tcgen05.mma as MN-major
// operands: the 128 B row of a pair is 32 consecutive M (or N) elements, eight pair rows are one K = 8 instruction, LBO =
// the 16 KB between channel blocks, SBO = 512 B between 4-row groups.  16 MMAs consume a tile; the fp32 accumulator (c_in block x c_out block)
// stays in TMEM across ALL tiles of an offset and is flushed with red.global.add.v4 only when the offset changes or the
// CTA's run ends.  Warps 0-3 gather (cp.async, zero-fill for padding pairs), warp 4 issues the MMAs; the gather warps also
// do the rare flush.  M is always 128 (c_in blocks narrower than 128 are zero-filled by the gather), so accumulator row r
// lives in TMEM lane r.
constexpr int kWgtMaxStages = 3;

// MN-major descriptor for 32-bit operands: SWIZZLE_128B_BASE32B (layout type 1) is the only MN-major layout tcgen05 accepts
// for tf32 (cute: Layout_MN_SW128_32B_Atom = Swizzle<2,5,2> over [4 K rows][128 B]): the 32-byte chunk index of a row is
// XORed with (K row & 3); LBO = bytes between 32-element MN blocks, SBO = 512 B between 4-row K groups.
__device__ __forceinline__ uint64_t umma_desc_mn128_b32(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)(512 >> 4) << 32) |
         (1ull << 46) | (1ull << 61);
}

struct WgradTcArgs {
  SpconvArgs a;
  const float *dout;
  int64_t ld_dout;
  float *dW;
  int MB, NBk;        // rows (c_in) and columns (c_out) of dW handled by this CTA's blockIdx.y / .z: 128, <= 128
  int n_stages;
  int m64_mode;       // experiment: TMEM row mapping hypothesis for M = 64 (1: lane = row, 2: 16 rows per quadrant)
};

// 8 gather warps (16 pairs of a tile each) + 1 MMA warp: the gather is an instruction stream of ~70 instructions per 16-byte
// chunk-row, and with 4 gather warps (one per scheduler) it alone set the pace: 107 us for every channel count up to 64.
constexpr int kWgtGatherWarps = 16;
constexpr int kWgtThreads = (kWgtGatherWarps + 1) * 32;
constexpr int kWgtRows = kTileM / kWgtGatherWarps;   // pairs of a tile per gather warp
__global__ void __launch_bounds__(kWgtThreads, 1) spconv_wgrad_tc_kernel(const WgradTcArgs g) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t s_full[kWgtMaxStages], s_empty[kWgtMaxStages], s_accfull, s_accempty;
  __shared__ uint32_t s_tmem;
  const SpconvArgs &a = g.a;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = g.n_stages;
  const int mblk = g.MB / 32, nblk = (g.NBk + 31) / 32;          // 32-channel blocks of the two operands
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t blk_bytes = kTileM * kAtomBytes;                 // 16 KB: [128 pairs][128 B]
  const uint32_t stage_bytes = (uint32_t)(mblk + nblk) * blk_bytes;
  const int ci0 = blockIdx.y * g.MB, co0 = blockIdx.z * g.NBk;
  const int tmem_cols = g.NBk <= 32 ? 32 : g.NBk <= 64 ? 64 : 128;

  if (tid == 0) {
    for (int s = 0; s < S; s++) { mbar_init(&s_full[s], kWgtGatherWarps * 32); mbar_init(&s_empty[s], 1); }
    mbar_init(&s_accfull, 1); mbar_init(&s_accempty, 128);   // the flush is done by warps 0-3 (one TMEM lane quadrant each)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kWgtGatherWarps) tmem_alloc(&s_tmem, (uint32_t)tmem_cols);
  // Channels beyond c_in / c_out inside the MB x NBk block are zero in EVERY tile: clear the stages once and never touch
  // those chunks again (MB is always 128: with c_in = 64 that is half of the X chunks, with c_in = 4 31 of 32 — the gather
  // warps are instruction bound, ncu: 3 300 instructions per warp and tile, one warp per scheduler)
  const int x_chunks = min(mblk * 8, (a.c_in - ci0 + 3) / 4), g_chunks = min(nblk * 8, (a.c_out - co0 + 3) / 4);
  if (x_chunks < mblk * 8 || g_chunks < nblk * 8) {
    for (uint32_t o = (uint32_t)tid * 16u; o < (uint32_t)S * stage_bytes; o += (uint32_t)kWgtThreads * 16u)
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(base + o), "r"(0u) : "memory");
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  const int n_tiles = a.n_map_tiles + a.n_id_tiles;
  const int t_begin = blockIdx.x * a.tiles_per_cta;
  const int n_my = max(0, min(n_tiles, t_begin + a.tiles_per_cta) - t_begin);

  if (warp < kWgtGatherWarps) {
    // ================= gather (+ the rare flush) =================
    // live 16-byte chunks per pair: X chunks first, then G chunks (c_in, c_out are multiples of 4 on this path)
    const int cpr = x_chunks + g_chunks;
    const int items = kWgtRows * cpr;    // (pair, chunk) items of this warp per tile
    TileCursor cur;
    int next_in = -1, next_out = -1;
    if (n_my > 0) {
      cur.init(a, t_begin);
      cur.seek(a, t_begin);
      const int p = cur.pbase + warp * kWgtRows + (lane % kWgtRows);
      if (p < cur.pend) { next_in = cur.identity ? p : __ldg(a.imap + p); next_out = cur.identity ? p : __ldg(a.omap + p); }
    }
    int na = 0, cur_k = -1, n_flush = 0;

    auto flush = [&](int k, int upto) {
      // hand over every gathered tile, wait for the MMA warp's commit of the accumulator, add it into dW[k]
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      fence_proxy_async_smem();
      for (; na < upto; na++) mbar_arrive(&s_full[na % S]);
      if (warp >= 4) { n_flush++; return; }      // TMEM lane quadrants belong to warps 0-3
      mbar_wait(&s_accfull, (uint32_t)n_flush & 1u);
      tc_fence_after();
      int row = warp * 32 + lane;                // accumulator row (c_in index within the block) = TMEM lane
      bool has_row = true;
      if (g.MB == 64) {
        if (g.m64_mode == 2) { row = warp * 16 + lane; has_row = lane < 16; }
        else { has_row = warp < 2; }
      }
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
      float *wrow = g.dW + ((int64_t)k * a.c_in + ci0 + row) * a.c_out + co0;
      for (int c0 = 0; c0 < g.NBk; c0 += 16) {
        float v[16];
        tmem_ld16(taddr + (uint32_t)c0, v);
        if (has_row && ci0 + row < a.c_in) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            if (co0 + c0 + j + 3 < a.c_out) red_add_v4(wrow + c0 + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
            else {
#pragma unroll
              for (int e = 0; e < 4; e++)
                if (co0 + c0 + j + e < a.c_out) atomicAdd(wrow + c0 + j + e, v[j + e]);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&s_accempty);
      n_flush++;
    };

    for (int i = 0; i < n_my; i++) {
      const int s = i % S;
      const int my_in = next_in, my_out = next_out;
      const int tile_k = cur.offset(a);
      if (i + 1 < n_my) {
        cur.seek(a, t_begin + i + 1);
        const int p = cur.pbase + warp * kWgtRows + (lane % kWgtRows);
        next_in = next_out = -1;
        if (p < cur.pend) { next_in = cur.identity ? p : __ldg(a.imap + p); next_out = cur.identity ? p : __ldg(a.omap + p); }
      }
      if (tile_k != cur_k) {
        if (cur_k >= 0) flush(cur_k, i);
        cur_k = tile_k;
      }
      if (i >= S) mbar_wait(&s_empty[s], (uint32_t)(i / S - 1) & 1u);
      const uint32_t sX = base + (uint32_t)s * stage_bytes;
      int rl = lane / cpr, cc = lane % cpr;
      const int drl = 32 / cpr, dcc = 32 % cpr;
      for (int it = 0; it * 32 < items; it++) {
        const bool live = it * 32 + lane < items;    // an odd chunk count leaves half of the last round without work
        if (!live) rl = 0;
        const int r = warp * kWgtRows + rl;
        const bool is_x = cc < x_chunks;
        const int idx = is_x ? cc : cc - x_chunks;
        const int blk = idx >> 3, c = idx & 7;
        const int in_r = __shfl_sync(0xffffffffu, my_in, rl), out_r = __shfl_sync(0xffffffffu, my_out, rl);
        const int src_row = is_x ? in_r : out_r;   // (the select must happen in the READING lane)
        const int ch = (is_x ? ci0 : co0) + blk * 32 + c * 4;
        const bool ok = src_row >= 0;
        const float *src = is_x ? a.in + (int64_t)src_row * a.ld_in + ch : g.dout + (int64_t)src_row * g.ld_dout + ch;
        const uint32_t dst = sX + (uint32_t)((is_x ? 0 : mblk) + blk) * blk_bytes + (uint32_t)r * kAtomBytes +
                             (uint32_t)((((c >> 1) ^ (r & 3)) << 5) | ((c & 1) << 4));   // BASE32B swizzle
        if (live) cp_async16_zfill(dst, ok ? (const void *)src : (const void *)a.in, ok ? 16u : 0u);
        rl += drl; cc += dcc;
        if (cc >= cpr) { cc -= cpr; rl += 1; }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (i - na >= 1) {
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        fence_proxy_async_smem();
        mbar_arrive(&s_full[na % S]);
        na++;
      }
    }
    if (cur_k >= 0) flush(cur_k, n_my);
  } else {
    // ================= MMA issue =================
    if (lane == 0) {
      // D fp32, A/B tf32, both MN-major (bits 15, 16), N >> 3, M >> 4
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(g.NBk >> 3) << 17) |
                             ((uint32_t)(g.MB >> 4) << 24);
      TileCursor cur;
      if (n_my > 0) cur.init(a, t_begin);
      int cur_k = -1, n_flush = 0;
      bool first = true;
      for (int i = 0; i < n_my; i++) {
        const int s = i % S;
        cur.seek(a, t_begin + i);
        const int tile_k = cur.offset(a);
        if (tile_k != cur_k) {
          if (cur_k >= 0) {
            umma_commit(&s_accfull);                                   // accumulator of the previous offset complete
            mbar_wait(&s_accempty, (uint32_t)n_flush & 1u);            // ... and drained by the flush
            n_flush++;
          }
          cur_k = tile_k;
          first = true;
        }
        mbar_wait(&s_full[s], (uint32_t)(i / S) & 1u);
        tc_fence_after();
        const uint32_t sX = base + (uint32_t)s * stage_bytes, sG = sX + (uint32_t)mblk * blk_bytes;
        for (int ks = 0; ks < kTileM / 8; ks++) {
          const uint64_t ad = umma_desc_mn128_b32(sX + (uint32_t)ks * 1024u, blk_bytes);
          const uint64_t bd = umma_desc_mn128_b32(sG + (uint32_t)ks * 1024u, blk_bytes);
          umma<0>(tmem, ad, bd, idesc, (first && ks == 0) ? 0u : 1u);
        }
        first = false;
        umma_commit(&s_empty[s]);
      }
      if (cur_k >= 0) umma_commit(&s_accfull);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kWgtGatherWarps) tmem_dealloc(tmem, (uint32_t)tmem_cols);
}

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

struct TcGeometry {
  int n_atoms, k_pad, NT, grid_y, n_rows_pad, tmem_cols, esize;
  size_t smem_bytes;
};

TcGeometry tc_geometry(int kdim, int ndim, int precision) {
  TcGeometry g;
  const bool b16 = precision == SPCONV_BF16 || precision == SPCONV_FP16;
  const int epa = b16 ? 64 : 32;
  g.esize = b16 ? 2 : 4;
  g.n_atoms = (kdim + epa - 1) / epa;
  g.k_pad = g.n_atoms * epa;
  const int n16 = round_up(ndim, 16);
  g.NT = n16 <= kMaxNT ? n16 : kMaxNT;
  g.grid_y = (n16 + g.NT - 1) / g.NT;
  g.n_rows_pad = g.grid_y * g.NT;
  g.tmem_cols = 32;
  while (g.tmem_cols < g.NT) g.tmem_cols *= 2;
  const int ga = g.n_atoms < kMaxAtomsPerGroup ? g.n_atoms : kMaxAtomsPerGroup;
  g.smem_bytes = 1024 + (size_t)ga * kTileM * kAtomBytes + (size_t)ga * g.NT * kAtomBytes;
  return g;
}

}  // namespace

size_t spconv_workspace_bytes(int rows, int k_vol, int c_in, int c_out, int precision) {
  if (precision == SPCONV_FP32) return 256;
  // large enough for the forward (K = c_in, N = c_out) and the dX backward (K = c_out, N = c_in):
  // the prepared weights, plus (bf16) a converted copy of the gathered feature matrix of `rows` rows
  const TcGeometry f = tc_geometry(c_in, c_out, precision), b = tc_geometry(c_out, c_in, precision);
  const size_t wf = (size_t)k_vol * f.n_rows_pad * f.k_pad * f.esize, wb = (size_t)k_vol * b.n_rows_pad * b.k_pad * b.esize;
  size_t need = (wf > wb ? wf : wb) + 512;
  if (precision == SPCONV_BF16 || precision == SPCONV_FP16) need += (size_t)(rows > 0 ? rows : 0) * (f.k_pad > b.k_pad ? f.k_pad : b.k_pad) * 2 + 256;
  return need;
}

cudaError_t spconv_gemm(const SpconvProblem &p, void *workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (p.k_vol <= 0 || p.kdim <= 0 || p.ndim <= 0) return cudaErrorInvalidValue;
  if (p.precision != SPCONV_FP32 && p.precision != SPCONV_TF32 && p.precision != SPCONV_BF16 && p.precision != SPCONV_FP16)
    return cudaErrorInvalidValue;
  if (p.sum_nnz % kTileM != 0) return cudaErrorInvalidValue;   // qkpos must be quantised to 128 (q of the reference)
  cudaError_t e;
  if (!p.accumulate && p.out_rows > 0) {
    if (p.ld_out == p.ndim) {
      if ((e = cudaMemsetAsync(p.out, 0, sizeof(float) * (size_t)p.out_rows * p.ndim, stream)) != cudaSuccess) return e;
    } else if ((e = cudaMemset2DAsync(p.out, sizeof(float) * p.ld_out, 0, sizeof(float) * p.ndim, p.out_rows, stream)) != cudaSuccess) {
      return e;
    }
  }
  SpconvArgs a;
  a.kpos = p.kpos; a.qkpos = p.qkpos; a.imap = p.imap; a.omap = p.omap;
  a.in = p.in; a.out = p.out; a.ld_in = p.ld_in; a.ld_out = p.ld_out;
  a.k_vol = p.k_vol; a.c_in = p.kdim; a.c_out = p.ndim;
  a.f16 = p.precision == SPCONV_FP16 ? 1 : 0;
  a.n_map_tiles = (int)(p.sum_nnz / kTileM);
  a.mid_k = (p.k_vol % 2 == 1) ? p.k_vol / 2 : 0;   // src/cuda/spconv_cuda.cu:35
  a.id_rows = p.separate_mid ? p.in_rows : 0;
  a.n_id_tiles = (a.id_rows + kTileM - 1) / kTileM;
  const int n_tiles = a.n_map_tiles + a.n_id_tiles;
  if (n_tiles == 0) return cudaSuccess;
  ProfileScope prof(5, stream);

  // tensor path needs 16-byte aligned rows; otherwise (or when asked) the exact SIMT kernel
  const bool tc_ok = p.kdim % 4 == 0 && p.ndim % 4 == 0 && p.ld_in % 4 == 0 && p.ld_out % 4 == 0 &&
                     (reinterpret_cast<uintptr_t>(p.in) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0;
  if (p.precision == SPCONV_FP32 || !tc_ok) {
    a.Wt = nullptr; a.n_atoms = 0; a.n_rows_pad = 0; a.NT = 0; a.tmem_cols = 0; a.tiles_per_cta = 1;
    // register-blocked tiled kernel when the gathered rows can be read as float4; the warp-per-4-pairs kernel otherwise
    const bool in_vec4 = p.kdim % 4 == 0 && p.ld_in % 4 == 0 && (reinterpret_cast<uintptr_t>(p.in) & 15) == 0;
    if (in_vec4 && !getenv("DGS_SPCONV_FP32_SIMPLE")) {
      const int out_vec4 = p.ndim % 4 == 0 && p.ld_out % 4 == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0;
      if (p.ndim <= 16) {
        dim3 grid(n_tiles, 1);
        spconv_fgms_fp32_tiled_kernel<2><<<grid, 256, 0, stream>>>(a, p.W, p.w_sc, p.w_sn, p.w_sk, out_vec4);
      } else if (p.ndim <= 32) {
        dim3 grid(n_tiles, 1);
        spconv_fgms_fp32_tiled_kernel<4><<<grid, 256, 0, stream>>>(a, p.W, p.w_sc, p.w_sn, p.w_sk, out_vec4);
      } else {
        dim3 grid(n_tiles, (p.ndim + 63) / 64);
        spconv_fgms_fp32_tiled_kernel<8><<<grid, 256, 0, stream>>>(a, p.W, p.w_sc, p.w_sn, p.w_sk, out_vec4);
      }
      return cudaGetLastError();
    }
    const int64_t warps = (int64_t)n_tiles * (kTileM / kSimtPairs);
    const int blocks = (int)((warps * 32 + 255) / 256);
    spconv_fgms_simt_kernel<<<blocks, 256, 0, stream>>>(a, p.W, p.w_sc, p.w_sn, p.w_sk);
    return cudaGetLastError();
  }

  const TcGeometry g = tc_geometry(p.kdim, p.ndim, p.precision);
  const size_t w_bytes = (size_t)p.k_vol * g.n_rows_pad * g.k_pad * g.esize;
  if (workspace == nullptr || workspace_bytes < w_bytes + 256) return cudaErrorInvalidValue;
  uint8_t *Wt = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
  const int64_t total = (int64_t)p.k_vol * g.n_rows_pad * g.k_pad;
  const int prep_blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  if (p.precision == SPCONV_TF32)
    spconv_prep_weights_kernel<0><<<prep_blocks, 256, 0, stream>>>(p.W, Wt, p.k_vol, p.kdim, p.ndim, g.k_pad, g.n_rows_pad, g.NT, p.w_sc, p.w_sn, p.w_sk, a.f16);
  else
    spconv_prep_weights_kernel<1><<<prep_blocks, 256, 0, stream>>>(p.W, Wt, p.k_vol, p.kdim, p.ndim, g.k_pad, g.n_rows_pad, g.NT, p.w_sc, p.w_sn, p.w_sk, a.f16);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;

  a.Wt = Wt; a.n_atoms = g.n_atoms; a.n_rows_pad = g.n_rows_pad; a.NT = g.NT; a.tmem_cols = g.tmem_cols;
  a.feat = nullptr; a.feat_pitch = 0; a.feat_valid_bytes = 0; a.n_stages = 0;

  // pipelined kernel when at least two whole-K stages and the W panel fit in shared memory
  {
    const size_t stage_bytes = (size_t)g.n_atoms * kTileM * kAtomBytes, w_smem = (size_t)g.n_atoms * g.NT * kAtomBytes;
    const size_t stg_smem = (size_t)4 * 32 * (kStgCols + 4) * 4;   // scatter staging of the four epilogue warps
    // Persistent CTAs per SM: as many (<= 3: 288 threads x 64 registers) as still leave every CTA two whole-K stages, its W panel and two TMEM accumulators
    // (2 x tmem_cols of the SM's 512 columns).  One CTA per SM leaves the scatter — 4 warps issuing red.global.add.v4 —
    // as the bottleneck: MinkUNet 64 -> 64 tf32 forward 58.4 us with one CTA per SM, 44.0 us with two.
    int per_sm = 1, S = 0;
    for (int c = 3; c >= 1; c--) {
      if (option(OPT_SPCONV_CTAS) >= 1 && option(OPT_SPCONV_CTAS) <= 3 && c != option(OPT_SPCONV_CTAS)) continue;
      const size_t budget = (size_t)(225 / c) * 1024u - (c > 1 ? 1024u : 0u);   // 1 KB per extra CTA is reserved by the driver
      const int Sc = w_smem + stg_smem + 1024 < budget ? (int)((budget - w_smem - stg_smem - 1024) / stage_bytes) : 0;
      if ((Sc >= 2 && c * 2 * g.tmem_cols <= 512) || c == 1) { per_sm = c; S = Sc; break; }
    }
    if (S > kPipeMaxStages) S = kPipeMaxStages;
    bool pipe = S >= 2 && !getenv("DGS_SPCONV_NO_PIPE");
    if (pipe && p.precision != SPCONV_TF32) {   // converted (bf16 / fp16) feature copy lives behind the weights in the workspace
      uint8_t *feat = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(Wt + w_bytes) + 255) & ~(uintptr_t)255);
      const size_t f_bytes = (size_t)p.in_rows * g.k_pad * 2;
      if (feat + f_bytes > reinterpret_cast<uint8_t *>(workspace) + workspace_bytes) {
        pipe = false;
      } else if (p.in_rows > 0) {
        const int64_t chunks = (int64_t)p.in_rows * (g.k_pad / 8);
        const int cb = (int)((chunks + 255) / 256 < 8192 ? (chunks + 255) / 256 : 8192);
        spconv_to_bf16_kernel<<<cb, 256, 0, stream>>>(p.in, p.ld_in, p.in_rows, p.kdim, g.k_pad, reinterpret_cast<__nv_bfloat16 *>(feat), a.f16);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        a.feat = feat; a.feat_pitch = (int64_t)g.k_pad * 2; a.feat_valid_bytes = g.k_pad * 2;
      }
    } else if (pipe) {
      a.feat = reinterpret_cast<const uint8_t *>(p.in); a.feat_pitch = p.ld_in * 4; a.feat_valid_bytes = p.kdim * 4;
    }
    if (pipe && a.feat != nullptr) {
      a.n_stages = S;
      int ctas = device_sm_count() * per_sm / g.grid_y;
      if (ctas < 1) ctas = 1;
      a.tiles_per_cta = (n_tiles + ctas - 1) / ctas;
      dim3 grid((n_tiles + a.tiles_per_cta - 1) / a.tiles_per_cta, g.grid_y);
      const size_t smem = 1024 + (size_t)S * stage_bytes + w_smem + stg_smem;
#define DGS_PIPE(KIND_, DEPTH_)                                                                                             \
  do {                                                                                                                      \
    if ((e = cudaFuncSetAttribute(spconv_fgms_pipe_kernel<KIND_, DEPTH_>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                  (int)smem)) != cudaSuccess) return e;                                                     \
    spconv_fgms_pipe_kernel<KIND_, DEPTH_><<<grid, kPipeThreads, smem, stream>>>(a);                                        \
  } while (0)
      if (p.precision == SPCONV_TF32) { if (S >= 3) DGS_PIPE(0, 2); else DGS_PIPE(0, 1); }
      else { if (S >= 3) DGS_PIPE(1, 2); else DGS_PIPE(1, 1); }
#undef DGS_PIPE
      return cudaGetLastError();
    }
  }
  // a run of tiles per CTA amortises the W[k] panel; keep >= ~8 CTAs per SM in flight for latency hiding
  const int sms = device_sm_count();
  int tpc = n_tiles / (sms * 8);
  if (tpc < 1) tpc = 1;
  if (tpc > 8) tpc = 8;
  a.tiles_per_cta = tpc;
  dim3 grid((n_tiles + tpc - 1) / tpc, g.grid_y);
  if (p.precision == SPCONV_TF32) {
    if ((e = cudaFuncSetAttribute(spconv_fgms_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes)) != cudaSuccess) return e;
    spconv_fgms_tc_kernel<0><<<grid, 128, g.smem_bytes, stream>>>(a);
  } else {
    if ((e = cudaFuncSetAttribute(spconv_fgms_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes)) != cudaSuccess) return e;
    spconv_fgms_tc_kernel<1><<<grid, 128, g.smem_bytes, stream>>>(a);
  }
  return cudaGetLastError();
}

cudaError_t spconv_wgrad(int k_vol, int c_in, int c_out, const int *kpos, const int *qkpos, const int *imap, const int *omap,
                         int64_t sum_nnz, const float *in, int64_t ld_in, int in_rows, const float *dout, int64_t ld_dout,
                         float *dW, int precision, int separate_mid, cudaStream_t stream) {
  if (k_vol <= 0 || c_in <= 0 || c_out <= 0 || sum_nnz % kTileM != 0) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)k_vol * c_in * c_out, stream);
  if (e != cudaSuccess) return e;
  SpconvArgs a{};
  a.kpos = kpos; a.qkpos = qkpos; a.imap = imap; a.omap = omap; a.in = in; a.ld_in = ld_in;
  a.k_vol = k_vol; a.c_in = c_in; a.c_out = c_out;
  a.n_map_tiles = (int)(sum_nnz / kTileM);
  a.mid_k = (k_vol % 2 == 1) ? k_vol / 2 : 0;
  a.id_rows = separate_mid ? in_rows : 0;
  a.n_id_tiles = (a.id_rows + kTileM - 1) / kTileM;
  const int n_tiles = a.n_map_tiles + a.n_id_tiles;
  if (n_tiles == 0) return cudaSuccess;
  ProfileScope prof(6, stream);
  // tensor path (tf32 operands, fp32 accumulation in TMEM) for the tf32 / bf16 precisions when rows are 16-byte aligned
  const bool tc_ok = precision != SPCONV_FP32 && c_in % 4 == 0 && c_out % 4 == 0 && ld_in % 4 == 0 && ld_dout % 4 == 0 &&
                     (reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(dout) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(dW) & 15) == 0 && !getenv("DGS_SPCONV_WGRAD_SIMT");
  if (tc_ok) {
    WgradTcArgs g;
    g.a = a; g.dout = dout; g.ld_dout = ld_dout; g.dW = dW;
    g.MB = 128;   // always M = 128 (rows beyond c_in are zero-filled): one accumulator row per TMEM lane
    g.m64_mode = 0;
    if (const char *e = getenv("DGS_WGRAD_M64")) { g.m64_mode = atoi(e); if (g.m64_mode && c_in <= 64) g.MB = 64; }
    const int n16 = round_up(c_out, 16);
    g.NBk = n16 <= 128 ? n16 : 128;
    size_t stage_bytes = (size_t)(g.MB / 32 + (g.NBk + 31) / 32) * kTileM * kAtomBytes;
    if (2 * stage_bytes > 220u * 1024u && g.NBk > 64) {
      // a 128 x 128 block of dW needs 128 KB per stage: halve the c_out block so that two stages fit (the X panel is then
      // gathered once per c_out block; 128x128 channels 0.56 ms with the FMA kernel)
      g.NBk = 64;
      stage_bytes = (size_t)(g.MB / 32 + 2) * kTileM * kAtomBytes;
    }
    const int by = (c_in + g.MB - 1) / g.MB, bz = (c_out + g.NBk - 1) / g.NBk;
    int S = (int)((220u * 1024u) / stage_bytes);
    if (S > kWgtMaxStages) S = kWgtMaxStages;
    if (S >= 2) {
      g.n_stages = S;
      int ctas = device_sm_count() / (by * bz);
      if (ctas < 1) ctas = 1;
      g.a.tiles_per_cta = (n_tiles + ctas - 1) / ctas;
      const size_t smem = 1024 + (size_t)S * stage_bytes;
      if ((e = cudaFuncSetAttribute(spconv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
      dim3 grid((n_tiles + g.a.tiles_per_cta - 1) / g.a.tiles_per_cta, by, bz);
      spconv_wgrad_tc_kernel<<<grid, kWgtThreads, smem, stream>>>(g);
      return cudaGetLastError();
    }
  }
  const int by = (c_in + kWgBlk - 1) / kWgBlk, bz = (c_out + kWgBlk - 1) / kWgBlk;
  int tpc = n_tiles * by * bz / (device_sm_count() * 6);
  if (tpc < 1) tpc = 1;
  if (tpc > 16) tpc = 16;
  a.tiles_per_cta = tpc;
  const size_t smem = sizeof(float) * 2 * kTileM * kWgBlk;
  if ((e = cudaFuncSetAttribute(spconv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
  dim3 grid((n_tiles + tpc - 1) / tpc, by, bz);
  spconv_wgrad_kernel<<<grid, 256, smem, stream>>>(a, dout, ld_dout, dW);
  return cudaGetLastError();
}

}  // namespace dgs
