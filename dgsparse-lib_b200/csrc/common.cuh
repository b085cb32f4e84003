// common.cuh — shared device helpers for the sm_100a kernels (torch-free).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dgs {

// Op algebra of the reference: include/gspmm.h:13-14 (REDUCEOP / COMPUTEOP enum order is part of
// the gspmm-fp pybind surface, src/gspmm-fp/gspmm.cc:31-42).  C_COPY = "no edge value" kernels.
enum ReduceOp { R_SUM = 0, R_MAX = 1, R_MIN = 2, R_MEAN = 3 };
enum ComputeOp { C_ADD = 0, C_SUB = 1, C_MUL = 2, C_DIV = 3, C_COPY = 4, C_MASK = 5 };  // C_MASK: multiply, masked by the forward's arg index

constexpr int kMaxDst = 8;  // output fan-out (local C + NVLink peers) of the fused column-shard epilogue

// MAX/MIN identities exactly as the reference: (float)INT_MIN / (float)INT_MAX, include/gspmm.h:133-146.
template <int RED> __device__ __forceinline__ float reduce_identity() {
  if (RED == R_MAX) return -2147483648.0f;
  if (RED == R_MIN) return 2147483648.0f;
  return 0.0f;
}

// COMPUTE::compute(a = edge value, b = feature), src/gspmm-fp/gspmm.h:53-79 (Sub = b - a, Div = b / a).
template <int COMP> __device__ __forceinline__ float compute_op(float a, float b) {
  if (COMP == C_ADD) return a + b;
  if (COMP == C_SUB) return b - a;
  if (COMP == C_MUL || COMP == C_MASK) return a * b;
  if (COMP == C_DIV) return b / a;
  return b;
}

// smallest idx in [0, n) with a[idx] > x (a ascending, a[n-1] > x guaranteed by the caller)
__device__ __forceinline__ int upper_bound_i32(const int *__restrict__ a, int n, int x) {
  int lo = 0, hi = n - 1;  // answer in [lo, hi]
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (__ldg(a + mid) > x) hi = mid; else lo = mid + 1;
  }
  return lo;
}

// Row that owns nnz position p: rowptr[r] <= p < rowptr[r+1]; skips empty rows
// (same contract as findRow / binary_search_segment_number, src/util/cuda_util.cuh:53-89,167-174).
__device__ __forceinline__ int row_of_nnz(const int *__restrict__ rowptr, int M, int p) {
  return upper_bound_i32(rowptr, M + 1, p) - 1;
}

// The same, searching FORWARD from a row r0 known to start at or before p (rowptr[r0] <= p): one probe decides between a
// 32-row window and the rest of the matrix, then ONE bisection loop (the same code shape as the full search — a galloping
// version with two loops made the SpMM hot loop 20 % slower on a graph that never takes this path).  Rows advance by a
// handful at a time inside a segment, so this is ~6 loads where the full search is log2(M) dependent ones (on
// p2p-Gnutella31, 74 % empty rows, the full search was most of the kernel).
__device__ __forceinline__ int row_of_nnz_from(const int *__restrict__ rowptr, int M, int p, int r0) {
  int lo = r0 + 1, hi = min(M, r0 + 32);          // smallest idx in (r0, M] with rowptr[idx] > p
  if (__ldg(rowptr + hi) <= p) { lo = hi + 1; hi = M; }
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(rowptr + mid) > p) hi = mid; else lo = mid + 1;
  }
  return lo - 1;
}

// base + row * stride_bytes as ONE mad.wide.u32 (row strides are < 4 GiB; the product may exceed 32 bits)
__device__ __forceinline__ const char *row_addr(const char *base, unsigned row, unsigned stride_bytes) {
  unsigned long long out;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(out) : "r"(row), "r"(stride_bytes), "l"((unsigned long long)base));
  return reinterpret_cast<const char *>(out);
}

template <int VEC> struct VecT;
template <> struct VecT<1> { using f = float;  using i = int;  };
template <> struct VecT<2> { using f = float2; using i = int2; };
template <> struct VecT<4> { using f = float4; using i = int4; };

template <int VEC> __device__ __forceinline__ void ld_vec(float (&d)[VEC], const float *p) {
  if (VEC == 4) { float4 t = __ldg(reinterpret_cast<const float4 *>(p)); d[0] = t.x; d[1] = t.y; d[2] = t.z; d[3] = t.w; }
  else if (VEC == 2) { float2 t = __ldg(reinterpret_cast<const float2 *>(p)); d[0] = t.x; d[1] = t.y; }
  else d[0] = __ldg(p);
}

template <int VEC> __device__ __forceinline__ void ld_ivec(int (&d)[VEC], const int *p) {
  if (VEC == 4) { int4 t = __ldg(reinterpret_cast<const int4 *>(p)); d[0] = t.x; d[1] = t.y; d[2] = t.z; d[3] = t.w; }
  else if (VEC == 2) { int2 t = __ldg(reinterpret_cast<const int2 *>(p)); d[0] = t.x; d[1] = t.y; }
  else d[0] = __ldg(p);
}

template <int VEC> __device__ __forceinline__ void st_vec_cs(float *p, const float (&s)[VEC]) {
  if (VEC == 4) __stcs(reinterpret_cast<float4 *>(p), make_float4(s[0], s[1], s[2], s[3]));
  else if (VEC == 2) __stcs(reinterpret_cast<float2 *>(p), make_float2(s[0], s[1]));
  else __stcs(p, s[0]);
}

// One store to an NVLS multicast address: the NVSwitch replicates it into the same offset of every GPU bound to the
// multicast object (this one included), so a column-shard epilogue sends each finished row ONCE instead of once per peer.
template <int VEC> __device__ __forceinline__ void st_vec_multimem(float *p, const float (&s)[VEC]) {
  if (VEC == 4)
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(s[0]), "f"(s[1]), "f"(s[2]), "f"(s[3]) : "memory");
  else {
#pragma unroll
    for (int v = 0; v < VEC; v++) asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p + v), "f"(s[v]) : "memory");
  }
}

template <int VEC> __device__ __forceinline__ void st_vec_cs(int *p, const int (&s)[VEC]) {
  if (VEC == 4) __stcs(reinterpret_cast<int4 *>(p), make_int4(s[0], s[1], s[2], s[3]));
  else if (VEC == 2) __stcs(reinterpret_cast<int2 *>(p), make_int2(s[0], s[1]));
  else __stcs(p, s[0]);
}

}  // namespace dgs
