// kmap.cu — kernel-map construction for sparse 3-D convolution (the step before the gather-GEMM-scatter).
//
// Replaces (behaviour, not code) sparse_mapping of the reference, src/cuda/sparse_mapping.cu:20-161, and its kernels
// in include/cuda/sparse_mapping.cuh: coordsDownsample / coordsGenerator (:296-432), insertHash / insertVal (:27-63),
// _queryhash_subm / _queryhash_sp (:68-229), exclusive_scan_for_kernel_quantified.  The reference never registers this
// as an op, leaves the result as a dense input-major table map[k][in] = out with the pair counters bumped by atomics,
// and mixes two coordinate conventions between the down-sampler and the strided query; what is built here is the
// well-defined product the spconv op consumes (the MinkUNet fixtures' layout): pair lists (imap, omap) grouped by
// kernel offset, kpos / qkpos, DETERMINISTIC (within an offset the pairs are ordered by output index).
//
// Semantics (offset decode order and centring exactly as _queryhash_subm / _queryhash_sp):
//   tap kp = (kx * ksy + ky) * ksz + kz, kx in [0, ksx) ...;  off(k) = k - (ks-1)/2 (+1 for odd ks > 1)
//   submanifold:             out_coords = in_coords;        input = out + (tap - (ks - 1) / 2)
//   plain down-sampling (stride = 1 or ks per axis, padding 0):
//                            out_coords = sorted unique of floor(in / s) (batch kept);   input = out * s + off(tap)
//   general layer (coordsDownsampleExpand, any stride / padding / bounds):
//                            out_coords = sorted unique of (in - off(tap) + pad) / s over exact divisions inside [lo, hi];
//                            input = out * s - pad + off(tap)
// Integer, HBM / latency bound: one 64-bit key per coordinate (16 bits per component, biased), an open-addressing
// hash table of the input keys (2x over-provisioned, linear probing, atomicCAS), one thread per (offset, output) query
// writing a k-major hit table, an exclusive scan of the hit flags (cub::DeviceScan: position = final pair index, so the
// order is deterministic and kpos falls out of the scan), and one compaction pass.  Sort / unique / scan are CUB
// (toolkit library code, as the reference uses thrust for the same steps); hash insert, query and compaction are ours.
#include <cub/cub.cuh>
#include <cstdint>
#include "common.cuh"
#include "kmap.h"

namespace dgs {
namespace {

constexpr unsigned long long kEmptyKey = ~0ull;
constexpr int kBias = 1 << 15;

__device__ __forceinline__ unsigned long long pack_key(int b, int x, int y, int z) {
  return ((unsigned long long)(unsigned)(b & 0xffff) << 48) | ((unsigned long long)(unsigned)((x + kBias) & 0xffff) << 32) |
         ((unsigned long long)(unsigned)((y + kBias) & 0xffff) << 16) | (unsigned long long)(unsigned)((z + kBias) & 0xffff);
}
__device__ __forceinline__ unsigned hash_key(unsigned long long k) {   // 64 -> 32 bit mix (murmur3 finaliser)
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return (unsigned)k;
}
__device__ __forceinline__ int floor_div(int a, int s) { return (a >= 0) ? a / s : -((-a + s - 1) / s); }
// Offset of tap k along one axis of a strided layer: the decode shared by coordsDownsampleExpand and _queryhash_sp
// (include/cuda/sparse_mapping.cuh:365-370, :195-200): k - (ks-1)/2, plus 1 for odd ks > 1  ->  k for ks <= 3,
// k - 1 for ks = 4 and 5, k - 2 for ks = 6 and 7 ...
__device__ __forceinline__ int tap_offset(int k, int ks) { return k - (ks - 1) / 2 + ((ks % 2 == 0 || ks == 1) ? 0 : 1); }

struct ExpandArgs {
  int in_nnz, ksx, ksy, ksz, k_vol, sx, sy, sz, px, py, pz;
  int lo[3], hi[3];
  const int *in_coords;
  unsigned long long *keys;   // [in_nnz * k_vol]
};

// coordsDownsampleExpand (include/cuda/sparse_mapping.cuh:326-401): every (input, tap) proposes the output voxel
// (in - tap + padding) / stride when that division is exact and the result lies inside [lo, hi]; everything else
// becomes the empty key (sorts last).  Per-axis padding (the reference reads padding[0] for all three axes).
__global__ void __launch_bounds__(256) expand_keys_kernel(const ExpandArgs a) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)a.in_nnz * a.k_vol) return;
  const int i = (int)(t / a.k_vol), k = (int)(t % a.k_vol);
  const int4 v = __ldg(reinterpret_cast<const int4 *>(a.in_coords) + i);
  const int cx = v.y - tap_offset(k / (a.ksz * a.ksy), a.ksx) + a.px;
  const int cy = v.z - tap_offset((k / a.ksz) % a.ksy, a.ksy) + a.py;
  const int cz = v.w - tap_offset(k % a.ksz, a.ksz) + a.pz;
  unsigned long long key = kEmptyKey;
  if (cx % a.sx == 0 && cy % a.sy == 0 && cz % a.sz == 0) {
    const int ox = cx / a.sx, oy = cy / a.sy, oz = cz / a.sz;
    if (ox >= a.lo[0] && ox <= a.hi[0] && oy >= a.lo[1] && oy <= a.hi[1] && oz >= a.lo[2] && oz <= a.hi[2])
      key = pack_key(v.x, ox, oy, oz);
  }
  a.keys[t] = key;
}

// the empty key, if any candidate was rejected, is the last unique key: drop it from the count
__global__ void drop_empty_key_kernel(int *n_dev, const unsigned long long *keys) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const int n = *n_dev;
    if (n > 0 && keys[n - 1] == kEmptyKey) *n_dev = n - 1;
  }
}

__global__ void __launch_bounds__(256) downsample_keys_kernel(int n, const int *__restrict__ c, int sx, int sy, int sz,
                                                              unsigned long long *__restrict__ keys) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int4 v = __ldg(reinterpret_cast<const int4 *>(c) + i);
  keys[i] = pack_key(v.x, floor_div(v.y, sx), floor_div(v.z, sy), floor_div(v.w, sz));
}

__global__ void __launch_bounds__(256) unpack_keys_kernel(const int *__restrict__ n_dev, const unsigned long long *__restrict__ keys,
                                                          int *__restrict__ c) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *n_dev) return;
  const unsigned long long k = keys[i];
  reinterpret_cast<int4 *>(c)[i] = make_int4((int)(k >> 48), (int)((k >> 32) & 0xffff) - kBias, (int)((k >> 16) & 0xffff) - kBias,
                                             (int)(k & 0xffff) - kBias);
}

__global__ void __launch_bounds__(256) fill_keys_kernel(unsigned long long *t, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) t[i] = kEmptyKey;
}

// A key holds 16 bits per component: batch in [0, 65535], x / y / z in (-32768, 32768).  Anything else would alias
// another voxel, so it raises *bad instead (kpos_kernel then poisons knnz / kpos / qkpos with -1).
__device__ __forceinline__ bool key_in_range(const int4 v) {
  return v.x >= 0 && v.x <= 0xffff && v.y > -kBias && v.y < kBias && v.z > -kBias && v.z < kBias && v.w > -kBias && v.w < kBias;
}

// table_keys[slot] = key, table_val[slot] = input index.  Duplicate input coordinates keep the smallest index.
__global__ void __launch_bounds__(256) hash_insert_kernel(int n, const int *__restrict__ c, unsigned mask,
                                                          unsigned long long *__restrict__ tkeys, int *__restrict__ tval,
                                                          int *__restrict__ bad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int4 v = __ldg(reinterpret_cast<const int4 *>(c) + i);
  if (!key_in_range(v)) { *bad = 1; return; }
  const unsigned long long key = pack_key(v.x, v.y, v.z, v.w);
  unsigned slot = hash_key(key) & mask;
  while (true) {
    const unsigned long long prev = atomicCAS(tkeys + slot, kEmptyKey, key);
    if (prev == kEmptyKey || prev == key) { atomicMin(tval + slot, i); return; }
    slot = (slot + 1) & mask;
  }
}

struct QueryArgs {
  int out_nnz, ksx, ksy, ksz, k_vol, sx, sy, sz, skip_k, px, py, pz, subm;
  const int *out_coords;
  unsigned mask;
  const unsigned long long *tkeys;
  const int *tval;
  int *hit;    // [k_vol][out_nnz] input index or -1
  int *flag;   // [k_vol][out_nnz] 1 / 0 (scan input)
  int *bad;    // set when an output coordinate does not fit a key
};

// one thread per (offset k, output o), o fastest: coalesced table writes, the 27 probes of one output spread over blocks
__global__ void __launch_bounds__(256) query_kernel(const QueryArgs a) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)a.k_vol * a.out_nnz) return;
  const int k = (int)(t / a.out_nnz), o = (int)(t % a.out_nnz);
  const int4 v = __ldg(reinterpret_cast<const int4 *>(a.out_coords) + o);
  if (!key_in_range(v)) *a.bad = 1;
  int kx = k / (a.ksz * a.ksy), ky = (k / a.ksz) % a.ksy, kz = k % a.ksz;
  int x, y, z;
  if (a.subm) {   // _queryhash_subm: centred taps
    x = v.y + kx - (a.ksx - 1) / 2; y = v.z + ky - (a.ksy - 1) / 2; z = v.w + kz - (a.ksz - 1) / 2;
  } else {        // _queryhash_sp: input = out * stride - padding + tap offset
    x = v.y * a.sx - a.px + tap_offset(kx, a.ksx);
    y = v.z * a.sy - a.py + tap_offset(ky, a.ksy);
    z = v.w * a.sz - a.pz + tap_offset(kz, a.ksz);
  }
  int found = -1;
  if (k != a.skip_k && v.x >= 0 && v.x <= 0xffff && x > -kBias && x < kBias && y > -kBias && y < kBias && z > -kBias && z < kBias) {
    const unsigned long long key = pack_key(v.x, x, y, z);
    unsigned slot = hash_key(key) & a.mask;
    while (true) {
      const unsigned long long tk = __ldg(a.tkeys + slot);
      if (tk == key) { found = __ldg(a.tval + slot); break; }
      if (tk == kEmptyKey) break;
      slot = (slot + 1) & a.mask;
    }
  }
  a.hit[t] = found;
  a.flag[t] = found >= 0 ? 1 : 0;
}

__global__ void __launch_bounds__(256) compact_kernel(int k_vol, int out_nnz, const int *__restrict__ hit, const int *__restrict__ pos,
                                                      int *__restrict__ imap, int *__restrict__ omap) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)k_vol * out_nnz) return;
  const int in = hit[t];
  if (in >= 0) { const int p = pos[t]; imap[p] = in; omap[p] = (int)(t % out_nnz); }
}

// kpos[k] = pos[k * out_nnz]; knnz; qkpos = counts rounded up to q (exclusive_scan_for_kernel_quantified of the reference)
__global__ void kpos_kernel(int k_vol, int out_nnz, const int *__restrict__ pos, const int *__restrict__ flag, int q, int *knnz,
                            int *kpos, int *qkpos, const int *__restrict__ bad) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (*bad) {   // coordinates outside the 16-bit key range: no silent aliasing, the map is marked invalid
    for (int k = 0; k < k_vol; k++) { knnz[k] = -1; kpos[k] = -1; qkpos[k] = -1; }
    kpos[k_vol] = -1; qkpos[k_vol] = -1;
    return;
  }
  const int64_t last = (int64_t)k_vol * out_nnz - 1;
  const int total = last >= 0 ? pos[last] + flag[last] : 0;
  int qacc = 0;
  qkpos[0] = 0;
  for (int k = 0; k < k_vol; k++) {
    const int b = out_nnz > 0 ? pos[(int64_t)k * out_nnz] : 0;
    const int e = (k + 1 < k_vol && out_nnz > 0) ? pos[(int64_t)(k + 1) * out_nnz] : total;
    kpos[k] = b;
    knnz[k] = e - b;
    qacc += (e - b + q - 1) / q * q;
    qkpos[k + 1] = qacc;
  }
  kpos[k_vol] = total;
}

inline size_t up256(size_t x) { return (x + 255) / 256 * 256; }
inline unsigned table_size(int n) {
  unsigned s = 1024;
  while (s < 2u * (unsigned)(n > 0 ? n : 1)) s <<= 1;
  return s;
}

}  // namespace

size_t kmap_workspace_bytes(int in_nnz, int out_nnz, int k_vol) {
  const size_t n = (size_t)(in_nnz > out_nnz ? in_nnz : out_nnz);
  const size_t cells = (size_t)k_vol * (size_t)(out_nnz > 0 ? out_nnz : 1);
  size_t scan_tmp = 0, sort_tmp = 0, uniq_tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, (int *)nullptr, (int *)nullptr, (int)cells);
  cub::DeviceRadixSort::SortKeys(nullptr, sort_tmp, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (int)n);
  cub::DeviceSelect::Unique(nullptr, uniq_tmp, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (int *)nullptr, (int)n);
  size_t tmp = scan_tmp > sort_tmp ? scan_tmp : sort_tmp;
  if (uniq_tmp > tmp) tmp = uniq_tmp;
  const size_t build = up256(table_size(in_nnz) * 8ull) + up256(table_size(in_nnz) * 4ull) + 3 * up256(cells * 4);
  const size_t down = 2 * up256(n * 8);
  return up256(tmp) + (build > down ? build : down) + 1024;
}

cudaError_t kmap_downsample(int in_nnz, const int *in_coords, int sx, int sy, int sz, int *out_coords, int *out_nnz_dev,
                            void *workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (in_nnz < 0 || sx < 1 || sy < 1 || sz < 1) return cudaErrorInvalidValue;
  if (in_nnz == 0) return cudaMemsetAsync(out_nnz_dev, 0, sizeof(int), stream);
  size_t sort_tmp = 0, uniq_tmp = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, sort_tmp, (unsigned long long *)nullptr, (unsigned long long *)nullptr, in_nnz);
  cub::DeviceSelect::Unique(nullptr, uniq_tmp, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (int *)nullptr, in_nnz);
  const size_t tmp = up256(sort_tmp > uniq_tmp ? sort_tmp : uniq_tmp);
  if (workspace == nullptr || workspace_bytes < tmp + 2 * up256((size_t)in_nnz * 8)) return cudaErrorInvalidValue;
  char *w = static_cast<char *>(workspace);
  unsigned long long *k0 = reinterpret_cast<unsigned long long *>(w + tmp);
  unsigned long long *k1 = reinterpret_cast<unsigned long long *>(w + tmp + up256((size_t)in_nnz * 8));
  const int blocks = (in_nnz + 255) / 256;
  downsample_keys_kernel<<<blocks, 256, 0, stream>>>(in_nnz, in_coords, sx, sy, sz, k0);
  size_t t = sort_tmp;
  cudaError_t e = cub::DeviceRadixSort::SortKeys(w, t, k0, k1, in_nnz, 0, 64, stream);   // batch -> x -> y -> z order
  if (e != cudaSuccess) return e;
  t = uniq_tmp;
  if ((e = cub::DeviceSelect::Unique(w, t, k1, k0, out_nnz_dev, in_nnz, stream)) != cudaSuccess) return e;
  unpack_keys_kernel<<<blocks, 256, 0, stream>>>(out_nnz_dev, k0, out_coords);
  return cudaGetLastError();
}

size_t kmap_expand_workspace_bytes(int in_nnz, int k_vol) {
  const size_t n = (size_t)(in_nnz > 0 ? in_nnz : 1) * (size_t)(k_vol > 0 ? k_vol : 1);
  size_t sort_tmp = 0, uniq_tmp = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, sort_tmp, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (int)n);
  cub::DeviceSelect::Unique(nullptr, uniq_tmp, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (int *)nullptr, (int)n);
  return up256(sort_tmp > uniq_tmp ? sort_tmp : uniq_tmp) + 2 * up256(n * 8) + 1024;
}

cudaError_t kmap_downsample_expand(int in_nnz, const int *in_coords, int ksx, int ksy, int ksz, int sx, int sy, int sz, int px,
                                   int py, int pz, const int *lo, const int *hi, int *out_coords, int *out_nnz_dev,
                                   void *workspace, size_t workspace_bytes, cudaStream_t stream) {
  const int k_vol = ksx * ksy * ksz;
  if (in_nnz < 0 || k_vol < 1 || sx < 1 || sy < 1 || sz < 1 || lo == nullptr || hi == nullptr) return cudaErrorInvalidValue;
  if ((int64_t)in_nnz * k_vol > 0x7fffffffLL) return cudaErrorInvalidValue;
  if (in_nnz == 0) return cudaMemsetAsync(out_nnz_dev, 0, sizeof(int), stream);
  const int n = in_nnz * k_vol;
  size_t sort_tmp = 0, uniq_tmp = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, sort_tmp, (unsigned long long *)nullptr, (unsigned long long *)nullptr, n);
  cub::DeviceSelect::Unique(nullptr, uniq_tmp, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (int *)nullptr, n);
  const size_t tmp = up256(sort_tmp > uniq_tmp ? sort_tmp : uniq_tmp);
  if (workspace == nullptr || workspace_bytes < tmp + 2 * up256((size_t)n * 8)) return cudaErrorInvalidValue;
  char *w = static_cast<char *>(workspace);
  unsigned long long *k0 = reinterpret_cast<unsigned long long *>(w + tmp);
  unsigned long long *k1 = reinterpret_cast<unsigned long long *>(w + tmp + up256((size_t)n * 8));
  ExpandArgs a;
  a.in_nnz = in_nnz; a.ksx = ksx; a.ksy = ksy; a.ksz = ksz; a.k_vol = k_vol; a.sx = sx; a.sy = sy; a.sz = sz;
  a.px = px; a.py = py; a.pz = pz;
  for (int d = 0; d < 3; d++) {   // keys hold 16 biased bits per axis
    a.lo[d] = lo[d] > -kBias + 1 ? lo[d] : -kBias + 1;
    a.hi[d] = hi[d] < kBias - 1 ? hi[d] : kBias - 1;
  }
  a.in_coords = in_coords; a.keys = k0;
  const int blocks = (n + 255) / 256;
  expand_keys_kernel<<<blocks, 256, 0, stream>>>(a);
  size_t t = sort_tmp;
  cudaError_t e = cub::DeviceRadixSort::SortKeys(w, t, k0, k1, n, 0, 64, stream);   // batch -> x -> y -> z, empty key last
  if (e != cudaSuccess) return e;
  t = uniq_tmp;
  if ((e = cub::DeviceSelect::Unique(w, t, k1, k0, out_nnz_dev, n, stream)) != cudaSuccess) return e;
  drop_empty_key_kernel<<<1, 32, 0, stream>>>(out_nnz_dev, k0);
  unpack_keys_kernel<<<blocks, 256, 0, stream>>>(out_nnz_dev, k0, out_coords);
  return cudaGetLastError();
}

cudaError_t kmap_build(int in_nnz, const int *in_coords, int out_nnz, const int *out_coords, int ksx, int ksy, int ksz, int sx,
                       int sy, int sz, int q, int skip_mid, int *imap, int *omap, int *knnz, int *kpos, int *qkpos,
                       void *workspace, size_t workspace_bytes, cudaStream_t stream) {
  return kmap_build_ex(in_nnz, in_coords, out_nnz, out_coords, ksx, ksy, ksz, sx, sy, sz, 0, 0, 0,
                       (sx == 1 && sy == 1 && sz == 1) ? 1 : 0, q, skip_mid, imap, omap, knnz, kpos, qkpos, workspace,
                       workspace_bytes, stream);
}

cudaError_t kmap_build_ex(int in_nnz, const int *in_coords, int out_nnz, const int *out_coords, int ksx, int ksy, int ksz, int sx,
                          int sy, int sz, int px, int py, int pz, int subm, int q, int skip_mid, int *imap, int *omap, int *knnz,
                          int *kpos, int *qkpos, void *workspace, size_t workspace_bytes, cudaStream_t stream) {
  const int k_vol = ksx * ksy * ksz;
  if (in_nnz < 0 || out_nnz < 0 || k_vol < 1 || sx < 1 || sy < 1 || sz < 1 || q < 1) return cudaErrorInvalidValue;
  if ((int64_t)k_vol * out_nnz > 0x7fffffffLL) return cudaErrorInvalidValue;
  if (workspace == nullptr || workspace_bytes < kmap_workspace_bytes(in_nnz, out_nnz, k_vol)) return cudaErrorInvalidValue;
  const unsigned tsize = table_size(in_nnz);
  const size_t cells = (size_t)k_vol * out_nnz;
  size_t scan_tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, (int *)nullptr, (int *)nullptr, (int)(cells > 0 ? cells : 1));
  char *w = static_cast<char *>(workspace);
  void *tmp = w; w += up256(scan_tmp);
  unsigned long long *tkeys = reinterpret_cast<unsigned long long *>(w); w += up256(tsize * 8ull);
  int *tval = reinterpret_cast<int *>(w); w += up256(tsize * 4ull);
  int *hit = reinterpret_cast<int *>(w); w += up256((cells > 0 ? cells : 1) * 4);
  int *flag = reinterpret_cast<int *>(w); w += up256((cells > 0 ? cells : 1) * 4);
  int *pos = reinterpret_cast<int *>(w); w += up256((cells > 0 ? cells : 1) * 4);
  int *bad = reinterpret_cast<int *>(w);   // inside the 1 KB tail kmap_workspace_bytes() reserves
  cudaError_t e;
  if ((e = cudaMemsetAsync(bad, 0, sizeof(int), stream)) != cudaSuccess) return e;
  fill_keys_kernel<<<(tsize + 255) / 256, 256, 0, stream>>>(tkeys, (int)tsize);
  if ((e = cudaMemsetAsync(tval, 0x7f, tsize * 4ull, stream)) != cudaSuccess) return e;
  if (in_nnz > 0) hash_insert_kernel<<<(in_nnz + 255) / 256, 256, 0, stream>>>(in_nnz, in_coords, tsize - 1, tkeys, tval, bad);
  if (cells > 0) {
    QueryArgs a;
    // centre tap (k_vol / 2 for odd volumes, 0 otherwise: src/cuda/spconv_cuda.cu:35) left to spconv's separate_mid
    a.skip_k = skip_mid ? ((k_vol % 2 == 1) ? k_vol / 2 : 0) : -1;
    a.out_nnz = out_nnz; a.ksx = ksx; a.ksy = ksy; a.ksz = ksz; a.k_vol = k_vol; a.sx = sx; a.sy = sy; a.sz = sz;
    a.px = px; a.py = py; a.pz = pz; a.subm = subm;
    a.out_coords = out_coords; a.mask = tsize - 1; a.tkeys = tkeys; a.tval = tval; a.hit = hit; a.flag = flag;
    a.bad = bad;
    const int blocks = (int)((cells + 255) / 256);
    query_kernel<<<blocks, 256, 0, stream>>>(a);
    size_t t = scan_tmp;
    if ((e = cub::DeviceScan::ExclusiveSum(tmp, t, flag, pos, (int)cells, stream)) != cudaSuccess) return e;
    compact_kernel<<<blocks, 256, 0, stream>>>(k_vol, out_nnz, hit, pos, imap, omap);
  }
  kpos_kernel<<<1, 32, 0, stream>>>(k_vol, out_nnz, pos, flag, q, knnz, kpos, qkpos, bad);
  return cudaGetLastError();
}

}  // namespace dgs
