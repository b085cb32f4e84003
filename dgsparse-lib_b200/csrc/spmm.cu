// spmm.cu — host dispatcher for the row-segment SpMM (geometry selection, segment sizing, workspace
// carving, the two launches).  Replaces the launch logic of src/cuda/spmm_cuda.cu:14-253,
// src/ge-spmm/gespmm.cc:29-134 and src/gspmm-fp/gspmm.cu:406-473 of the reference.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <mutex>
#include "spmm.h"
#include "options.h"
#include "spmm_rowseg.cuh"
#include "spmm_rowpar.cuh"

namespace dgs {

#define DGS_DECL_LOOKUP(V, G) SpmmKernel spmm_lookup_v##V##_g##G(int red, int comp, bool arg);
DGS_DECL_LOOKUP(4, 4) DGS_DECL_LOOKUP(4, 8) DGS_DECL_LOOKUP(4, 16) DGS_DECL_LOOKUP(4, 32)
DGS_DECL_LOOKUP(1, 4) DGS_DECL_LOOKUP(1, 8) DGS_DECL_LOOKUP(1, 16) DGS_DECL_LOOKUP(1, 32)
SpmmLaunchFn spmm_rowpar_lookup_v4_g4(int red, int comp, bool arg);
SpmmLaunchFn spmm_rowpar_lookup_v4_g8(int red, int comp, bool arg);
SpmmLaunchFn spmm_rowpar_lookup_v4_g16(int red, int comp, bool arg);

// pdl: the fix-up grid is a programmatic dependent of the SpMM grid launched just before it on the same stream — it may
// start while that grid runs (its empty-row half is independent) and waits for it with griddepcontrol.wait.
bool profile_is_on();

template <int RED, bool ARG, int FV>
static cudaError_t launch_fixup_one(const SpmmArgs &a, int blocks, bool pdl, cudaStream_t s) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)blocks);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, spmm_fixup_kernel<RED, ARG, FV>, a);
}

template <bool ARG, int FV> static cudaError_t launch_fixup(int red, const SpmmArgs &a, int blocks, bool pdl, cudaStream_t s) {
  switch (red) {
  case R_MAX: return launch_fixup_one<R_MAX, ARG, FV>(a, blocks, pdl, s);
  case R_MIN: return launch_fixup_one<R_MIN, ARG, FV>(a, blocks, pdl, s);
  default: return launch_fixup_one<R_SUM, false, FV>(a, blocks, pdl, s);
  }
}

// ---- bench-only per-launch timing -----------------------------------------------------------------
namespace {
struct ProfRec { int id; cudaEvent_t a, b; };
bool g_prof_on = false;
std::vector<ProfRec> g_prof;
}  // namespace

bool profile_is_on() { return g_prof_on; }

int profile_enable(bool on) {
  g_prof_on = on;
  if (!on) {
    for (auto &r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_prof.clear();
  }
  return 0;
}

ProfileScope::ProfileScope(int kernel_id, cudaStream_t s) : slot(-1), stream(s) {
  if (!g_prof_on) return;
  ProfRec r;
  r.id = kernel_id;
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
  cudaEventRecord(r.a, s);
  g_prof.push_back(r);
  slot = (int)g_prof.size() - 1;
}

ProfileScope::~ProfileScope() {
  if (slot >= 0) cudaEventRecord(g_prof[slot].b, stream);
}

int profile_collect(int max_records, int *kernel_ids, float *ms) {
  int n = 0;
  for (auto &r : g_prof) {
    cudaEventSynchronize(r.b);
    float t = 0.0f;
    cudaEventElapsedTime(&t, r.a, r.b);
    if (n < max_records) { kernel_ids[n] = r.id; ms[n] = t; n++; }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_prof.clear();
  return n;
}

int device_sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// ---- graph notes: which matrices may take the single-launch row-parallel kernel -------------------------------------
// The row-parallel kernel (spmm_rowpar.cuh) wins in the latency regime but walks a row serially, so it needs to know that
// the matrix has no long row — which neither C ABI tells us.  The library therefore remembers, per (device, rowptr, M),
// what its own kernels saw: the first call on a matrix takes the segment path, whose fix-up grid scans every row anyway
// and sets a mapped host word when one exceeds kRowParLimit; an event marks the end of that scan.  Later calls read the
// word WITHOUT synchronising (cudaEventQuery) and switch to the row-parallel kernel if it stayed clear.  The note is only
// a hint: the row-parallel kernel is correct for any matrix and raises the same word when it meets a long row (a CSR
// rewritten in place under the same pointer), which sends the following calls back to the segment path.
namespace {
constexpr int64_t kRowParMaxNnz = 4 << 20;
struct GraphNote {
  const int *rowptr = nullptr;
  int m = -1, dev = -1;
  int verdict = -1;      // -1 unknown, 0 no row longer than kRowParLimit, 1 has long rows
  bool pending = false;  // a full scan is enqueued; `ev` marks its end
  cudaEvent_t ev = nullptr;
  int *flag_host = nullptr, *flag_dev = nullptr;
};
std::mutex g_note_mu;
GraphNote g_notes[32];
int g_note_clock = 0;

thread_local int g_last_path = 0;   // 0 = row-segment kernel + fix-up, 1 = row-parallel single launch

int rowpar_mode() {   // option spmm_rowpar: 0 = never, 1 = always (tests), unset (-1) = by graph note
  const int v = option(OPT_SPMM_ROWPAR);
  return v < 0 ? -1 : (v ? 1 : 0);
}

// -> index of the note (creating it on first sight), or -1
int note_lookup(const int *rowptr, int m, int *verdict, int **flag_dev, bool *want_scan) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  std::lock_guard<std::mutex> lk(g_note_mu);
  int idx = -1;
  for (int i = 0; i < 32; i++)
    if (g_notes[i].rowptr == rowptr && g_notes[i].m == m && g_notes[i].dev == dev) { idx = i; break; }
  if (idx < 0) {
    idx = g_note_clock++ % 32;
    GraphNote &n = g_notes[idx];
    if (n.flag_host == nullptr) {
      if (cudaHostAlloc((void **)&n.flag_host, sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess ||
          cudaHostGetDevicePointer((void **)&n.flag_dev, n.flag_host, 0) != cudaSuccess ||
          cudaEventCreateWithFlags(&n.ev, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        n.flag_host = nullptr; n.rowptr = nullptr;
        return -1;
      }
    }
    *(volatile int *)n.flag_host = 0;
    n.rowptr = rowptr; n.m = m; n.dev = dev; n.verdict = -1; n.pending = false;
  }
  GraphNote &n = g_notes[idx];
  if (n.pending) {
    const cudaError_t q = cudaEventQuery(n.ev);
    if (q == cudaSuccess) { n.pending = false; n.verdict = *(volatile int *)n.flag_host ? 1 : 0; }
    else cudaGetLastError();   // cudaErrorNotReady is not an error
  } else if (n.verdict == 0 && *(volatile int *)n.flag_host) {
    n.verdict = 1;             // a row-parallel launch met a long row: the matrix changed under this pointer
  }
  *verdict = n.verdict;
  *flag_dev = n.flag_dev;
  *want_scan = n.verdict < 0 && !n.pending;
  return idx;
}

void notes_forget() {
  std::lock_guard<std::mutex> lk(g_note_mu);
  for (auto &n : g_notes) { n.rowptr = nullptr; n.m = -1; n.verdict = -1; n.pending = false; }
}

void note_scan_enqueued(int idx, const int *rowptr, cudaStream_t s) {
  std::lock_guard<std::mutex> lk(g_note_mu);
  GraphNote &n = g_notes[idx];
  if (n.rowptr != rowptr) return;   // recycled meanwhile
  if (cudaEventRecord(n.ev, s) == cudaSuccess) n.pending = true;
}
}  // namespace

// Graph notes allocate mapped memory, record and query events: none of that belongs into a stream capture, and a captured
// launch must not depend on what a note says at capture time.  While capturing, the segment path is taken unconditionally.
static bool stream_is_capturing(cudaStream_t s) {
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(s, &st) != cudaSuccess) { cudaGetLastError(); return true; }   // e.g. legacy stream during a global capture
  return st != cudaStreamCaptureStatusNone;
}

int spmm_last_path() { return g_last_path; }
void spmm_forget_graph_notes() { notes_forget(); }

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int pow2_at_least(int x, int lo, int hi) {
  int g = lo;
  while (g < x && g < hi) g <<= 1;
  return g;
}

int device_l2_bytes() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 126 << 20;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrL2CacheSize, dev) != cudaSuccess || n <= 0) n = 126 << 20;
    cached[dev] = n;
  }
  return cached[dev];
}

// Width (columns) of the B panel one pass over the nnz stream gathers from: 64 (16 lanes x float4), wider matrices run
// as 64-column panels one after the other (blockIdx.y), so that the gathered panel, K rows x 256 B, is what has to stay
// L2-resident, not all of B (measured on B200 against 128-column panels: reddit-like N=128 3.97 -> 3.29 ms, N=256
// 8.66 -> 6.84 ms).
// DEAD END, measured in round 2 (profiles/r02_exp_panels_*.jsonl, r02_ncu_panel_products128_*.csv, kernel kept under
// tools/dead_ends/): narrower panels for operands that do not fit the L2 (products-like N=128: B = 1.25 GB).  32 / 16 / 8
// columns: 9.2 / 20.4 / 43.1 ms against 8.65 ms at 64 — DRAM reads 57 -> 121 -> 249 GB, i.e. they DOUBLE per halving:
// (1) the L2 is tagged per 128 B line, so a 32 B slice per 512 B row occupies a whole line and the "78 MB" panel needs 313 MB of
// lines; a miss costs ~126 B of DRAM whatever the request size; (2) even fully resident (reddit-like N=128, 15 MB of
// slices) the gather is bound by REQUESTS, not bytes: 141 / 171 / 183 G requests/s at 128 / 64 / 32 B per request, i.e.
// 18 / 10.9 / 5.9 TB/s — a repacked contiguous 8-column panel would top out at 1.98 G requests / 183 G/s = 10.8 ms.
// Only >= 128 B per request runs at full L2 rate.  Option spmm_panel = 32 keeps the 8-lane geometry reachable for experiments.
static int pick_panel(int N, int64_t K, bool narrow_ok) {
  (void)N; (void)K; (void)narrow_ok;
  return option(OPT_SPMM_PANEL) == 32 ? 32 : 64;
}

// Lane-group geometry for feature width N and panel width W.  vec4 needs 16-byte aligned rows everywhere.
static void pick_geometry(int N, int W, bool can_vec4, int *vec, int *G, bool *narrow) {
  *narrow = false;
  if (can_vec4) {
    *vec = 4;
    // N=16 -> 4 lanes, 32 -> 8, >= 64 -> 16 lanes x float4; a narrower panel only when the matrix is wider than it
    *G = pow2_at_least((N + 3) / 4, 4, 16);
    if (W < 64 && N > W) *G = W / 4;
  } else {
    *vec = 1;
    *G = pow2_at_least(N, 4, 32);
  }
}

// Segment length (nnz per lane group).  Depends only on (N, nnz, with_arg, #SMs) so that
// spmm_workspace_bytes() and the launch agree.  Aim: segs_per_group() segments per resident group (dynamic balance
// through the block scheduler) while the partial workspace stays below kWorkspaceCap.
static constexpr size_t kWorkspaceCap = 192u << 20;
// Segments per resident lane group: more segments = finer dynamic balance through the block scheduler, fewer = fewer rows
// cut by a segment boundary (those are finished by the fix-up kernel, which on N GPUs is an NVLink-ingress burst).
// Measured on B200 (option spmm_segs): the main kernel does not care (reddit@64 1.573 / 1.568 / 1.567 / 1.570 / 1.575 ms
// for 8 / 4 / 3 / 2 / 1, products@128 8.95 / 8.93 / 8.95 ms for 8 / 4 / 3) while the fix-up shrinks with the segment count.
// Round 2 sweep on reddit@64 (segments per group / longest segment: step, kernel, fix-up ms): 3 / 8192: 1.580, 1.560, 0.0128;
// 2 / 8192: 1.580, 1.563, 0.0108; 1 / 16384: 1.584, 1.569, 0.0087; 1 / 32768: 1.581, 1.566, 0.0087 — level within noise.
// ONE segment per resident group, up to 16384 nonzeros, everywhere — the SAME layout whether the epilogue stores locally or
// fans out to other GPUs, so that the column-sharded result stays bit-identical to the single-GPU one (rows are folded at
// the same places) while as few rows as possible go through the fix-up grid's burst over NVLink.
// That holds for the single-panel 16-lane geometry (N = 64: the headline and every rank of the column shard).  Elsewhere
// four segments per group are better or level (`tools/bench_vs_ref.py`, ms, 1 vs 4): reddit-like N = 32 0.93 vs 0.86, N = 512
// 14.28 vs 13.43, products-like N = 32 2.77 vs 2.35 — with several column panels per launch the gathered panel changes at every
// transition and one long wave per panel makes that a long stretch in which two 60 MB panels compete for the L2; the 8-lane
// geometry (four row walkers per warp) has a longer tail in a single wave.
static int segs_per_group(int G, int panels) {
  const int v = option(OPT_SPMM_SEGS);
  if (v >= 1 && v <= 64) return v;
  return (G == 16 && panels == 1) ? 1 : 4;
}

// blocks_per_sm: CTAs of the kernel that will run that fit on one SM (SpmmKernel::blocks_per_sm; 3 when sizing the workspace,
// the most any flavour reaches = the most segments)
static int pick_chunk(int N, int64_t nnz, bool with_arg, int G, int blocks_per_sm) {
  const int64_t resident_groups = (int64_t)device_sm_count() * blocks_per_sm * (kSpmmThreads / G);
  const int spg = segs_per_group(G, (N + 4 * G - 1) / (4 * G));   // 4 * G = columns per panel in the 16-byte geometries
  int64_t chunk = (nnz + resident_groups * spg - 1) / (resident_groups * spg);
  // Small matrices (latency regime): one 32-nnz batch per segment spreads them over more SMs; with two or more column
  // panels (N > 64) the extra cut rows cost more than that buys.  p2p-Gnutella31 / ca-CondMat, us per call, min 32 | 64 | 128:
  // N=32 26.7 | 34.9 | 55.3 and 20.5 | 22.6 | 34.5;  N=64 24.7 | 29.5 | 41.0 and 20.5 | 22.4 | 28.2;  N=128 35.0 | 32.9 | 43.0
  // and 30.7 | 26.4 | 32.9.
  // Only while the GPU is underfilled, though: once 64-nnz segments give every resident group one (arxiv-like, 1.17 M nnz),
  // halving them only doubles the per-segment overhead (N=32 0.065 -> 0.088 ms).
  // Between the two (arxiv-like, 1.17 M nnz): ONE wave of blocks over all column panels, every resident group one segment —
  // a fixed 64-nnz segment left 1.28 waves at N = 32 (58.6 us against 45) and 2.56 at N = 128 (135 us against 114).
  if (chunk < 2 * kBatch) {
    const int min_chunk = (N <= 64) ? kBatch : 2 * kBatch;
    const int panels = (N + 4 * G - 1) / (4 * G);
    const int64_t per_panel = resident_groups / panels > 0 ? resident_groups / panels : 1;
    chunk = (nnz + per_panel - 1) / per_panel;
    if (chunk < min_chunk) chunk = min_chunk;
  }
  int cap = option(OPT_SPMM_CHUNK_CAP);
  if (cap < 64) cap = 16384;
  if (chunk > cap) {
    // longer matrices: a WHOLE number of segments per resident group again — a capped segment length would leave a sliver
    // of a second wave (products@128: 7 551 segments on 7 104 resident groups = 1.06 waves cost 9.6 ms against 8.7 ms)
    const int64_t rounds = (nnz + resident_groups * cap - 1) / (resident_groups * cap);
    chunk = (nnz + resident_groups * rounds - 1) / (resident_groups * rounds);
  }
  const size_t per_chunk = (size_t)2 * N * 4 * (with_arg ? 2 : 1) + 4;
  const int64_t min_for_ws = (int64_t)(((size_t)nnz * per_chunk + kWorkspaceCap - 1) / kWorkspaceCap);
  if (chunk < min_for_ws) chunk = min_for_ws;
  return (int)((chunk + kBatch - 1) / kBatch * kBatch);
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static void geometry_for(int N, int64_t nnz, bool with_arg, bool can_vec4, int W, int blocks_per_sm, int *vec, int *G, bool *narrow,
                         int *chunk, int *num_chunks) {
  pick_geometry(N, W, can_vec4, vec, G, narrow);
  *chunk = pick_chunk(N, nnz, with_arg, *G, blocks_per_sm);
  *num_chunks = (int)((nnz + *chunk - 1) / *chunk);
}

size_t spmm_workspace_bytes(int N, int64_t nnz, bool with_arg) {
  if (nnz <= 0 || N <= 0) return 256;
  // the segment length differs between the geometries (vec4 at every panel width, scalar); size for the largest need
  size_t need = 0;
  for (int pass = 0; pass < 3; pass++) {
    static const int widths[3] = {64, 32, 64};
    int vec, G, chunk, nc;
    bool narrow;
    geometry_for(N, nnz, with_arg, pass < 2, widths[pass], 3, &vec, &G, &narrow, &chunk, &nc);
    size_t b = align_up((size_t)nc * 4, 256) + align_up((size_t)nc * 2 * N * 4, 256) * (with_arg ? 2 : 1);
    if (b > need) need = b;
  }
  return need + 256;
}

cudaError_t spmm_csr(const SpmmProblem &p, void *workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (p.M <= 0 || p.N <= 0) return cudaSuccess;
  if (p.n_dst < 1 || p.n_dst > kMaxDst) return cudaErrorInvalidValue;
  const bool with_arg = p.E != nullptr;
  if (with_arg && p.reduce != R_MAX && p.reduce != R_MIN) return cudaErrorInvalidValue;
  const int comp = (p.compute == C_MASK) ? C_MASK : (p.val == nullptr) ? C_COPY : p.compute;
  if (comp == C_MASK && (p.mask == nullptr || p.reduce != R_SUM || with_arg)) return cudaErrorInvalidValue;

  bool can_vec4 = (p.N % 4 == 0) && (p.ldb % 4 == 0) && (p.ldc % 4 == 0) && aligned16(p.B) && aligned16(workspace);
  for (int d = 0; d < p.n_dst; d++) can_vec4 = can_vec4 && aligned16(p.dst[d]);
  if (with_arg) can_vec4 = can_vec4 && aligned16(p.E) && (p.lde % 4 == 0);
  if (comp == C_MASK) can_vec4 = can_vec4 && aligned16(p.mask) && (p.ldm % 4 == 0);

  SpmmArgs a;
  a.M = p.M; a.N = p.N; a.nnz = (int)p.nnz;
  a.rowptr = p.rowptr; a.col = p.col; a.val = p.val;
  a.B = p.B; a.ldb = p.ldb; a.ldc = p.ldc;
  a.E = p.E; a.lde = p.lde;
  a.mask = p.mask; a.ldm = p.ldm;
  a.mean = (p.reduce == R_MEAN);
  a.nnz_dev = p.nnz_on_device ? 1 : 0;
  a.nnz_report = p.nnz_report;
  a.n_dst = p.n_dst;
  a.mcast = p.mcast;
  if (p.mcast && p.n_dst != 1) return cudaErrorInvalidValue;
  for (int d = 0; d < kMaxDst; d++) a.dst[d] = d < p.n_dst ? p.dst[d] : nullptr;

  int vec = 1, G = 32;
  bool narrow = false;
  a.chunk = kBatch; a.num_chunks = 0;
  a.part_val = nullptr; a.part_arg = nullptr; a.tail_row = nullptr;
  a.hub_flag = nullptr; a.hub_limit = kRowParLimit;

  // latency regime: single-launch row-parallel kernel when this matrix is known to have short rows only
  int note = -1;
  bool want_scan = false;
  const int rp_mode = rowpar_mode();
  if (p.nnz > 0 && p.nnz <= kRowParMaxNnz && can_vec4 && (comp == C_MUL || comp == C_COPY) && rp_mode != 0 &&
      !stream_is_capturing(stream)) {
    int verdict = -1, *flag_dev = nullptr;
    note = note_lookup(p.rowptr, p.M, &verdict, &flag_dev, &want_scan);
    if (note >= 0) a.hub_flag = flag_dev;
    if ((note >= 0 && verdict == 0) || rp_mode == 1) {
      pick_geometry(p.N, 64, true, &vec, &G, &narrow);
      SpmmLaunchFn fn = G == 4 ? spmm_rowpar_lookup_v4_g4(p.reduce, comp, with_arg)
                      : G == 8 ? spmm_rowpar_lookup_v4_g8(p.reduce, comp, with_arg)
                               : spmm_rowpar_lookup_v4_g16(p.reduce, comp, with_arg);
      if (fn != nullptr) {
        const int gpb = kSpmmThreads / G;
        dim3 grid((p.M + gpb - 1) / gpb, (p.N + G * vec - 1) / (G * vec));
        ProfileScope prof(1, stream);
        g_last_path = 1;
        return fn(a, grid, stream);
      }
    }
    if (!want_scan) a.hub_flag = nullptr;   // the segment path only reports while the note is still open
  }
  g_last_path = 0;
  if (p.nnz > 0) {
    const int W = pick_panel(p.N, p.K > 0 ? p.K : p.M, false);
    pick_geometry(p.N, W, can_vec4, &vec, &G, &narrow);
    SpmmKernel kern;
    const int key = vec * 100 + G;
    switch (key) {
    case 404: kern = spmm_lookup_v4_g4(p.reduce, comp, with_arg); break;
    case 408: kern = spmm_lookup_v4_g8(p.reduce, comp, with_arg); break;
    case 416: kern = spmm_lookup_v4_g16(p.reduce, comp, with_arg); break;
    case 432: kern = spmm_lookup_v4_g32(p.reduce, comp, with_arg); break;
    case 104: kern = spmm_lookup_v1_g4(p.reduce, comp, with_arg); break;
    case 108: kern = spmm_lookup_v1_g8(p.reduce, comp, with_arg); break;
    case 116: kern = spmm_lookup_v1_g16(p.reduce, comp, with_arg); break;
    case 132: kern = spmm_lookup_v1_g32(p.reduce, comp, with_arg); break;
    }
    if (!kern) return cudaErrorInvalidValue;
    int bps = kern.blocks_per_sm();
    if (bps > 3) bps = 3;      // the workspace is sized for at most 3 CTAs per SM worth of segments
    geometry_for(p.N, p.nnz, with_arg, can_vec4, W, bps, &vec, &G, &narrow, &a.chunk, &a.num_chunks);
    const size_t tail_b = align_up((size_t)a.num_chunks * 4, 256);
    const size_t part_b = align_up((size_t)a.num_chunks * 2 * p.N * 4, 256);
    const size_t need = tail_b + part_b * (with_arg ? 2 : 1);
    if (workspace == nullptr || workspace_bytes < need) return cudaErrorInvalidValue;
    char *w = static_cast<char *>(workspace);
    a.tail_row = reinterpret_cast<int *>(w);
    a.part_val = reinterpret_cast<float *>(w + tail_b);
    a.part_arg = with_arg ? reinterpret_cast<int *>(w + tail_b + part_b) : nullptr;
    SpmmLaunchFn fn = kern.launch;
    if (fn == nullptr) return cudaErrorInvalidValue;
    const int gpb = kSpmmThreads / G;
    dim3 grid((a.num_chunks + gpb - 1) / gpb, (p.N + G * vec - 1) / (G * vec));
    cudaError_t e;
    {
      ProfileScope prof(1, stream);
      e = fn(a, grid, stream);
    }
    if (e != cudaSuccess) return e;
  }
  const int fv = can_vec4 ? 4 : 1;
  const int64_t fold_threads = (int64_t)a.num_chunks * (p.N / fv);
  const int64_t empty_threads = ((int64_t)p.M + 31) / 32 * 32;
  const int64_t threads = fold_threads > empty_threads ? fold_threads : empty_threads;
  const int blocks = (int)((threads + 255) / 256);
  // dependent launch only right behind the SpMM grid, and not while the bench brackets the two launches with events
  const bool pdl = p.nnz > 0 && !profile_is_on() && option(OPT_SPMM_NO_PDL) != 1;
  cudaError_t fe;
  {
    ProfileScope prof(2, stream);
    if (can_vec4)
      fe = with_arg ? launch_fixup<true, 4>(p.reduce, a, blocks, pdl, stream) : launch_fixup<false, 4>(p.reduce, a, blocks, pdl, stream);
    else
      fe = with_arg ? launch_fixup<true, 1>(p.reduce, a, blocks, pdl, stream) : launch_fixup<false, 1>(p.reduce, a, blocks, pdl, stream);
  }
  if (fe == cudaSuccess && note >= 0 && want_scan && a.hub_flag != nullptr) note_scan_enqueued(note, p.rowptr, stream);   // the fix-up grid scanned every row
  return fe;
}

}  // namespace dgs
