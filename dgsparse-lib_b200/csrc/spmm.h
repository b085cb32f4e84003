// spmm.h — internal C++ interface between the C ABI (cabi.cu) and the kernels.  Torch-free.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace dgs {

struct SpmmProblem {
  int M = 0, N = 0;
  int K = 0;                    // rows of B (columns of A); 0 = unknown, taken as M (square adjacency).  Only sizes the column panels.
  int64_t nnz = 0;
  const int *rowptr = nullptr;
  const int *col = nullptr;
  const float *val = nullptr;   // null -> no edge value (copy_u / has_value = false)
  const float *B = nullptr;
  int64_t ldb = 0;
  int n_dst = 1;
  int mcast = 0;                // dst[0] is an NVLS multicast address covering every rank's C (requires n_dst == 1)
  float *dst[8] = {nullptr};    // dst[0] = C; further entries = NVLink peers' C (column-shard epilogue)
  int64_t ldc = 0;
  int *E = nullptr;             // arg column index out (MAX / MIN only)
  int64_t lde = 0;
  int reduce = 0;               // dgs::ReduceOp
  int compute = 2;              // dgs::ComputeOp (C_MUL)
  const int *mask = nullptr;    // compute == C_MASK: the forward's arg tensor, row stride ldm
  int64_t ldm = 0;
  // Legacy entry points without an nnz argument: `nnz` above is only a HINT (what rowptr[M] held on an earlier call); the
  // kernels read the true value on the device and re-derive the segment layout, and report it into *nnz_report (mapped
  // host memory, may be null) for the next call.  No device->host copy, no synchronisation.
  bool nnz_on_device = false;
  int *nnz_report = nullptr;
};

size_t spmm_workspace_bytes(int N, int64_t nnz, bool with_arg);
cudaError_t spmm_csr(const SpmmProblem &p, void *workspace, size_t workspace_bytes, cudaStream_t stream);
int device_sm_count();
void spmm_forget_graph_notes();   // drop what the library remembers about matrices it has seen (tests, benchmarks)
int spmm_last_path();   // which kernel family the calling thread's last spmm_csr launched: 0 row-segment (+ fix-up), 1 row-parallel

// bench-only launch timing (see dgs_profile_enable in include/dgsparse_b200.h)
int profile_enable(bool on);
int profile_collect(int max_records, int *kernel_ids, float *ms);
struct ProfileScope {   // RAII: records an event pair around the launches issued inside the scope
  ProfileScope(int kernel_id, cudaStream_t s);
  ~ProfileScope();
  int slot;
  cudaStream_t stream;
};

struct SddmmProblem {
  int M = 0, K = 0;
  int64_t nnz = 0;
  const int *rowptr = nullptr;  // CSR rows (null for COO)
  const int *row = nullptr;     // COO row index per edge (null for CSR)
  const int *col = nullptr;
  const float *D1 = nullptr;    // [M, K] row-major, leading dim ld1
  const float *D2 = nullptr;    // [*, K] row-major, leading dim ld2
  int64_t ld1 = 0, ld2 = 0;
  const int *E = nullptr;       // arg mask [M, K] (max/min backward), null otherwise
  int mean = 0;                 // divide by the row degree (CSR only)
  float *out = nullptr;         // [nnz]
};
cudaError_t sddmm(const SddmmProblem &p, cudaStream_t stream);
void sddmm_last_geometry(int *wpc, int *ctas_per_sm, int *chunk);   // of the calling thread's last sddmm(); ring kernel: wpc > 0

size_t csr2csc_workspace_bytes(int M, int ncols, int64_t nnz);
// colptr[ncols+1], row[nnz], val_t[nnz] (optional), perm[nnz] (optional): stable transpose
cudaError_t csr2csc(int M, int ncols, int64_t nnz, const int *rowptr, const int *col, const float *val,
                    int *colptr, int *row, float *val_t, int *perm, void *workspace, size_t workspace_bytes,
                    cudaStream_t stream);

}  // namespace dgs
