/*
 * dgsparse.h — the legacy dgSPARSE C ABI, served by libdgsparse_b200.so.
 *
 * Drop-in for the reference's public header /root/reference/include/dgsparse.h:4-19 and for the
 * GE-SpMM / SDDMM library headers src/ge-spmm/gespmm.h:9-40, src/sddmm/sddmm.h:6-12 (same symbol
 * names, argument order and meaning; authoritative export list = nm -D lib/dgsparse.so).
 *
 * Conventions kept from the reference: every pointer is a DEVICE pointer owned by the caller;
 * dense matrices are row-major; `k` of spmm_cuda is the dense column count N (gespmm.cc:114);
 * calls are asynchronous on the legacy default stream (stream 0); `out` is fully overwritten.
 * Differences: any K works for sddmm_cuda_csr (q13); errors are reported on
 * stderr and through dgs_last_error() instead of exit() (q14).
 */
#ifndef DGSPARSE_H
#define DGSPARSE_H

#ifdef __cplusplus
extern "C" {
#endif

/* replaces spmm_cuda, src/ge-spmm/gespmm.cc:114-129:  out[m,k] = A_csr[m,*] . dense[*,k] */
void spmm_cuda(int m, int k, int *rowptr, int *colind, float *values, float *dense, float *out);

/* replaces spmm_cuda_no_edge_value, src/ge-spmm/gespmm.cc:131-134 (values ignored, treated as 1) */
void spmm_cuda_no_edge_value(int m, int k, int *rowptr, int *colind, float *values, float *dense, float *out);

/* replaces sddmm_cuda_coo, src/sddmm/sddmm.cu:8-23:  out[e] = dot(D1[rowind[e],:k], D2[colind[e],:k]) */
void sddmm_cuda_coo(int k, int nnz, int *rowind, int *colind, float *D1, float *D2, float *out);

/* replaces sddmm_cuda_csr, src/sddmm/sddmm.cu:25-41 */
void sddmm_cuda_csr(int m, int k, int nnz, int *rowptr, int *colind, float *D1, float *D2, float *out);

/* Declared by the reference header (include/dgsparse.h:17-18) but defined nowhere in it.  Provided
 * here with the only sensible reading: per row r and head h, softmax over the row's nonzeros of
 * values[p*head + h]. */
void edge_softmax_cuda(int mrows, int head, int *rowptr, float *values, float *softmax);

/* src/ge-spmm/gespmm.h:9-33 */
struct SpMatCsrDescr_t {
  int nrow;
  int ncol;
  int nnz;      /* may be -1 (spmm_cuda passes that): then read from indptr[nrow] */
  int *indptr;
  int *indices;
  float *data;  /* may be NULL: no edge value */
};

enum gespmmAlg_t {
  GESPMM_ALG_SEQREDUCE_ROWBALANCE = 0,
  GESPMM_ALG_PARREDUCE_ROWBALANCE,
  GESPMM_ALG_SEQREDUCE_NNZBALANCE,
  GESPMM_ALG_PARREDUCE_NNZBALANCE,
  GESPMM_ALG_SEQREDUCE_ROWBALANCE_NON_TRANSPOSE,
  GESPMM_ALG_PARREDUCE_ROWBALANCE_NON_TRANSPOSE,
  GESPMM_ALG_SEQREDUCE_NNZBALANCE_NON_TRANSPOSE,
  GESPMM_ALG_PARREDUCE_NNZBALANCE_NON_TRANSPOSE,
  GESPMM_ALG_ROWCACHING_ROWBALANCE,
  GESPMM_ALG_ROWCACHING_NNZBALANCE,
  GESPMM_ALG_DEFAULT
};

/* replaces gespmmCsrSpMM, src/ge-spmm/gespmm.cc:29-111.  `alg` is accepted and ignored: every
 * value runs the one row-segment kernel (all reference algorithms compute the same C).
 * transpose_BC = true: B[ncol,N], C[nrow,N] row-major.  false: column-major (ldB = ncol, ldC = nrow),
 * which needs spmatA.ncol. */
#ifdef __cplusplus
void gespmmCsrSpMM(const SpMatCsrDescr_t spmatA, float *B, const int N, float *C, bool transpose_BC, gespmmAlg_t alg);
#else
void gespmmCsrSpMM(const struct SpMatCsrDescr_t spmatA, float *B, const int N, float *C, _Bool transpose_BC,
                   enum gespmmAlg_t alg);
#endif

/* The older SpMV/SpMM API of the same library, src/ge-spmm/gespmm_v2.h:8-38 (exported by lib/dgsparse.so):
 * dnOutput[nr, nv] = A[nr, nc] . dnInput[nc, nv] for CSR (rowPtr) or COO (rowIdx, sorted by row) inputs.
 * Every algorithm value computes the same product and runs the one row-segment kernel; the output is
 * OVERWRITTEN (the reference's COO / merge algorithms accumulate with atomicAdd into a caller-zeroed buffer,
 * gespmm_csrcoo_v2.cu:94-214 — same result).  values may be NULL (treated as 1, __guard_load_default_one).
 * Any nv works (the reference handles nv in {1, 2, 4, 8, 16, 32} only). */
enum SPMV_SPMM_ALG { ALG_CSR_SCALAR, ALG_CSR_VECTOR, ALG_COO_SCALAR, ALG_COO_VECTOR };
enum SparseFormat { SPARSE_FORMAT_CSR, SPARSE_FORMAT_COO };
enum DenseLayout { DENSE_ROW_MAJOR, DENSE_COL_MAJOR };

/* replaces cuda_csr_coo_spmm, src/ge-spmm/gespmm_csrcoo_v2.cu:606-790.  rowPtr may be NULL for the COO algorithms
 * (it is then rebuilt from rowIdx); rowIdx is ignored when rowPtr is given. */
#ifdef __cplusplus
void cuda_csr_coo_spmm(SPMV_SPMM_ALG kAlg, DenseLayout layout, const int nr, const int nc, const int nnz, const int nv,
                       const int *rowPtr, const int *rowIdx, const int *colIdx, const float *values,
                       const float *dnInput, float *dnOutput);
#else
void cuda_csr_coo_spmm(enum SPMV_SPMM_ALG kAlg, enum DenseLayout layout, const int nr, const int nc, const int nnz,
                       const int nv, const int *rowPtr, const int *rowIdx, const int *colIdx, const float *values,
                       const float *dnInput, float *dnOutput);
#endif

/* replaces cuda_csr_spmm, src/ge-spmm/gespmm_v2.cu:569-760.  algo_code 0..3 (row-scalar, row-vector, merge-scalar,
 * merge-vector) all compute the same product; layout_code 0 = column-major, 1 = row-major dense operands. */
void cuda_csr_spmm(int algo_code, int layout_code, int nr, int nc, int nv, int nnz, int *csrRowPtr, int *csrCol,
                   float *csrVal, float *vin, float *vout);

#ifdef __cplusplus
}
#endif
#endif /* DGSPARSE_H */
