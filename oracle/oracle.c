/*
 * oracle.c — CPU restatement of the dgSPARSE hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may load this library.  The product path (dgsparse-lib_b200/) never links or calls it.
 *
 * Every function restates one piece of the reference (paths relative to /root/reference):
 *
 *   oracle_spmm            example/util/sp_util.hpp:62-84   (serial fp32 accumulation in nnz order)
 *                          include/cuda/spmm_cuda.cuh:27-54 (REDUCE ops, arg index E, empty row -> 0 / -1,
 *                                                            MEAN divides by the row's nnz count)
 *                          include/gspmm.h:19-146           (MAX/MIN identities INT_MIN/INT_MAX as float,
 *                                                            strict compare -> first extremum wins)
 *                          src/gspmm-fp/gspmm.h:53-79, src/gspmm-fp/gspmm.cu:210-249
 *                                                           (COMPUTE ops: a = edge value, b = feature;
 *                                                            Sub = b - a, Div = b / a; empty row -> 0)
 *   oracle_spmm_f64        same loops with a double accumulator (tolerance anchor for long rows)
 *   oracle_sddmm_csr       example/util/sp_util.hpp:87-112  + MEAN scaling of
 *                          include/cuda/sddmm_cuda.cuh:222-311 (divide each edge by its row degree)
 *   oracle_sddmm_coo       src/sddmm/coosddmm_ebalance.cuh:5-212 semantics (dot of D1[row[e]], D2[col[e]])
 *   oracle_spmm_mask       include/cuda/spmm_cuda.cuh:400-433 — the INTENDED semantics (only entries with
 *                          E[col, v] == row contribute; the reference leaves val_pre_red uninitialised
 *                          when the mask is false)
 *   oracle_sddmm_csr_mask  include/cuda/sddmm_cuda.cuh:403-507 (only feature positions with
 *                          E[row, c] == col contribute)
 *   oracle_csr2csc         include/cuda/csr2csc.cuh:8-26 (cusparseCsr2cscEx2 ALG1) pinned by
 *                          test/test_csr2csr.py:42-49 against scipy tocsc(): stable counting sort,
 *                          rows ascending within each column
 *   oracle_spconv          test/test_spconv.py:17-53 (cpu_compute): out[omap[p]] += in[imap[p]] . W[k]
 *
 * Floating point: compiled with -ffp-contract=off so `acc += a * b` is a separate multiply and add,
 * exactly as the reference's host functions behave on baseline x86-64.
 */
#include <limits.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

enum { R_SUM = 0, R_MAX = 1, R_MIN = 2, R_MEAN = 3 };      /* include/gspmm.h:13 */
enum { C_ADD = 0, C_SUB = 1, C_MUL = 2, C_DIV = 3, C_COPY = 4 }; /* include/gspmm.h:14 (+COPY = no edge value) */

static inline float compute_op(int cop, float a /*edge*/, float b /*feat*/) {
  switch (cop) {
  case C_ADD: return a + b;
  case C_SUB: return b - a;   /* src/gspmm-fp/gspmm.h:67-72 */
  case C_MUL: return a * b;
  case C_DIV: return b / a;   /* src/gspmm-fp/gspmm.h:74-79 */
  default:    return b;       /* topo kernels: copy_u */
  }
}

static inline float reduce_init(int rop) {
  if (rop == R_MAX) return (float)INT_MIN;  /* include/gspmm.h:137-138 */
  if (rop == R_MIN) return (float)INT_MAX;  /* include/gspmm.h:139-140 */
  return 0.0f;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* out[M,N] row-major (leading dim N), B row-major with leading dim ldb (>= N).
 * val == NULL  => edge value 1 (include/cuda/cuda_util.cuh:276-290 __guard_load_default_one);
 * E   == NULL  => arg index not produced. E holds the COLUMN index of the winning nnz, -1 if none. */
void oracle_spmm(int M, int N, const int *rowptr, const int *col, const float *val,
                 const float *B, int64_t ldb, int rop, int cop, float *out, int *E) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t r = 0; r < M; r++) {
    int start = rowptr[r], end = rowptr[r + 1];
    for (int v = 0; v < N; v++) {
      float res = reduce_init(rop);
      int arg = -1;
      if (end - start > 0) {
        for (int p = start; p < end; p++) {
          int c = col[p];
          float a = val ? val[p] : 1.0f;
          float x = compute_op(val ? cop : C_COPY, a, B[(int64_t)c * ldb + v]); /* 1.0f * b == b */
          if (rop == R_MAX) {
            if (res < x) { arg = c; }
            res = (res < x) ? x : res;   /* MAX(a,b) ((a<b)?b:a) include/gspmm.h:17 */
          } else if (rop == R_MIN) {
            if (res > x) { arg = c; }
            res = (res < x) ? res : x;   /* MIN(a,b) ((a<b)?a:b) include/gspmm.h:16 */
          } else {
            res = res + x;
          }
        }
        if (rop == R_MEAN) res /= (float)(end - start);
      } else {
        res = 0.0f;
      }
      out[r * N + v] = res;
      if (E) E[r * N + v] = arg;
    }
  }
}

/* Same semantics, double accumulator for SUM/MEAN (tolerance anchor). MAX/MIN are exact anyway. */
void oracle_spmm_f64(int M, int N, const int *rowptr, const int *col, const float *val,
                     const float *B, int64_t ldb, int rop, int cop, double *out) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t r = 0; r < M; r++) {
    int start = rowptr[r], end = rowptr[r + 1];
    for (int v = 0; v < N; v++) {
      double res = (double)reduce_init(rop);
      if (end - start > 0) {
        for (int p = start; p < end; p++) {
          int c = col[p];
          double a = val ? (double)val[p] : 1.0, b = (double)B[(int64_t)c * ldb + v], x;
          int op = val ? cop : C_COPY;
          switch (op) {
          case C_ADD: x = a + b; break;
          case C_SUB: x = b - a; break;
          case C_MUL: x = a * b; break;
          case C_DIV: x = b / a; break;
          default: x = b;
          }
          if (rop == R_MAX) res = (res < x) ? x : res;
          else if (rop == R_MIN) res = (res < x) ? res : x;
          else res += x;
        }
        if (rop == R_MEAN) res /= (double)(end - start);
      } else {
        res = 0.0;
      }
      out[r * N + v] = res;
    }
  }
}

/* out[e] = sum_k D1[row(e),k] * D2[col(e),k]; mean != 0 divides by the row degree. */
void oracle_sddmm_csr(int M, int K, const int *rowptr, const int *col, const float *D1,
                      const float *D2, int mean, float *out) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t i = 0; i < M; i++) {
    int lb = rowptr[i], hb = rowptr[i + 1];
    for (int p = lb; p < hb; p++) {
      float acc = 0;
      const float *a = D1 + i * K, *b = D2 + (int64_t)col[p] * K;
      for (int k = 0; k < K; k++) acc += a[k] * b[k];
      if (mean && hb - lb > 0) acc /= (float)(hb - lb);
      out[p] = acc;
    }
  }
}

void oracle_sddmm_csr_f64(int M, int K, const int *rowptr, const int *col, const float *D1,
                          const float *D2, int mean, double *out) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t i = 0; i < M; i++) {
    int lb = rowptr[i], hb = rowptr[i + 1];
    for (int p = lb; p < hb; p++) {
      double acc = 0;
      const float *a = D1 + i * K, *b = D2 + (int64_t)col[p] * K;
      for (int k = 0; k < K; k++) acc += (double)a[k] * (double)b[k];
      if (mean && hb - lb > 0) acc /= (double)(hb - lb);
      out[p] = acc;
    }
  }
}

void oracle_sddmm_coo(int K, int64_t nnz, const int *row, const int *col, const float *D1,
                      const float *D2, float *out) {
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < nnz; e++) {
    float acc = 0;
    const float *a = D1 + (int64_t)row[e] * K, *b = D2 + (int64_t)col[e] * K;
    for (int k = 0; k < K; k++) acc += a[k] * b[k];
    out[e] = acc;
  }
}

/* Max/min backward wrt dense, called on the CSC arrays (ptr = colptr, idx = row):
 * out[j,v] = sum_{p in ptr[j]..ptr[j+1]} [E[idx[p],v] == j] * val[p] * G[idx[p],v]. */
void oracle_spmm_mask(int M, int N, const int *ptr, const int *idx, const float *val,
                      const float *G, const int *E, float *out) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t j = 0; j < M; j++) {
    for (int v = 0; v < N; v++) {
      float res = 0;
      for (int p = ptr[j]; p < ptr[j + 1]; p++) {
        int i = idx[p];
        if (E[(int64_t)i * N + v] == (int)j) res += (val ? val[p] : 1.0f) * G[(int64_t)i * N + v];
      }
      out[j * N + v] = res;
    }
  }
}

/* Max/min backward wrt values: out[e] = sum_c [E[row(e),c] == col(e)] * D1[row(e),c] * D2[col(e),c]. */
void oracle_sddmm_csr_mask(int M, int K, const int *rowptr, const int *col, const float *D1,
                           const float *D2, const int *E, float *out) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t i = 0; i < M; i++) {
    for (int p = rowptr[i]; p < rowptr[i + 1]; p++) {
      float acc = 0;
      int c = col[p];
      for (int k = 0; k < K; k++)
        if (E[i * K + k] == c) acc += D1[i * K + k] * D2[(int64_t)c * K + k];
      out[p] = acc;
    }
  }
}

/* Stable CSR -> CSC. perm[q] = CSR position of the q-th CSC entry (exact int32).
 * val / val_t may be NULL. */
void oracle_csr2csc(int M, int ncols, const int *rowptr, const int *col, const float *val,
                    int *colptr, int *row, float *val_t, int *perm) {
  int nnz = rowptr[M];
  memset(colptr, 0, sizeof(int) * ((size_t)ncols + 1));
  for (int p = 0; p < nnz; p++) colptr[col[p] + 1]++;
  for (int c = 0; c < ncols; c++) colptr[c + 1] += colptr[c];
  int *cursor = (int *)malloc(sizeof(int) * ((size_t)ncols + 1));
  memcpy(cursor, colptr, sizeof(int) * ((size_t)ncols + 1));
  for (int r = 0; r < M; r++) {
    for (int p = rowptr[r]; p < rowptr[r + 1]; p++) {
      int q = cursor[col[p]]++;
      row[q] = r;
      if (val_t && val) val_t[q] = val[p];
      if (perm) perm[q] = p;
    }
  }
  free(cursor);
}

/* Sparse 3-D convolution gather-GEMM-scatter, test/test_spconv.py:17-53 (cpu_compute):
 * for each kernel offset k and each pair p in [kpos[k], kpos[k+1]):
 *   out[omap[p], :] += in[imap[p], :] @ W[k]        W: [k_vol, c_in, c_out] row-major;
 * then, with `precompute` (test_spconv.py:46-52, the separate_mid form of a submanifold layer):
 *   out[i, :] += in[i, :] @ W[mid_k] for every output row i, mid_k = k_vol / 2 for odd k_vol, else 0.
 * Per output element the additions happen in the same order as cpu_compute (k, pair, c_in ascending; product
 * rounded to fp32, then added), so the result is bit-identical to it (pinned by tests/test_oracle_golden.py). */
void oracle_spconv(int k_vol, int c_in, int c_out, int out_nnz, const int *kpos, const int *imap,
                   const int *omap, const float *in, const float *W, float *out, int precompute) {
  memset(out, 0, sizeof(float) * (size_t)out_nnz * c_out);
  for (int k = 0; k < k_vol; k++) {
    const float *Wk = W + (size_t)k * c_in * c_out;
    for (int p = kpos[k]; p < kpos[k + 1]; p++) {
      const float *x = in + (size_t)imap[p] * c_in;
      float *y = out + (size_t)omap[p] * c_out;
      for (int ci = 0; ci < c_in; ci++) {
        float xv = x[ci];
        for (int co = 0; co < c_out; co++) y[co] += xv * Wk[(size_t)ci * c_out + co];
      }
    }
  }
  if (precompute) {
    const int mid_k = (k_vol % 2 == 1) ? k_vol / 2 : 0;
    const float *Wk = W + (size_t)mid_k * c_in * c_out;
    for (int i = 0; i < out_nnz; i++) {
      const float *x = in + (size_t)i * c_in;
      float *y = out + (size_t)i * c_out;
      for (int ci = 0; ci < c_in; ci++) {
        float xv = x[ci];
        for (int co = 0; co < c_out; co++) y[co] += xv * Wk[(size_t)ci * c_out + co];
      }
    }
  }
}
