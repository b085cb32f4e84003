// torch_ops.cpp — the reference's PyTorch operator boundary as a COMPILED, loadable op library:
// dgsparse/_spmm_cuda.so, found by PathFinder and loaded with torch.ops.load_library exactly as
// /root/reference/dgsparse/__init__.py:16-26 does, usable from Python, TorchScript and libtorch C++ alike.
//
// Registers TORCH_LIBRARY(dgsparse_spmm): spmm_sum / spmm_max / spmm_min / spmm_mean / csr2csc with the schemas the
// reference's C++ signatures imply (src/spmm.cpp:15-33, 83-94, 264-270) and the same autograd contract (gradients for
// `values` (argument 2) and `dense` (argument 6) only, src/spmm.cpp:76-78), plus csr2csc_perm / sddmm_csr / sddmm_coo
// which the Python package uses.  No kernels here: every op calls the torch-free C ABI of libdgsparse_b200.so
// (include/dgsparse_b200.h) on the CURRENT stream, with outputs and scratch from the caching allocator.
//
// Deviations from the reference, all fixes: `algorithm` accepted and ignored (q2), the arg index E only
// materialised for max / min (q6), launches on the current stream under a device guard with error checks (q14), exact
// int32 CSR->CSC permutation (q10), mean backward wrt dense scaled by the degree of the SOURCE row.
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGraphsC10Utils.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/autograd.h>
#include <torch/library.h>

#include <vector>

#include "../../include/dgsparse_b200.h"

namespace {

using at::Tensor;
using torch::autograd::AutogradContext;
using torch::autograd::tensor_list;

void check(int rc, const char *what) { TORCH_CHECK(rc == 0, what, " failed: ", dgs_last_error(), " (cudaError ", rc, ")"); }

void require_cuda(const Tensor &t, const char *name) {
  TORCH_CHECK(t.is_cuda(), "dgsparse (B200) ops need CUDA tensors; there is no CPU path (", name, " is on ", t.device(), ")");
}
// (every microsecond counts on the reference's own benchmark shape — a Cora-sized graph, ~12 us per call — so the common
//  case of an already contiguous tensor does not go through the dispatcher)
Tensor i32c(const Tensor &t, const char *name) {
  require_cuda(t, name);
  TORCH_CHECK(t.scalar_type() == at::kInt, name, " must be int32 (got ", t.scalar_type(), ")");   // dgsparse/storage.py:27-82
  return t.is_contiguous() ? t : t.contiguous();
}
Tensor f32c(const Tensor &t, const char *name) {
  require_cuda(t, name);
  TORCH_CHECK(t.scalar_type() == at::kFloat, name, " must be float32 (got ", t.scalar_type(), ")");
  return t.is_contiguous() ? t : t.contiguous();
}
void *stream_of(const Tensor &t) { return at::cuda::getCurrentCUDAStream(t.get_device()).stream(); }
// Kernel scratch.  Small requests reuse one buffer per (thread, device, stream) — work on one stream is ordered, so the next
// call may overwrite it — instead of paying a caching-allocator round trip per call; large ones, and anything inside a
// CUDA-graph capture (the graph must own its memory), are allocated per call.
Tensor scratch(size_t bytes, const Tensor &like) {
  const int64_t n = (int64_t)(bytes < 256 ? 256 : bytes);
  constexpr int64_t kCached = 8 << 20;
  if (n > kCached || c10::cuda::currentStreamCaptureStatusMayInitCtx() != c10::cuda::CaptureStatus::None)
    return at::empty({n}, like.options().dtype(at::kByte));
  struct Slot { int device = -1; void *stream = nullptr; Tensor buf; };
  static thread_local Slot slots[4];
  static thread_local int next = 0;
  const int device = like.get_device();
  void *stream = stream_of(like);
  for (Slot &sl : slots)
    if (sl.device == device && sl.stream == stream && sl.buf.defined() && sl.buf.numel() >= n) return sl.buf;
  Slot &sl = slots[next++ % 4];
  sl.device = device; sl.stream = stream;
  sl.buf = at::empty({n < (1 << 20) ? (int64_t)(1 << 20) : n}, like.options().dtype(at::kByte));
  return sl.buf;
}
const int *iptr(const Tensor &t) { return t.defined() ? t.data_ptr<int>() : nullptr; }
const float *fptr(const Tensor &t) { return t.defined() ? t.data_ptr<float>() : nullptr; }

// out[,E] = generalized CSR SpMM (values undefined -> no edge value)
Tensor spmm_fwd(const Tensor &rowptr_, const Tensor &col_, const Tensor &values_, const Tensor &dense_, int reduce,
                Tensor *E_out) {
  const Tensor rowptr = i32c(rowptr_, "rowptr"), col = i32c(col_, "col"), dense = f32c(dense_, "dense");
  TORCH_CHECK(dense.dim() == 2, "dense must be 2-D [K, N]");
  Tensor values;
  if (values_.defined()) {
    values = f32c(values_, "values");
    if (values.dim() != 1) values = values.reshape({-1});
    TORCH_CHECK(values.numel() == col.numel(), "values must have one entry per nonzero");
  }
  const int64_t M = rowptr.numel() - 1, N = dense.size(1), nnz = col.numel();
  const c10::cuda::CUDAGuard guard(dense.device());
  Tensor out = at::empty({M, N}, dense.options());
  Tensor E;
  if (E_out) E = at::empty({M, N}, dense.options().dtype(at::kInt));
  if (M > 0 && N > 0) {
    const size_t wsb = dgs_spmm_workspace_bytes((int)N, nnz, E_out != nullptr);
    Tensor ws = scratch(wsb, dense);
    check(dgs_spmm_csr_k((int)M, (int)dense.size(0), (int)N, nnz, iptr(rowptr), iptr(col), fptr(values), fptr(dense),
                         dense.stride(0), out.data_ptr<float>(), N, E_out ? E.data_ptr<int>() : nullptr, E_out ? N : 0, reduce,
                         DGS_MUL, ws.data_ptr(), (size_t)ws.numel(), stream_of(dense)), "dgs_spmm_csr_k");
  }
  if (E_out) *E_out = E;
  return out;
}

// max / min backward wrt dense on the CSC arrays (spmm_cuda_with_mask, src/cuda/spmm_cuda.cu:255-303)
Tensor spmm_masked(const Tensor &ptr_, const Tensor &idx_, const Tensor &values_, const Tensor &grad_, const Tensor &E_) {
  const Tensor ptr = i32c(ptr_, "ptr"), idx = i32c(idx_, "idx"), grad = f32c(grad_, "grad"), E = i32c(E_, "E");
  Tensor values;
  if (values_.defined()) values = f32c(values_, "values").reshape({-1});
  const int64_t M = ptr.numel() - 1, N = grad.size(1), nnz = idx.numel();
  const c10::cuda::CUDAGuard guard(grad.device());
  Tensor out = at::empty({M, N}, grad.options());
  if (M > 0 && N > 0) {
    Tensor ws = scratch(dgs_spmm_workspace_bytes((int)N, nnz, 0), grad);
    check(dgs_spmm_csr_mask((int)M, (int)N, nnz, iptr(ptr), iptr(idx), fptr(values), fptr(grad), grad.stride(0), iptr(E),
                            E.stride(0), out.data_ptr<float>(), N, ws.data_ptr(), (size_t)ws.numel(), stream_of(grad)),
          "dgs_spmm_csr_mask");
  }
  return out;
}

// [1, nnz] = dot(D1[row(e)], D2[col(e)]) (+ MEAN scaling, + arg mask): sddmm_cuda_csr(Tensor...), src/cuda/spmm_cuda.cu:330-361
Tensor sddmm_csr_impl(const Tensor &rowptr_, const Tensor &col_, const Tensor &D1_, const Tensor &D2_, const Tensor &E_, bool mean) {
  const Tensor rowptr = i32c(rowptr_, "rowptr"), col = i32c(col_, "col"), D1 = f32c(D1_, "D1"), D2 = f32c(D2_, "D2");
  TORCH_CHECK(D1.dim() == 2 && D2.dim() == 2 && D1.size(1) == D2.size(1), "D1 [M, K] and D2 [*, K] must share K");
  Tensor E;
  if (E_.defined()) E = i32c(E_, "E");
  const int64_t M = rowptr.numel() - 1, K = D1.size(1), nnz = col.numel();
  const c10::cuda::CUDAGuard guard(D1.device());
  Tensor out = at::empty({1, nnz}, D1.options());
  if (nnz > 0)
    check(dgs_sddmm_csr((int)M, (int)K, nnz, iptr(rowptr), iptr(col), fptr(D1), D1.stride(0), fptr(D2), D2.stride(0), iptr(E),
                        mean ? 1 : 0, out.data_ptr<float>(), stream_of(D1)), "dgs_sddmm_csr");
  return out;
}

Tensor sddmm_csr_op(const Tensor &rowptr, const Tensor &colind, const Tensor &D1, const Tensor &D2) {
  return sddmm_csr_impl(rowptr, colind, D1, D2, Tensor(), false);
}

Tensor sddmm_coo_op(const Tensor &rowind_, const Tensor &colind_, const Tensor &D1_, const Tensor &D2_) {
  const Tensor row = i32c(rowind_, "rowind"), col = i32c(colind_, "colind"), D1 = f32c(D1_, "D1"), D2 = f32c(D2_, "D2");
  TORCH_CHECK(D1.dim() == 2 && D2.dim() == 2 && D1.size(1) == D2.size(1), "D1 [M, K] and D2 [*, K] must share K");
  const int64_t nnz = col.numel();
  const c10::cuda::CUDAGuard guard(D1.device());
  Tensor out = at::zeros({nnz}, D1.options());   // src/cuda/spmm_cuda.cu:314
  if (nnz > 0)
    check(dgs_sddmm_coo((int)D1.size(1), nnz, iptr(row), iptr(col), fptr(D1), D1.stride(0), fptr(D2), D2.stride(0),
                        out.data_ptr<float>(), stream_of(D1)), "dgs_sddmm_coo");
  return out;
}

// stable CSR -> CSC: colptr, row, optional transposed values, optional exact permutation
std::vector<Tensor> csr2csc_impl(const Tensor &rowptr_, const Tensor &col_, const Tensor &values_, int64_t ncols, bool want_perm) {
  const Tensor rowptr = i32c(rowptr_, "rowptr"), col = i32c(col_, "colind");
  Tensor values;
  if (values_.defined()) values = f32c(values_, "values").reshape({-1});
  const int64_t M = rowptr.numel() - 1, nnz = col.numel();
  const c10::cuda::CUDAGuard guard(col.device());
  const auto io = col.options();
  Tensor colptr = at::empty({ncols + 1}, io), row = at::empty({nnz}, io);
  Tensor val_t = values.defined() ? at::empty({nnz}, values.options()) : Tensor();
  Tensor perm = want_perm ? at::empty({nnz}, io) : Tensor();
  Tensor ws = scratch(dgs_csr2csc_workspace_bytes((int)M, (int)ncols, nnz), col);
  check(dgs_csr2csc((int)M, (int)ncols, nnz, iptr(rowptr), iptr(col), fptr(values), colptr.data_ptr<int>(), row.data_ptr<int>(),
                    val_t.defined() ? val_t.data_ptr<float>() : nullptr, perm.defined() ? perm.data_ptr<int>() : nullptr,
                    ws.data_ptr(), (size_t)ws.numel(), stream_of(col)), "dgs_csr2csc");
  return {colptr, row, val_t, perm};
}

// [colptr, row, values_t] for a square matrix: csr2csc_cuda, src/cuda/spmm_cuda.cu:384-414
std::vector<Tensor> csr2csc_op(const Tensor &rowptr, const Tensor &colind, const Tensor &values) {
  auto r = csr2csc_impl(rowptr, colind, values, rowptr.numel() - 1, false);
  return {r[0], r[1], r[2]};
}
// [colptr, row, perm] with the exact int32 permutation (fixes dgsparse/storage.py:164-169)
std::vector<Tensor> csr2csc_perm_op(const Tensor &rowptr, const Tensor &colind, int64_t ncols) {
  auto r = csr2csc_impl(rowptr, colind, Tensor(), ncols, true);
  return {r[0], r[1], r[3]};
}

// values.view({-1,1}).index_select(0, csr2csc).view(-1), src/spmm.cpp:70-71
Tensor t_values(const Tensor &values, const Tensor &csr2csc, bool has_value) {
  if (!has_value) return Tensor();
  return values.reshape({-1}).index_select(0, csr2csc);
}

// grad wrt dense = A^T-shaped CSC SpMM; the CSC built by Storage has max(M, ncols) columns, dense may have more rows
Tensor csc_spmm(const Tensor &colptr, const Tensor &row, const Tensor &tv, const Tensor &grad_out, const Tensor &dense,
                const Tensor &mask) {
  Tensor g = mask.defined() ? spmm_masked(colptr, row, tv, grad_out, mask) : spmm_fwd(colptr, row, tv, grad_out, DGS_SUM, nullptr);
  const int64_t ncsc = colptr.numel() - 1, k = dense.size(0);
  if (ncsc == k) return g;
  Tensor out = at::zeros({k, grad_out.size(1)}, grad_out.options());
  const int64_t n = ncsc < k ? ncsc : k;
  out.narrow(0, 0, n).copy_(g.narrow(0, 0, n));
  return out;
}

tensor_list grads(const Tensor &grad_value, const Tensor &grad_dense) {
  return {Tensor(), Tensor(), grad_value, Tensor(), Tensor(), Tensor(), grad_dense, Tensor(), Tensor()};
}

struct SpMMSum : public torch::autograd::Function<SpMMSum> {   // src/spmm.cpp:36-81
  static Tensor forward(AutogradContext *ctx, Tensor rowptr, Tensor col, Tensor values, Tensor colptr, Tensor row, Tensor csr2csc,
                        Tensor dense, bool has_value, int64_t algorithm) {
    (void)algorithm;
    Tensor out = spmm_fwd(rowptr, col, has_value ? values : Tensor(), dense, DGS_SUM, nullptr);
    ctx->saved_data["has_value"] = has_value;
    ctx->save_for_backward({rowptr, col, values, colptr, row, csr2csc, dense});
    return out;
  }
  static tensor_list backward(AutogradContext *ctx, tensor_list grad_outs) {
    const bool has_value = ctx->saved_data["has_value"].toBool();
    auto s = ctx->get_saved_variables();
    const Tensor &rowptr = s[0], &col = s[1], &values = s[2], &colptr = s[3], &row = s[4], &csr2csc = s[5], &dense = s[6];
    const Tensor grad_out = grad_outs[0].contiguous();
    Tensor gv, gd;
    if (has_value && ctx->needs_input_grad(2)) gv = sddmm_csr_impl(rowptr, col, grad_out, dense, Tensor(), false).view_as(values);
    if (ctx->needs_input_grad(6)) gd = csc_spmm(colptr, row, t_values(values, csr2csc, has_value), grad_out, dense, Tensor());
    return grads(gv, gd);
  }
};

template <int REDUCE> struct SpMMArg : public torch::autograd::Function<SpMMArg<REDUCE>> {   // SpMMMax :96-142, SpMMMin :152-198
  static Tensor forward(AutogradContext *ctx, Tensor rowptr, Tensor col, Tensor values, Tensor colptr, Tensor row, Tensor csr2csc,
                        Tensor dense, bool has_value, int64_t algorithm) {
    (void)algorithm;
    Tensor E;
    Tensor out = spmm_fwd(rowptr, col, has_value ? values : Tensor(), dense, REDUCE, &E);
    ctx->saved_data["has_value"] = has_value;
    ctx->save_for_backward({rowptr, col, values, colptr, row, csr2csc, dense, E});
    return out;
  }
  static tensor_list backward(AutogradContext *ctx, tensor_list grad_outs) {
    const bool has_value = ctx->saved_data["has_value"].toBool();
    auto s = ctx->get_saved_variables();
    const Tensor &rowptr = s[0], &col = s[1], &values = s[2], &colptr = s[3], &row = s[4], &csr2csc = s[5], &dense = s[6], &E = s[7];
    const Tensor grad_out = grad_outs[0].contiguous();
    Tensor gv, gd;
    if (has_value && ctx->needs_input_grad(2)) gv = sddmm_csr_impl(rowptr, col, grad_out, dense, E, false).view_as(values);
    if (ctx->needs_input_grad(6)) gd = csc_spmm(colptr, row, t_values(values, csr2csc, has_value), grad_out, dense, E);
    return grads(gv, gd);
  }
};

struct SpMMMean : public torch::autograd::Function<SpMMMean> {   // src/spmm.cpp:208-253
  static Tensor forward(AutogradContext *ctx, Tensor rowptr, Tensor col, Tensor values, Tensor colptr, Tensor row, Tensor csr2csc,
                        Tensor dense, bool has_value, int64_t algorithm) {
    (void)algorithm;
    Tensor out = spmm_fwd(rowptr, col, has_value ? values : Tensor(), dense, DGS_MEAN, nullptr);
    ctx->saved_data["has_value"] = has_value;
    ctx->save_for_backward({rowptr, col, values, colptr, row, csr2csc, dense});
    return out;
  }
  static tensor_list backward(AutogradContext *ctx, tensor_list grad_outs) {
    const bool has_value = ctx->saved_data["has_value"].toBool();
    auto s = ctx->get_saved_variables();
    const Tensor &rowptr = s[0], &col = s[1], &values = s[2], &colptr = s[3], &row = s[4], &csr2csc = s[5], &dense = s[6];
    const Tensor grad_out = grad_outs[0].contiguous();
    Tensor gv, gd;
    if (has_value && ctx->needs_input_grad(2)) gv = sddmm_csr_impl(rowptr, col, grad_out, dense, Tensor(), true).view_as(values);
    if (ctx->needs_input_grad(6)) {
      // d out[r] / d dense[c] = val(r, c) / deg(r): scale each CSC entry by its source row's degree
      const int64_t M = rowptr.numel() - 1;
      Tensor deg = (rowptr.narrow(0, 1, M) - rowptr.narrow(0, 0, M)).to(at::kFloat).clamp_min_(1.0);
      Tensor inv = deg.reciprocal_().index_select(0, row);
      Tensor tv = t_values(values, csr2csc, has_value);
      tv = tv.defined() ? tv * inv : inv;
      gd = csc_spmm(colptr, row, tv, grad_out, dense, Tensor());
    }
    return grads(gv, gd);
  }
};

Tensor spmm_sum(Tensor rowptr, Tensor col, Tensor values, Tensor colptr, Tensor row, Tensor csr2csc, Tensor dense, bool has_value,
                int64_t algorithm) {
  return SpMMSum::apply(rowptr, col, values, colptr, row, csr2csc, dense, has_value, algorithm);
}
Tensor spmm_max(Tensor rowptr, Tensor col, Tensor values, Tensor colptr, Tensor row, Tensor csr2csc, Tensor dense, bool has_value,
                int64_t algorithm) {
  return SpMMArg<DGS_MAX>::apply(rowptr, col, values, colptr, row, csr2csc, dense, has_value, algorithm);
}
Tensor spmm_min(Tensor rowptr, Tensor col, Tensor values, Tensor colptr, Tensor row, Tensor csr2csc, Tensor dense, bool has_value,
                int64_t algorithm) {
  return SpMMArg<DGS_MIN>::apply(rowptr, col, values, colptr, row, csr2csc, dense, has_value, algorithm);
}
Tensor spmm_mean(Tensor rowptr, Tensor col, Tensor values, Tensor colptr, Tensor row, Tensor csr2csc, Tensor dense, bool has_value,
                 int64_t algorithm) {
  return SpMMMean::apply(rowptr, col, values, colptr, row, csr2csc, dense, has_value, algorithm);
}

}  // namespace

// Same registration as src/spmm.cpp:264-270 (schemas inferred from the C++ signatures), plus the three helper ops.
TORCH_LIBRARY(dgsparse_spmm, m) {
  m.def("spmm_sum", &spmm_sum);
  m.def("spmm_max", &spmm_max);
  m.def("spmm_min", &spmm_min);
  m.def("spmm_mean", &spmm_mean);
  m.def("csr2csc", &csr2csc_op);
  m.def("csr2csc_perm", &csr2csc_perm_op);
  m.def("sddmm_csr", &sddmm_csr_op);
  m.def("sddmm_coo", &sddmm_coo_op);
}
