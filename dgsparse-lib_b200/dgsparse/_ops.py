"""torch.ops.dgsparse_spmm.* — the reference's operator boundary (src/spmm.cpp:264-270), re-registered
over the B200 C ABI.  Same five schemas, same argument order, same autograd contract (gradients for
`values` (arg 2) and `dense` (arg 6) only, src/spmm.cpp:76-78).

`algorithm` is accepted and ignored: every value gives algorithm-0 semantics.
Deviations from the reference, all fixes: E is only materialised for max/min (q6);
the mean backward wrt dense scales by the degree of the SOURCE row, which is the true gradient
(the reference divides by the column degree, src/spmm.cpp:245-246, right only for regular graphs).
"""
import torch

from . import _kernels as K
from . import _lib

_SCHEMA = ("(Tensor rowptr, Tensor col, Tensor values, Tensor colptr, Tensor row, Tensor csr2csc, "
           "Tensor dense, bool has_value, int algorithm) -> Tensor")

_library = torch.library.Library("dgsparse_spmm", "DEF")
for _name in ("spmm_sum", "spmm_max", "spmm_min", "spmm_mean"):
    _library.define(_name + _SCHEMA)
_library.define("csr2csc(Tensor rowptr, Tensor colind, Tensor values) -> Tensor[]")
_library.define("csr2csc_perm(Tensor rowptr, Tensor colind, int ncols) -> Tensor[]")
_library.define("sddmm_csr(Tensor rowptr, Tensor colind, Tensor D1, Tensor D2) -> Tensor")
_library.define("sddmm_coo(Tensor rowind, Tensor colind, Tensor D1, Tensor D2) -> Tensor")


def _t_values(values, csr2csc, has_value):
    # values.view({-1,1}).index_select(0, csr2csc).view(-1), src/spmm.cpp:70-71
    if not has_value:
        return None
    return values.reshape(-1).index_select(0, csr2csc.long() if csr2csc.dtype != torch.int32 else csr2csc)


def _grad_dense_rows(colptr, dense):
    """The CSC built by Storage has max(M, ncols) columns; dense may have more rows than that."""
    return colptr.numel() - 1, dense.size(0)


def _csc_spmm(colptr, row, t_values, grad_out, dense, reduce=_lib.SUM, mask=None):
    ncsc, k = _grad_dense_rows(colptr, dense)
    if mask is None:
        g = K.spmm(colptr, row, t_values, grad_out, reduce, _lib.MUL)
    else:
        g = K.spmm_with_mask(colptr, row, t_values, grad_out, mask)
    if ncsc == k:
        return g
    out = grad_out.new_zeros((k, grad_out.size(1)))
    n = min(ncsc, k)
    out[:n] = g[:n]
    return out


class SpMMSum(torch.autograd.Function):  # src/spmm.cpp:36-81
    @staticmethod
    def forward(ctx, rowptr, col, values, colptr, row, csr2csc, dense, has_value, algorithm):
        out = K.spmm(rowptr, col, values if has_value else None, dense, _lib.SUM, _lib.MUL)
        ctx.has_value = has_value
        ctx.save_for_backward(rowptr, col, values, colptr, row, csr2csc, dense)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        rowptr, col, values, colptr, row, csr2csc, dense = ctx.saved_tensors
        grad_out = grad_out.contiguous()
        grad_value = grad_dense = None
        if ctx.has_value and ctx.needs_input_grad[2]:
            grad_value = K.sddmm_csr(rowptr, col, grad_out, dense).view_as(values)
        if ctx.needs_input_grad[6]:
            grad_dense = _csc_spmm(colptr, row, _t_values(values, csr2csc, ctx.has_value), grad_out, dense)
        return None, None, grad_value, None, None, None, grad_dense, None, None


class _SpMMArg(torch.autograd.Function):  # SpMMMax src/spmm.cpp:96-142, SpMMMin :152-198
    REDUCE = _lib.MAX

    @classmethod
    def _fwd(cls, ctx, rowptr, col, values, colptr, row, csr2csc, dense, has_value, algorithm):
        out, E = K.spmm(rowptr, col, values if has_value else None, dense, cls.REDUCE, _lib.MUL, with_arg=True)
        ctx.has_value = has_value
        ctx.save_for_backward(rowptr, col, values, colptr, row, csr2csc, dense, E)
        ctx.mark_non_differentiable(E)
        return out

    @staticmethod
    def _bwd(ctx, grad_out):
        rowptr, col, values, colptr, row, csr2csc, dense, E = ctx.saved_tensors
        grad_out = grad_out.contiguous()
        grad_value = grad_dense = None
        if ctx.has_value and ctx.needs_input_grad[2]:
            grad_value = K.sddmm_csr(rowptr, col, grad_out, dense, E=E).view_as(values)
        if ctx.needs_input_grad[6]:
            grad_dense = _csc_spmm(colptr, row, _t_values(values, csr2csc, ctx.has_value), grad_out, dense, mask=E)
        return None, None, grad_value, None, None, None, grad_dense, None, None


class SpMMMax(_SpMMArg):
    REDUCE = _lib.MAX

    @staticmethod
    def forward(ctx, *args):
        ctx.set_materialize_grads(True)
        return SpMMMax._fwd(ctx, *args)

    @staticmethod
    def backward(ctx, grad_out):
        return _SpMMArg._bwd(ctx, grad_out)


class SpMMMin(_SpMMArg):
    REDUCE = _lib.MIN

    @staticmethod
    def forward(ctx, *args):
        ctx.set_materialize_grads(True)
        return SpMMMin._fwd(ctx, *args)

    @staticmethod
    def backward(ctx, grad_out):
        return _SpMMArg._bwd(ctx, grad_out)


class SpMMMean(torch.autograd.Function):  # src/spmm.cpp:208-253
    @staticmethod
    def forward(ctx, rowptr, col, values, colptr, row, csr2csc, dense, has_value, algorithm):
        out = K.spmm(rowptr, col, values if has_value else None, dense, _lib.MEAN, _lib.MUL)
        ctx.has_value = has_value
        ctx.save_for_backward(rowptr, col, values, colptr, row, csr2csc, dense)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        rowptr, col, values, colptr, row, csr2csc, dense = ctx.saved_tensors
        grad_out = grad_out.contiguous()
        grad_value = grad_dense = None
        if ctx.has_value and ctx.needs_input_grad[2]:
            grad_value = K.sddmm_csr(rowptr, col, grad_out, dense, mean=True).view_as(values)
        if ctx.needs_input_grad[6]:
            # d out[r] / d dense[c] = val(r,c) / deg(r): scale each CSC entry by its source row's degree
            deg = (rowptr[1:] - rowptr[:-1]).to(torch.float32)
            inv = torch.reciprocal(deg.clamp_(min=1.0)).index_select(0, row.long())
            tv = _t_values(values, csr2csc, ctx.has_value)
            tv = inv if tv is None else tv * inv
            grad_dense = _csc_spmm(colptr, row, tv, grad_out, dense)
        return None, None, grad_value, None, None, None, grad_dense, None, None


def _spmm_sum(rowptr, col, values, colptr, row, csr2csc, dense, has_value, algorithm):
    return SpMMSum.apply(rowptr, col, values, colptr, row, csr2csc, dense, has_value, algorithm)


def _spmm_max(rowptr, col, values, colptr, row, csr2csc, dense, has_value, algorithm):
    return SpMMMax.apply(rowptr, col, values, colptr, row, csr2csc, dense, has_value, algorithm)


def _spmm_min(rowptr, col, values, colptr, row, csr2csc, dense, has_value, algorithm):
    return SpMMMin.apply(rowptr, col, values, colptr, row, csr2csc, dense, has_value, algorithm)


def _spmm_mean(rowptr, col, values, colptr, row, csr2csc, dense, has_value, algorithm):
    return SpMMMean.apply(rowptr, col, values, colptr, row, csr2csc, dense, has_value, algorithm)


def _csr2csc(rowptr, colind, values):
    """[colptr, row, values_t] for a square matrix, as csr2csc_cuda (src/cuda/spmm_cuda.cu:384-414)."""
    colptr, row, val_t, _ = K.csr2csc(rowptr, colind, values, ncols=rowptr.numel() - 1, want_perm=False)
    return [colptr, row, val_t]


def _csr2csc_perm(rowptr, colind, ncols):
    """[colptr, row, perm] with the exact int32 permutation (fixes dgsparse/storage.py:164-169)."""
    colptr, row, _, perm = K.csr2csc(rowptr, colind, None, ncols=ncols, want_perm=True)
    return [colptr, row, perm]


def _sddmm_csr(rowptr, colind, D1, D2):
    return K.sddmm_csr(rowptr, colind, D1, D2)


def _sddmm_coo(rowind, colind, D1, D2):
    return K.sddmm_coo(rowind, colind, D1, D2)


_library.impl("spmm_sum", _spmm_sum, "CompositeImplicitAutograd")
_library.impl("spmm_max", _spmm_max, "CompositeImplicitAutograd")
_library.impl("spmm_min", _spmm_min, "CompositeImplicitAutograd")
_library.impl("spmm_mean", _spmm_mean, "CompositeImplicitAutograd")
_library.impl("csr2csc", _csr2csc, "CompositeExplicitAutograd")
_library.impl("csr2csc_perm", _csr2csc_perm, "CompositeExplicitAutograd")
_library.impl("sddmm_csr", _sddmm_csr, "CompositeExplicitAutograd")
_library.impl("sddmm_coo", _sddmm_coo, "CompositeExplicitAutograd")
