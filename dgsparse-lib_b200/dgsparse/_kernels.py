"""Tensor-level launchers over the C ABI: allocate outputs/workspace with torch (plumbing), pass raw
pointers and the current stream to libdgsparse_b200.so.  Mirrors the host launchers of the reference
torch face, src/cuda/spmm_cuda.cu (spmm_cuda :14, spmm_cuda_with_mask :255, sddmm_cuda_coo :305,
sddmm_cuda_csr :331, sddmm_cuda_csr_with_mask :363, csr2csc_cuda :384)."""
import torch

from . import _lib
from ._lib import lib, check, ptr, stream_of, require_cuda


def _i32c(t, name):
    if t.dtype != torch.int32:
        raise TypeError(f"{name} must be int32 (dgsparse/storage.py:27-60), got {t.dtype}")
    return t.contiguous()


def _f32c(t, name):
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    return t.contiguous()


def _workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def spmm(rowptr, col, values, dense, reduce=_lib.SUM, compute=_lib.MUL, with_arg=False, out=None):
    """out[,E] = generalized CSR SpMM.  values=None -> no edge value."""
    require_cuda(rowptr, col, values, dense)
    rowptr, col = _i32c(rowptr, "rowptr"), _i32c(col, "col")
    dense = _f32c(dense, "dense")
    if dense.dim() != 2:
        raise ValueError("dense must be 2-D [K, N]")
    if values is not None:
        values = _f32c(values, "values").reshape(-1)
        if values.numel() != col.numel():
            raise ValueError("values must have one entry per nonzero")
    M, N, nnz = rowptr.numel() - 1, dense.size(1), col.numel()
    with torch.cuda.device(dense.device):
        if out is None:
            out = torch.empty((M, N), dtype=torch.float32, device=dense.device)
        E = torch.empty((M, N), dtype=torch.int32, device=dense.device) if with_arg else None
        if M == 0 or N == 0:
            return (out, E) if with_arg else out
        ws_bytes = lib.dgs_spmm_workspace_bytes(N, nnz, int(with_arg))
        ws = _workspace(ws_bytes, dense.device)
        check(lib.dgs_spmm_csr_k(M, dense.size(0), N, nnz, ptr(rowptr), ptr(col), ptr(values), ptr(dense), dense.stride(0),
                                 ptr(out), out.stride(0), ptr(E), N if with_arg else 0, int(reduce), int(compute),
                                 ptr(ws), ws.numel(), stream_of(dense)), "dgs_spmm_csr_k")
    return (out, E) if with_arg else out


def spmm_with_mask(ptr_, idx, values, grad, E):
    """Max/min backward wrt dense on the CSC arrays (spmm_cuda_with_mask, src/cuda/spmm_cuda.cu:255-303)."""
    require_cuda(ptr_, idx, values, grad, E)
    ptr_, idx, grad, E = _i32c(ptr_, "ptr"), _i32c(idx, "idx"), _f32c(grad, "grad"), _i32c(E, "E")
    if values is not None:
        values = _f32c(values, "values").reshape(-1)
    M, N, nnz = ptr_.numel() - 1, grad.size(1), idx.numel()
    with torch.cuda.device(grad.device):
        out = torch.empty((M, N), dtype=torch.float32, device=grad.device)
        if M == 0 or N == 0:
            return out
        ws_bytes = lib.dgs_spmm_workspace_bytes(N, nnz, 0)
        ws = _workspace(ws_bytes, grad.device)
        check(lib.dgs_spmm_csr_mask(M, N, nnz, ptr(ptr_), ptr(idx), ptr(values), ptr(grad), grad.stride(0), ptr(E),
                                    E.stride(0), ptr(out), N, ptr(ws), ws.numel(), stream_of(grad)),
              "dgs_spmm_csr_mask")
    return out


def sddmm_csr(rowptr, col, D1, D2, mean=False, E=None):
    """[1, nnz] like the reference torch face (src/cuda/spmm_cuda.cu:342)."""
    require_cuda(rowptr, col, D1, D2, E)
    rowptr, col, D1, D2 = _i32c(rowptr, "rowptr"), _i32c(col, "col"), _f32c(D1, "D1"), _f32c(D2, "D2")
    if E is not None:
        E = _i32c(E, "E")
    M, K, nnz = rowptr.numel() - 1, D1.size(1), col.numel()
    with torch.cuda.device(D1.device):
        out = torch.empty((1, nnz), dtype=torch.float32, device=D1.device)
        if nnz:
            check(lib.dgs_sddmm_csr(M, K, nnz, ptr(rowptr), ptr(col), ptr(D1), D1.stride(0), ptr(D2), D2.stride(0),
                                    ptr(E), int(bool(mean)), ptr(out), stream_of(D1)), "dgs_sddmm_csr")
    return out


def sddmm_coo(row, col, D1, D2, out=None):
    """[nnz] like the reference torch face (src/cuda/spmm_cuda.cu:314).  out: optional preallocated fp32 [nnz]."""
    require_cuda(row, col, D1, D2, out)
    row, col, D1, D2 = _i32c(row, "row"), _i32c(col, "col"), _f32c(D1, "D1"), _f32c(D2, "D2")
    K, nnz = D1.size(1), col.numel()
    with torch.cuda.device(D1.device):
        if out is None:
            out = torch.zeros((nnz,), dtype=torch.float32, device=D1.device)
        elif out.dtype != torch.float32 or out.numel() != nnz or not out.is_contiguous():
            raise TypeError("out must be a contiguous float32 tensor with nnz elements")
        if nnz:
            check(lib.dgs_sddmm_coo(K, nnz, ptr(row), ptr(col), ptr(D1), D1.stride(0), ptr(D2), D2.stride(0),
                                    ptr(out), stream_of(D1)), "dgs_sddmm_coo")
    return out


def csr2csc(rowptr, col, values=None, ncols=None, want_perm=True):
    """-> (colptr int32[ncols+1], row int32[nnz], values_t f32[nnz] | None, perm int32[nnz] | None).
    ncols defaults to M (the reference assumes a square matrix, src/cuda/spmm_cuda.cu:409) but is
    widened to max(col)+1 by the caller when the matrix is not square."""
    require_cuda(rowptr, col, values)
    rowptr, col = _i32c(rowptr, "rowptr"), _i32c(col, "col")
    if values is not None:
        values = _f32c(values, "values").reshape(-1)
    M, nnz = rowptr.numel() - 1, col.numel()
    if ncols is None:
        ncols = M
    dev = col.device
    with torch.cuda.device(dev):
        colptr = torch.empty(ncols + 1, dtype=torch.int32, device=dev)
        row = torch.empty(nnz, dtype=torch.int32, device=dev)
        val_t = torch.empty(nnz, dtype=torch.float32, device=dev) if values is not None else None
        perm = torch.empty(nnz, dtype=torch.int32, device=dev) if want_perm else None
        ws_bytes = lib.dgs_csr2csc_workspace_bytes(M, ncols, nnz)
        ws = _workspace(ws_bytes, dev)
        check(lib.dgs_csr2csc(M, ncols, nnz, ptr(rowptr), ptr(col), ptr(values), ptr(colptr), ptr(row), ptr(val_t),
                              ptr(perm), ptr(ws), ws.numel(), stream_of(col)), "dgs_csr2csc")
    return colptr, row, val_t, perm


def edge_softmax(rowptr, values, head=1):
    require_cuda(rowptr, values)
    rowptr, values = _i32c(rowptr, "rowptr"), _f32c(values, "values")
    out = torch.empty_like(values)
    with torch.cuda.device(values.device):
        check(lib.dgs_edge_softmax(rowptr.numel() - 1, head, ptr(rowptr), ptr(values), ptr(out), stream_of(values)),
              "dgs_edge_softmax")
    return out
