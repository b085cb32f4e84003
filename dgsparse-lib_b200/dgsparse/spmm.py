"""spmm_{sum,max,min,mean}(sparse, dense, algorithm) — mirror of dgsparse/spmm.py:5-106: unpack the
SparseTensor's storage into the nine positional arguments of torch.ops.dgsparse_spmm.<op>."""
import torch

from .tensor import SparseTensor


def _unpack(sparse: SparseTensor):
    s = sparse.storage
    return s.rowptr(), s.col(), s.values(), s.colptr(), s.row(), s.csr2csc()


def spmm_sum(sparse: SparseTensor, dense: torch.Tensor, algorithm=0) -> torch.Tensor:
    r"""Sparse @ dense with sum reduction."""
    return torch.ops.dgsparse_spmm.spmm_sum(*_unpack(sparse), dense, sparse.has_value, algorithm)


def spmm_mean(sparse: SparseTensor, dense: torch.Tensor, algorithm=0) -> torch.Tensor:
    r"""Sparse @ dense with mean reduction (divides by the row's nonzero count)."""
    return torch.ops.dgsparse_spmm.spmm_mean(*_unpack(sparse), dense, sparse.has_value, algorithm)


def spmm_max(sparse: SparseTensor, dense: torch.Tensor, algorithm=0) -> torch.Tensor:
    r"""Sparse @ dense with max reduction (empty rows give 0)."""
    return torch.ops.dgsparse_spmm.spmm_max(*_unpack(sparse), dense, sparse.has_value, algorithm)


def spmm_min(sparse: SparseTensor, dense: torch.Tensor, algorithm=0) -> torch.Tensor:
    r"""Sparse @ dense with min reduction (empty rows give 0)."""
    return torch.ops.dgsparse_spmm.spmm_min(*_unpack(sparse), dense, sparse.has_value, algorithm)
