"""SDDMM entry points (torch face of src/cuda/spmm_cuda.cu:305-361; the reference never registers
them as ops but its backward calls them)."""
from . import _kernels as K


def sddmm_csr(rowptr, col, D1, D2, reduce="sum"):
    """[1, nnz]: out[e] = dot(D1[row(e)], D2[col(e)]); reduce='mean' divides by the row degree."""
    return K.sddmm_csr(rowptr, col, D1, D2, mean=(reduce == "mean"))


def sddmm_coo(row, col, D1, D2):
    """[nnz]"""
    return K.sddmm_coo(row, col, D1, D2)
