"""Storage — mirror of dgsparse/storage.py:6-174 of the reference (same constructor, accessors and
validation), with one fix: the CSR->CSC permutation is an exact int32 array from our own transpose
instead of a float32 arange pushed through cuSPARSE (storage.py:159-174 is exact only below 2^24
nonzeros)."""
from typing import Optional

import torch


class Storage(object):
    _row: Optional[torch.Tensor]
    _rowptr: Optional[torch.Tensor]
    _col: Optional[torch.Tensor]
    _values: Optional[torch.Tensor]
    _colptr: torch.Tensor
    _csr2csc: torch.Tensor
    _colcount: Optional[torch.Tensor]

    def __init__(self, row=None, rowptr=None, col=None, values=None, colptr=None, csr2csc=None, csc2csr=None,
                 colcount=None):
        assert row is not None or rowptr is not None
        assert col is not None
        assert col.dtype == torch.int
        assert col.dim() == 1
        col = col.contiguous()

        M = 0
        if rowptr is not None:
            M = rowptr.numel() - 1
        elif row is not None and row.numel() > 0:
            M = int(row.max()) + 1
        N = int(col.max()) + 1 if col.numel() > 0 else 0
        self.sparse_sizes = (M, N)
        self.nnz = col.size(0)

        if row is not None:
            assert row.dtype == torch.int
            assert row.device == col.device
            assert row.dim() == 1
            assert row.numel() == col.numel()
            row = row.contiguous()
        if rowptr is not None:
            assert rowptr.dtype == torch.int
            assert rowptr.device == col.device
            assert rowptr.dim() == 1
            assert rowptr.numel() - 1 == self.sparse_sizes[0]
            rowptr = rowptr.contiguous()
        else:
            # COO rows -> rowptr; the reference leaves rowptr None and then fails in csr2csc.  (row, col, values) must
            # already be in CSR order (row non-decreasing): bincount + cumsum is only a row pointer for sorted rows.
            if row.numel() > 1 and not bool((row[1:] >= row[:-1]).all()):
                raise ValueError("Storage: COO `row` must be sorted (non-decreasing); sort (row, col, values) together first")
            counts = torch.bincount(row.long(), minlength=M)
            rowptr = torch.zeros(M + 1, dtype=torch.int32, device=col.device)
            rowptr[1:] = torch.cumsum(counts, 0).to(torch.int32)
        if values is not None:
            assert values.device == col.device
            assert values.size(0) == self.nnz
            values = values.contiguous()
        else:
            values = torch.ones((self.nnz), dtype=torch.float, device=col.device)
        if colptr is not None:
            assert colptr.device == col.device
            assert colptr.dim() == 1
            colptr = colptr.contiguous()
        if csr2csc is not None:
            assert csr2csc.device == col.device
            assert csr2csc.dim() == 1
            assert csr2csc.numel() == col.size(0)
            csr2csc = csr2csc.contiguous()
        if colcount is not None:
            assert colcount.device == col.device
            assert colcount.dim() == 1
            colcount = colcount.contiguous()

        self._row = row
        self._rowptr = rowptr
        self._col = col
        self._values = values
        self._colptr = colptr
        self._csr2csc = csr2csc
        self._colcount = colcount
        self.csr2csc_convert()

    @classmethod
    def empty(self):
        row = torch.tensor([], dtype=torch.int)
        col = torch.tensor([], dtype=torch.int)
        return Storage(row=row, rowptr=None, col=col, values=None, colptr=None, csc2csr=None, csr2csc=None,
                       colcount=None)

    def _get(self, name):
        v = getattr(self, name)
        if v is None:
            raise ValueError
        return v

    def row(self) -> torch.Tensor:
        return self._get("_row")

    def rowptr(self) -> torch.Tensor:
        return self._get("_rowptr")

    def col(self) -> torch.Tensor:
        return self._get("_col")

    def colptr(self) -> torch.Tensor:
        return self._get("_colptr")

    def values(self) -> torch.Tensor:
        return self._get("_values")

    def csr2csc(self) -> torch.Tensor:
        return self._get("_csr2csc")

    def csr2csc_convert(self):
        """Eager CSC build, as dgsparse/storage.py:100,159-174: sets _colptr, _row (CSC row indices)
        and _csr2csc (CSC position -> CSR position).

        Deviation: `_row` is ALWAYS replaced by the CSC-ordered row indices, because that is what the backward kernels read
        it as (src/spmm.cpp:52-80 passes it as the CSC index array).  The reference keeps a user-supplied COO `row`
        (`if self._row is None`), i.e. CSR-ordered rows, and its backward then gathers with the wrong indices; after this
        call `row()` therefore returns CSC order here, whatever was passed to the constructor."""
        if self._csr2csc is not None:
            return self._csr2csc
        if not self._col.is_cuda:
            # host-side containers (e.g. Storage.empty()) carry no CSC; ops need CUDA tensors anyway
            return None
        ncols = max(self.sparse_sizes[0], self.sparse_sizes[1])
        colptr, row, csr2csc = torch.ops.dgsparse_spmm.csr2csc_perm(self._rowptr, self._col, ncols)
        self._row = row
        if self._colptr is None:
            self._colptr = colptr
        self._csr2csc = csr2csc
        return csr2csc
