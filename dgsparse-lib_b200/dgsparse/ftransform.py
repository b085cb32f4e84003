"""csr2csc(sparse) — mirror of dgsparse/ftransform.py:6-10."""
from typing import Tuple

import torch

from .tensor import SparseTensor


def csr2csc(sparse: SparseTensor) -> Tuple[torch.Tensor]:
    s = sparse.storage
    return torch.ops.dgsparse_spmm.csr2csc(s._rowptr, s._col, s._values)
