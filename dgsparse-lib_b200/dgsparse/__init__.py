"""dgsparse (B200-native) — same public surface as the reference package (dgsparse/__init__.py:1-49):
spmm_sum/max/min/mean, SparseTensor, Storage, csr2csc, torch.ops.dgsparse_spmm.*, dgsparse._C.

The CUDA library is loaded eagerly; importing this package without it raises ImportError (no CPU path).
"""
import torch

from . import _C  # noqa: F401
from . import _ops  # noqa: F401  (registers torch.ops.dgsparse_spmm.*)
from .spmm import spmm_max, spmm_mean, spmm_sum, spmm_min
from .tensor import SparseTensor
from .storage import Storage
from .ftransform import csr2csc
from . import gspmm, sddmm, spconv, sparse_mapping  # noqa: F401  (spconv registers torch.ops.dgsparse_spconv.spconv)
from . import nn  # noqa: F401,E402  (dgsparse/__init__.py:44 `from . import nn`; after SparseTensor / spmm_* exist)

__version__ = "0.1+b200"

cuda_version = _C.cuda_version()
if torch.version.cuda is not None and cuda_version != -1:  # dgsparse/__init__.py:28-42
    major = cuda_version // 1000
    t_major = int(torch.version.cuda.split(".")[0])
    if t_major != major:
        raise RuntimeError(f"PyTorch was built with CUDA {torch.version.cuda} but dgsparse (B200) with CUDA "
                           f"{major}.{(cuda_version % 1000) // 10}; the major versions must match.")

__all__ = ["spmm_sum", "spmm_max", "spmm_min", "spmm_mean", "Storage", "SparseTensor", "csr2csc"]
