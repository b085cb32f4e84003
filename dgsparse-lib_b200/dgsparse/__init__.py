"""dgsparse (B200-native) — same public surface as the reference package (dgsparse/__init__.py:1-49):
spmm_sum/max/min/mean, SparseTensor, Storage, csr2csc, torch.ops.dgsparse_spmm.*, dgsparse._C.

The CUDA library is loaded eagerly; importing this package without it raises ImportError (no CPU path).

torch.ops.dgsparse_spmm.* comes from the COMPILED op library dgsparse/_spmm_cuda.so (csrc/torch_ops.cpp), found with
PathFinder and loaded with torch.ops.load_library exactly as dgsparse/__init__.py:16-26 of the reference does; the same
file serves TorchScript and libtorch C++ callers.  When it has not been built (build.py build_torch), or with
DGSPARSE_PY_OPS=1, the same schemas are registered from Python instead (dgsparse/_ops.py, over ctypes).
"""
import importlib.machinery
import os
import os.path as osp

import torch

from . import _lib  # noqa: F401  (loads libdgsparse_b200.so first: _spmm_cuda.so and _C.so link against it)
from . import _C  # noqa: F401  (the pybind extension _C.so when built, else the ctypes mirror _C.py)

_spec = importlib.machinery.PathFinder().find_spec("_spmm_cuda", [osp.dirname(__file__)])
if _spec is not None and not os.environ.get("DGSPARSE_PY_OPS"):
    torch.ops.load_library(_spec.origin)
    ops_backend = "compiled (" + osp.basename(_spec.origin) + ")"
else:
    from . import _ops  # noqa: F401  (registers torch.ops.dgsparse_spmm.* from Python)
    ops_backend = "python (_ops.py)"
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
