"""dgsparse._C — mirror of src/version.cpp:11-21."""
from ._lib import lib


def cuda_version() -> int:
    return int(lib.dgs_cuda_version())
