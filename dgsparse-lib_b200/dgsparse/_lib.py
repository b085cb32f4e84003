"""ctypes binding of libdgsparse_b200.so (the C ABI of include/dgsparse_b200.h + include/dgsparse.h).

There is NO fallback: if the shared library is missing this module raises ImportError telling the
user how to build it.  Nothing here (or anywhere in this package) imports oracle/.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libdgsparse_b200.so")

_vp, _i32, _i64, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_size_t

SUM, MAX, MIN, MEAN = 0, 1, 2, 3                    # include/gspmm.h:13
ADD, SUB, MUL, DIV, COPY, MASKMUL = 0, 1, 2, 3, 4, 5  # include/gspmm.h:14 (+ COPY / MASKMUL)
SPCONV_FP32, SPCONV_TF32, SPCONV_BF16, SPCONV_FP16 = 0, 1, 2, 3       # include/dgsparse_b200.h dgsSpconvPrecision

# every symbol include/*.h declares: (restype, argtypes)
SIGNATURES = {
    "dgs_version": (_i32, []),
    "dgs_cuda_version": (_i32, []),
    "dgs_last_error": (ctypes.c_char_p, []),
    "dgs_sm_count": (_i32, []),
    "dgs_spmm_workspace_bytes": (_sz, [_i32, _i64, _i32]),
    "dgs_spmm_csr": (_i32, [_i32, _i32, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _i64, _i32, _i32, _vp, _sz, _vp]),
    "dgs_spmm_last_path": (_i32, []),
    "dgs_spmm_forget_graph_notes": (None, []),
    "dgs_sddmm_last_geometry": (None, [_vp, _vp, _vp]),
    "dgs_legacy_scratch_release": (_i32, []),
    "dgs_set_option": (_i32, [ctypes.c_char_p, _i32]),
    "dgs_spmm_csr_k": (_i32, [_i32, _i32, _i32, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _i64, _i32, _i32, _vp, _sz, _vp]),
    "dgs_spmm_csr_multi": (_i32, [_i32, _i32, _i64, _vp, _vp, _vp, _vp, _i64, _i32, _vp, _i64, _i32, _i32, _vp, _sz, _vp]),
    "dgs_spmm_csr_mcast": (_i32, [_i32, _i32, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _i32, _i32, _vp, _sz, _vp]),
    "dgs_mcast_barrier": (_i32, [_vp, _vp, ctypes.c_uint, _vp]),
    "dgs_spmm_csr_mask": (_i32, [_i32, _i32, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _sz, _vp]),
    "dgs_sddmm_csr": (_i32, [_i32, _i32, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _i32, _vp, _vp]),
    "dgs_sddmm_coo": (_i32, [_i32, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _vp]),
    "dgs_csr2csc_workspace_bytes": (_sz, [_i32, _i32, _i64]),
    "dgs_csr2csc": (_i32, [_i32, _i32, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "dgs_edge_softmax": (_i32, [_i32, _i32, _vp, _vp, _vp, _vp]),
    "dgs_spconv_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32, _i32]),
    "dgs_spconv_fwd": (_i32, [_i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i32, _i32, _vp, _sz, _vp]),
    "dgs_spconv_bwd": (_i32, [_i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i32, _i32,
                              _vp, _sz, _vp]),
    "dgs_kmap_workspace_bytes": (_sz, [_i32, _i32, _i32]),
    "dgs_kmap_downsample": (_i32, [_i32, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _sz, _vp]),
    "dgs_kmap_build": (_i32, [_i32, _vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp,
                              _vp, _sz, _vp]),
    "dgs_kmap_expand_workspace_bytes": (_sz, [_i32, _i32]),
    "dgs_kmap_downsample_expand": (_i32, [_i32, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp,
                                          _vp, _sz, _vp]),
    "dgs_kmap_build_ex": (_i32, [_i32, _vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32,
                                 _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "dgs_ipc_export": (_i32, [_vp, _vp, ctypes.POINTER(_i64)]),
    "dgs_ipc_open": (_i32, [_vp, ctypes.POINTER(_vp)]),
    "dgs_ipc_close": (_i32, [_vp]),
    "dgs_profile_enable": (_i32, [_i32]),
    "dgs_profile_collect": (_i32, [_i32, _vp, _vp]),
    "dgs_csr_upload": (_i32, [_i32, _i32, _i64, _vp, _vp, _vp, ctypes.POINTER(ctypes.c_void_p)]),
    "dgs_spmm_csr_resident_host": (_i32, [_vp, _i32, _vp, _vp, _vp, _i32, _i32]),
    "dgs_csr_free": (_i32, [_vp]),
    "dgs_spmm_csr_host": (_i32, [_i32, _i32, _i32, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32]),
    "dgs_sddmm_csr_host": (_i32, [_i32, _i32, _i32, _i64, _vp, _vp, _vp, _vp, _vp]),
    # legacy dgSPARSE symbols (include/dgsparse.h)
    "spmm_cuda": (None, [_i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "spmm_cuda_no_edge_value": (None, [_i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "sddmm_cuda_coo": (None, [_i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "sddmm_cuda_csr": (None, [_i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "edge_softmax_cuda": (None, [_i32, _i32, _vp, _vp, _vp]),
    # the older SpMV/SpMM API, src/ge-spmm/gespmm_v2.h
    "cuda_csr_coo_spmm": (None, [_i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cuda_csr_spmm": (None, [_i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
}


class SpMatCsrDescr_t(ctypes.Structure):  # src/ge-spmm/gespmm.h:9-16
    _fields_ = [("nrow", _i32), ("ncol", _i32), ("nnz", _i32), ("indptr", _vp), ("indices", _vp), ("data", _vp)]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"dgsparse (B200): CUDA library not found at {LIB_PATH}. Build it with "
            f"`python dgsparse-lib_b200/build.py` (needs nvcc, sm_100a). There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    lib.gespmmCsrSpMM.restype = None
    lib.gespmmCsrSpMM.argtypes = [SpMatCsrDescr_t, _vp, _i32, _vp, ctypes.c_bool, _i32]
    return lib


lib = _load()


def check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed: {lib.dgs_last_error().decode()} (cudaError {rc})")


def ptr(t):
    return None if t is None else t.data_ptr()


def stream_of(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("dgsparse (B200) ops need CUDA tensors; there is no CPU path "
                               f"(got a tensor on {t.device})")
