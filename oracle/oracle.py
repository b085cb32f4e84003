"""ctypes face of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this module; nothing under dgsparse-lib_b200/ does.  See oracle/oracle.c for the reference
file:line each function restates, and oracle/ref_host_shim.cpp for the compiled reference itself.
"""
import ctypes
import os
import subprocess
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "_build", "liboracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libref_host.so")
_REF_CUDA_SO = os.path.join(_HERE, "_ref", "libref_cuda.so")

REDUCE = {"sum": 0, "max": 1, "min": 2, "mean": 3}            # include/gspmm.h:13
COMPUTE = {"add": 0, "sub": 1, "mul": 2, "div": 3, "copy": 4}  # include/gspmm.h:14 (+copy)

_i32p = ctypes.POINTER(ctypes.c_int)
_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)


def build(force=False):
    """Compile oracle/_build/liboracle.so (and oracle/_ref when /root/reference is present)."""
    if force or not os.path.exists(_ORACLE_SO) or (
            os.path.getmtime(_ORACLE_SO) < os.path.getmtime(os.path.join(_HERE, "oracle.c"))):
        subprocess.check_call(["make", "-C", _HERE, "-s", "_build/liboracle.so"])
    if os.path.exists("/root/reference/example/util/sp_util.hpp") and (
            force or not os.path.exists(_REF_SO)):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
    if os.path.exists("/root/reference/src/ge-spmm/gespmm.cc") and (
            force or not os.path.exists(_REF_CUDA_SO)):
        subprocess.check_call(["make", "-C", _HERE, "-s", "refcuda"])
    # the reference's own example drivers linked against OUR library (tests/test_reference_drivers_gpu.py)
    ours = os.path.join(os.path.dirname(_HERE), "dgsparse-lib_b200", "lib", "libdgsparse_b200.so")
    drv = os.path.join(_HERE, "_ref", "spmm_example.out")
    if os.path.exists("/root/reference/example/ge-spmm/spmm.cu") and os.path.exists(ours) and (
            force or not os.path.exists(drv) or os.path.getmtime(drv) < os.path.getmtime(
                os.path.join(os.path.dirname(_HERE), "include", "dgsparse.h"))):
        subprocess.check_call(["make", "-C", _HERE, "-s", "refdrivers"])


def _p(a, typ):
    if a is None:
        return ctypes.cast(None, typ)
    return a.ctypes.data_as(typ)


def _i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


_lib = None
_ref = None
_lock = threading.Lock()


def lib():
    global _lib
    with _lock:
        if _lib is None:
            build()
            _lib = ctypes.CDLL(_ORACLE_SO)
        return _lib


def ref_lib():
    """The reference's own host code (oracle/_ref/libref_host.so) or None if it was never built."""
    global _ref
    with _lock:
        if _ref is None:
            if not os.path.exists(_REF_SO):
                try:
                    build()
                except Exception:
                    pass
            if os.path.exists(_REF_SO):
                _ref = ctypes.CDLL(_REF_SO)
        return _ref


_ref_cuda = None


class SpMatCsrDescr(ctypes.Structure):
    """struct SpMatCsrDescr_t of the reference, src/ge-spmm/gespmm.h:9-16 (passed by value)."""
    _fields_ = [("nrow", ctypes.c_int), ("ncol", ctypes.c_int), ("nnz", ctypes.c_int),
                ("indptr", ctypes.c_void_p), ("indices", ctypes.c_void_p), ("data", ctypes.c_void_p)]


def ref_cuda_lib():
    """The reference's own CUDA C ABI compiled for sm_100a (oracle/_ref/libref_cuda.so) or None.

    Device-pointer entry points, all launching on the legacy default stream and returning void:
      spmm_cuda(m, N, rowptr, col, val, B, C)               src/ge-spmm/gespmm.cc:114-123
      spmm_cuda_no_edge_value(m, N, rowptr, col, _, B, C)    src/ge-spmm/gespmm.cc:125-134
      gespmmCsrSpMM(SpMatCsrDescr_t, B, N, C, transpose, alg) src/ge-spmm/gespmm.cc:29-111
      sddmm_cuda_csr(m, k, nnz, rowptr, col, D1, D2, out)    src/sddmm/sddmm.cu:25-41
      sddmm_cuda_coo(k, nnz, row, col, D1, D2, out)          src/sddmm/sddmm.cu:8-23
    """
    global _ref_cuda
    with _lock:
        if _ref_cuda is None and os.path.exists(_REF_CUDA_SO):
            L = ctypes.CDLL(_REF_CUDA_SO)
            vp, ci = ctypes.c_void_p, ctypes.c_int
            L.spmm_cuda.argtypes = [ci, ci, vp, vp, vp, vp, vp]
            L.spmm_cuda.restype = None
            L.spmm_cuda_no_edge_value.argtypes = [ci, ci, vp, vp, vp, vp, vp]
            L.spmm_cuda_no_edge_value.restype = None
            L.gespmmCsrSpMM.argtypes = [SpMatCsrDescr, vp, ci, vp, ctypes.c_bool, ci]
            L.gespmmCsrSpMM.restype = None
            L.sddmm_cuda_csr.argtypes = [ci, ci, ci, vp, vp, vp, vp, vp]
            L.sddmm_cuda_csr.restype = None
            L.sddmm_cuda_coo.argtypes = [ci, ci, vp, vp, vp, vp, vp]
            L.sddmm_cuda_coo.restype = None
            # older API, src/ge-spmm/gespmm_v2.h:19-37
            L.cuda_csr_coo_spmm.argtypes = [ci, ci, ci, ci, ci, ci, vp, vp, vp, vp, vp, vp]
            L.cuda_csr_coo_spmm.restype = None
            L.cuda_csr_spmm.argtypes = [ci, ci, ci, ci, ci, ci, vp, vp, vp, vp, vp]
            L.cuda_csr_spmm.restype = None
            _ref_cuda = L
        return _ref_cuda


_ref_gspmm = None


def ref_gspmm_module():
    """The reference's own gspmm-fp pybind module `spmm` (oracle/_ref/spmm.so, built by oracle/build_ref_gspmm.sh from
    src/gspmm-fp/gspmm.{cu,cc} unmodified) or None.  GSpMM_u_e(rowptr, colind, edge_val, feat, REDUCEOP, COMPUTEOP) and
    GSpMM_u(rowptr, colind, feat, REDUCEOP) on CUDA tensors (src/gspmm-fp/gspmm.cc:27-44)."""
    global _ref_gspmm
    with _lock:
        path = os.path.join(_HERE, "_ref", "spmm.so")
        if _ref_gspmm is None and os.path.exists(path):
            import importlib.util
            import torch  # noqa: F401  (libtorch must be loaded before the extension)
            spec = importlib.util.spec_from_file_location("spmm", path)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            _ref_gspmm = mod
        return _ref_gspmm


_ref_spconv = None


def ref_spconv_module():
    """The reference's own sparse-convolution CUDA behind a no-algorithm pybind shim (oracle/_ref/_ref_spconv.so, built by
    oracle/build_ref_spconv.sh from src/cuda/sparse_mapping.cu + src/cuda/spconv_cuda.cu unmodified) or None:
    sparse_mapping(...), spconv_fwd_fused(...), spconv_bwd_fused(...) on CUDA tensors."""
    global _ref_spconv
    with _lock:
        path = os.path.join(_HERE, "_ref", "_ref_spconv.so")
        if _ref_spconv is None and os.path.exists(path):
            import importlib.util
            import torch  # noqa: F401  (libtorch must be loaded before the extension)
            spec = importlib.util.spec_from_file_location("_ref_spconv", path)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            _ref_spconv = mod
        return _ref_spconv


def num_threads():
    return int(lib().oracle_num_threads())


def set_num_threads(n):
    lib().oracle_set_num_threads(int(n))


def spmm(rowptr, col, val, B, reduce="sum", compute="mul", with_arg=False):
    """out[,E] per include/cuda/spmm_cuda.cuh:27-54 (+ gspmm COMPUTE ops). val=None -> no edge value."""
    rowptr, col, B = _i32(rowptr), _i32(col), _f32(B)
    val = None if val is None else _f32(val).reshape(-1)
    M, N = rowptr.size - 1, B.shape[1]
    out = np.empty((M, N), np.float32)
    E = np.empty((M, N), np.int32) if with_arg else None
    lib().oracle_spmm(ctypes.c_int(M), ctypes.c_int(N), _p(rowptr, _i32p), _p(col, _i32p),
                      _p(val, _f32p), _p(B, _f32p), ctypes.c_int64(B.shape[1]),
                      ctypes.c_int(REDUCE[reduce]), ctypes.c_int(COMPUTE[compute]),
                      _p(out, _f32p), _p(E, _i32p))
    return (out, E) if with_arg else out


def spmm_f64(rowptr, col, val, B, reduce="sum", compute="mul"):
    rowptr, col, B = _i32(rowptr), _i32(col), _f32(B)
    val = None if val is None else _f32(val).reshape(-1)
    M, N = rowptr.size - 1, B.shape[1]
    out = np.empty((M, N), np.float64)
    lib().oracle_spmm_f64(ctypes.c_int(M), ctypes.c_int(N), _p(rowptr, _i32p), _p(col, _i32p),
                          _p(val, _f32p), _p(B, _f32p), ctypes.c_int64(B.shape[1]),
                          ctypes.c_int(REDUCE[reduce]), ctypes.c_int(COMPUTE[compute]),
                          _p(out, _f64p))
    return out


def sddmm_csr(rowptr, col, D1, D2, mean=False, f64=False):
    rowptr, col, D1, D2 = _i32(rowptr), _i32(col), _f32(D1), _f32(D2)
    M, K = rowptr.size - 1, D1.shape[1]
    out = np.empty(col.size, np.float64 if f64 else np.float32)
    fn = lib().oracle_sddmm_csr_f64 if f64 else lib().oracle_sddmm_csr
    fn(ctypes.c_int(M), ctypes.c_int(K), _p(rowptr, _i32p), _p(col, _i32p), _p(D1, _f32p),
       _p(D2, _f32p), ctypes.c_int(int(mean)), _p(out, _f64p if f64 else _f32p))
    return out


def sddmm_coo(row, col, D1, D2):
    row, col, D1, D2 = _i32(row), _i32(col), _f32(D1), _f32(D2)
    out = np.empty(col.size, np.float32)
    lib().oracle_sddmm_coo(ctypes.c_int(D1.shape[1]), ctypes.c_int64(col.size), _p(row, _i32p),
                           _p(col, _i32p), _p(D1, _f32p), _p(D2, _f32p), _p(out, _f32p))
    return out


def spmm_mask(ptr, idx, val, G, E):
    ptr, idx, G, E = _i32(ptr), _i32(idx), _f32(G), _i32(E)
    val = None if val is None else _f32(val).reshape(-1)
    M, N = ptr.size - 1, G.shape[1]
    out = np.empty((M, N), np.float32)
    lib().oracle_spmm_mask(ctypes.c_int(M), ctypes.c_int(N), _p(ptr, _i32p), _p(idx, _i32p),
                           _p(val, _f32p), _p(G, _f32p), _p(E, _i32p), _p(out, _f32p))
    return out


def sddmm_csr_mask(rowptr, col, D1, D2, E):
    rowptr, col, D1, D2, E = _i32(rowptr), _i32(col), _f32(D1), _f32(D2), _i32(E)
    M, K = rowptr.size - 1, D1.shape[1]
    out = np.empty(col.size, np.float32)
    lib().oracle_sddmm_csr_mask(ctypes.c_int(M), ctypes.c_int(K), _p(rowptr, _i32p), _p(col, _i32p),
                                _p(D1, _f32p), _p(D2, _f32p), _p(E, _i32p), _p(out, _f32p))
    return out


def csr2csc(rowptr, col, val=None, ncols=None):
    """-> (colptr, row, val_t, perm); stable (rows ascending within a column)."""
    rowptr, col = _i32(rowptr), _i32(col)
    val = None if val is None else _f32(val).reshape(-1)
    M = rowptr.size - 1
    if ncols is None:
        ncols = M  # the reference assumes square (src/cuda/spmm_cuda.cu:409)
    colptr = np.empty(ncols + 1, np.int32)
    row = np.empty(col.size, np.int32)
    perm = np.empty(col.size, np.int32)
    val_t = None if val is None else np.empty(col.size, np.float32)
    lib().oracle_csr2csc(ctypes.c_int(M), ctypes.c_int(ncols), _p(rowptr, _i32p), _p(col, _i32p),
                         _p(val, _f32p), _p(colptr, _i32p), _p(row, _i32p), _p(val_t, _f32p),
                         _p(perm, _i32p))
    return colptr, row, val_t, perm


def spconv(kpos, imap, omap, in_feats, W, out_nnz, precompute=False):
    """cpu_compute(feats, weights, out_size, knnz, imap, omap, precompute), test/test_spconv.py:17-53."""
    kpos, imap, omap, in_feats, W = _i32(kpos), _i32(imap), _i32(omap), _f32(in_feats), _f32(W)
    k_vol, c_in, c_out = W.shape
    out = np.empty((out_nnz, c_out), np.float32)
    lib().oracle_spconv(ctypes.c_int(k_vol), ctypes.c_int(c_in), ctypes.c_int(c_out),
                        ctypes.c_int(out_nnz), _p(kpos, _i32p), _p(imap, _i32p), _p(omap, _i32p),
                        _p(in_feats, _f32p), _p(W, _f32p), _p(out, _f32p), ctypes.c_int(int(bool(precompute))))
    return out


# ---- the compiled reference (oracle/_ref) -------------------------------------------------------

def ref_spmm_host(rowptr, col, val, B, K=None):
    """spmm_reference_host<int,float> (/root/reference/example/util/sp_util.hpp:62-84), unmodified."""
    r = ref_lib()
    if r is None:
        raise RuntimeError("oracle/_ref/libref_host.so not built (reference tree absent)")
    rowptr, col, val, B = _i32(rowptr), _i32(col), _f32(val), _f32(B)
    M, N = rowptr.size - 1, B.shape[1]
    out = np.empty((M, N), np.float32)
    r.ref_spmm_host(ctypes.c_int(M), ctypes.c_int(N), ctypes.c_int(K or B.shape[0]),
                    _p(rowptr, _i32p), _p(col, _i32p), _p(val, _f32p), _p(B, _f32p), _p(out, _f32p))
    return out


def ref_spmm_host_threads(rowptr, col, val, B, threads, out=None, rows=None):
    """Row-block parallel driver around the unmodified reference loop (one block per host thread,
    blocks balanced by nnz).  ctypes releases the GIL, so the blocks run concurrently."""
    r = ref_lib()
    if r is None:
        raise RuntimeError("oracle/_ref/libref_host.so not built (reference tree absent)")
    rowptr, col, val, B = _i32(rowptr), _i32(col), _f32(val), _f32(B)
    M, N = rowptr.size - 1, B.shape[1]
    r0, r1 = (0, M) if rows is None else rows
    if out is None:
        out = np.empty((M, N), np.float32)
    lo, hi = int(rowptr[r0]), int(rowptr[r1])
    targets = lo + (hi - lo) * np.arange(1, threads) / threads
    cuts = [r0] + [int(x) for x in np.searchsorted(rowptr[r0:r1 + 1], targets) + r0] + [r1]
    cuts = [min(max(c, r0), r1) for c in cuts]

    def work(a, b):
        if b > a:
            r.ref_spmm_host_rows(ctypes.c_int(a), ctypes.c_int(b), ctypes.c_int(N),
                                 ctypes.c_int(B.shape[0]), _p(rowptr, _i32p), _p(col, _i32p),
                                 _p(val, _f32p), _p(B, _f32p), _p(out, _f32p))

    ts = [threading.Thread(target=work, args=(cuts[i], cuts[i + 1])) for i in range(threads)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return out


def ref_sddmm_host(rowptr, col, D1, D2):
    """sddmm_reference_host<int,float> (/root/reference/example/util/sp_util.hpp:87-112)."""
    r = ref_lib()
    if r is None:
        raise RuntimeError("oracle/_ref/libref_host.so not built (reference tree absent)")
    rowptr, col, D1, D2 = _i32(rowptr), _i32(col), _f32(D1), _f32(D2)
    M, K = rowptr.size - 1, D1.shape[1]
    out = np.zeros(col.size, np.float32)
    r.ref_sddmm_host(ctypes.c_int(M), ctypes.c_int(D2.shape[0]), ctypes.c_int(K),
                     ctypes.c_int(col.size), _p(rowptr, _i32p), _p(col, _i32p), _p(D1, _f32p),
                     _p(D2, _f32p), _p(out, _f32p))
    return out
