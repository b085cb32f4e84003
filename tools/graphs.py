"""Synthetic CSR workloads of SURVEY.md §8(d) and loaders for the committed fixtures.

Input manufacture only — neither product nor oracle.  Datasets (reddit, ogbn-products, ogbn-arxiv)
cannot be downloaded here, so each workload is a seeded synthetic graph with the real dataset's
M, nnz and degree skew.  Columns are drawn uniformly without replacement per row by tools/graphgen.c.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libgraphgen.so")
_GOLDEN = os.path.join(os.path.dirname(_HERE), "tests", "golden")


def build(force=False):
    src = os.path.join(_HERE, "graphgen.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["/usr/bin/gcc", "-O3", "-fopenmp", "-fPIC", "-shared", "-o", _SO, src])


_lib = None


def _l():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def uniform(n, seed, lo=0.0, hi=1.0):
    x = np.empty(int(n), np.float32)
    _l().gen_uniform(ctypes.c_int64(int(n)), ctypes.c_uint64(seed), ctypes.c_float(lo),
                     ctypes.c_float(hi), x.ctypes.data_as(ctypes.c_void_p))
    return x


def _fit_degrees(raw, target_nnz, dmin, dmax, rng):
    """Rescale positive weights to integer degrees in [dmin, dmax] summing exactly to target_nnz."""
    raw = np.asarray(raw, np.float64)
    lo, hi = 0.0, float(target_nnz) / raw.sum() * 64
    for _ in range(80):
        s = 0.5 * (lo + hi)
        tot = np.clip(np.rint(raw * s), dmin, dmax).sum()
        if tot < target_nnz:
            lo = s
        else:
            hi = s
    d = np.clip(np.rint(raw * hi), dmin, dmax).astype(np.int64)
    diff = int(target_nnz - d.sum())
    order = rng.permutation(d.size)
    i = 0
    while diff != 0:  # spread the rounding remainder one unit at a time
        r = order[i % d.size]
        step = 1 if diff > 0 else -1
        if dmin <= d[r] + step <= dmax:
            d[r] += step
            diff -= step
        i += 1
    return d


def from_degrees(deg, K, seed):
    deg = np.asarray(deg, np.int64)
    M = deg.size
    rowptr = np.zeros(M + 1, np.int64)
    np.cumsum(deg, out=rowptr[1:])
    assert rowptr[-1] < 2**31
    col = np.empty(int(rowptr[-1]), np.int32)
    _l().gen_columns(ctypes.c_int(M), ctypes.c_int(K), rowptr.ctypes.data_as(ctypes.c_void_p),
                     ctypes.c_uint64(seed), col.ctypes.data_as(ctypes.c_void_p))
    return rowptr.astype(np.int32), col


def from_degrees_community(deg, K, seed, comm_size, p_intra):
    """from_degrees with planted communities: rows [c*comm_size, (c+1)*comm_size) draw a column from their own community
    with probability p_intra, else uniformly (tools/graphgen.c gen_columns_community)."""
    deg = np.asarray(deg, np.int64)
    M = deg.size
    rowptr = np.zeros(M + 1, np.int64)
    np.cumsum(deg, out=rowptr[1:])
    assert rowptr[-1] < 2**31
    col = np.empty(int(rowptr[-1]), np.int32)
    _l().gen_columns_community(ctypes.c_int(M), ctypes.c_int(K), rowptr.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(seed),
                               ctypes.c_int(int(comm_size)), ctypes.c_int(int(round(p_intra * 1000))),
                               col.ctypes.data_as(ctypes.c_void_p))
    return rowptr.astype(np.int32), col


def reddit_like_communities(scale=1.0, comm_size=1024, p_intra=0.85):
    """The reddit-like workload (same M, nnz, degree law and seed as reddit_like) with LOCALITY: 85 % of a row's columns
    fall in the row's own community of `comm_size` consecutive nodes — the structure real social / co-purchase graphs have
    after a community-aware node ordering, which the uniform generator (the worst case for every cache) lacks."""
    M = max(64, int(round(232965 * scale)))
    nnz = int(round(114615892 * scale))
    rng = np.random.Generator(np.random.PCG64(20240001))
    raw = rng.lognormal(5.4, 1.3, M)
    deg = _fit_degrees(raw, nnz, 1, min(21657, M), rng)
    return from_degrees_community(deg, M, 20240001, comm_size, p_intra)


def reddit_like(scale=1.0):
    """M=K=232 965, nnz=114 615 892, lognormal(5.4, 1.3) degrees clipped to [1, 21 657] (§8d config 2).
    scale < 1 shrinks M and nnz together (same mean degree) for parity-sized cases."""
    M = max(64, int(round(232965 * scale)))
    nnz = int(round(114615892 * scale))
    rng = np.random.Generator(np.random.PCG64(20240001))
    raw = rng.lognormal(5.4, 1.3, M)
    deg = _fit_degrees(raw, nnz, 1, min(21657, M), rng)
    return from_degrees(deg, M, 20240001)


def products_like(scale=1.0):
    """M=K=2 449 029, nnz=123 718 280, power-law (alpha 2.1) degrees clipped to [1, 17 481] (config 3)."""
    M = max(64, int(round(2449029 * scale)))
    nnz = int(round(123718280 * scale))
    rng = np.random.Generator(np.random.PCG64(20240002))
    raw = (1.0 - rng.random(M)) ** (-1.0 / 1.1)  # Pareto tail, pdf ~ x^-2.1
    deg = _fit_degrees(raw, nnz, 1, min(17481, M), rng)
    return from_degrees(deg, M, 20240002)


def arxiv_like(scale=1.0):
    """M=K=169 343, nnz=1 166 243 (directed, mean degree 6.9, max ~13k) (config 4)."""
    M = max(64, int(round(169343 * scale)))
    nnz = int(round(1166243 * scale))
    rng = np.random.Generator(np.random.PCG64(20240003))
    raw = (1.0 - rng.random(M)) ** (-1.0 / 1.3)
    deg = _fit_degrees(raw, nnz, 0, min(13155, M), rng)
    return from_degrees(deg, M, 20240003)


def random_csr(M, K, nnz, seed, empty_frac=0.0, hub=0):
    """Small test matrix: ragged rows, an optional share of empty rows, optional hub rows."""
    rng = np.random.Generator(np.random.PCG64(seed))
    w = rng.random(M) ** 3 + 1e-3
    if empty_frac > 0:
        w[rng.random(M) < empty_frac] = 0
    for h in range(hub):
        w[rng.integers(0, M)] = w.sum() * 0.2
    if w.sum() == 0:
        w[0] = 1
    deg = np.floor(w / w.sum() * nnz).astype(np.int64)
    deg = np.minimum(deg, K)
    return from_degrees(deg, K, seed)


def load_fixture(name):
    """tests/golden/<name>.npz -> (rowptr int32, col int32, shape).  Made by tests/golden/make_fixtures.py
    from /root/reference/example/data/<name>.mtx via scipy mmread().tocsr(), as test/test_csr2csr.py:20-28."""
    z = np.load(os.path.join(_GOLDEN, name + ".npz"))
    return z["rowptr"].astype(np.int32), z["col"].astype(np.int32), tuple(int(x) for x in z["shape"])
