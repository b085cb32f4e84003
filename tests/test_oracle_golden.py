"""CPU: pins the oracle (oracle/oracle.c) against
  * the golden vectors produced by the UNMODIFIED reference host functions (tests/golden/*, made by
    tests/golden/make_fixtures.py from oracle/_ref/libref_host.so = example/util/sp_util.hpp:62-112),
  * scipy tocsc() as test/test_csr2csr.py:42-49 does,
  * torch.sparse.mm(csr, X, reduce) on CPU, the oracle of the reference's own tests
    (test/test_spmm.py:60-61,97-98,134-135),
  * the compiled reference itself (oracle/_ref) when it is present.
"""
import os

import numpy as np
import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ["p2p-Gnutella31", "ca-CondMat"]


@pytest.mark.parametrize("name", NAMES)
def test_spmm_matches_reference_golden(oracle, graphs, name):
    rowptr, col, (M, K) = graphs.load_fixture(name)
    g = np.load(os.path.join(GOLDEN, name + "_spmm32.npz"))
    sv, sb = (int(x) for x in g["seeds"])
    val = graphs.uniform(col.size, sv)
    B = graphs.uniform(K * 32, sb).reshape(K, 32)
    out = oracle.spmm(rowptr, col, val, B)
    # the restatement performs the same fp32 operations in the same order: bit-exact
    assert np.array_equal(out[g["rows"]], g["out_rows"])
    assert np.allclose(out.astype(np.float64).sum(0), g["colsum"], rtol=1e-12, atol=0)


@pytest.mark.parametrize("name", NAMES)
def test_sddmm_matches_reference_golden(oracle, graphs, name):
    rowptr, col, (M, K) = graphs.load_fixture(name)
    g = np.load(os.path.join(GOLDEN, name + "_sddmm32.npz"))
    s1, s2 = (int(x) for x in g["seeds"])
    D1 = graphs.uniform(M * 32, s1).reshape(M, 32)
    D2 = graphs.uniform(K * 32, s2).reshape(K, 32)
    out = oracle.sddmm_csr(rowptr, col, D1, D2)
    assert np.array_equal(out, g["out"])
    # COO flavour on the expanded rows gives the same numbers
    row = np.repeat(np.arange(M, dtype=np.int32), np.diff(rowptr))
    assert np.array_equal(oracle.sddmm_coo(row, col, D1, D2), g["out"])


@pytest.mark.parametrize("name", NAMES)
def test_csr2csc_matches_scipy_golden(oracle, graphs, name):
    rowptr, col, (M, K) = graphs.load_fixture(name)
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    val = graphs.uniform(col.size, 5)
    colptr, row, val_t, perm = oracle.csr2csc(rowptr, col, val, ncols=K)
    assert np.array_equal(colptr, z["colptr"])
    assert np.array_equal(row, z["row"])
    assert np.array_equal(perm, z["perm"])
    assert np.array_equal(val_t, val[z["perm"]])


@pytest.mark.parametrize("reduce,tname", [("sum", "sum"), ("max", "amax"), ("min", "amin"), ("mean", "mean")])
def test_spmm_reduce_matches_torch_sparse_cpu(oracle, graphs, reduce, tname):
    rowptr, col, (M, K) = graphs.load_fixture("p2p-Gnutella31")   # 46k empty rows
    N = 16
    B = graphs.uniform(K * N, 3, -1.0, 1.0).reshape(K, N)
    val = np.ones(col.size, np.float32)                            # test/utils.py:52
    out, E = oracle.spmm(rowptr, col, val, B, reduce=reduce, with_arg=True) if reduce in ("max", "min") else \
        (oracle.spmm(rowptr, col, val, B, reduce=reduce), None)
    A = torch.sparse_csr_tensor(torch.from_numpy(rowptr.astype(np.int64)), torch.from_numpy(col.astype(np.int64)),
                                torch.from_numpy(val), size=(M, K))
    ref = torch.sparse.mm(A, torch.from_numpy(B), tname).numpy() if reduce != "sum" else \
        torch.sparse.mm(A, torch.from_numpy(B)).numpy()
    assert np.allclose(out, ref, rtol=1e-5, atol=1e-5)   # terms in [-1, 1]: order-of-summation noise near 0
    if E is not None:
        deg = np.diff(rowptr)
        assert (E[deg == 0] == -1).all() and (out[deg == 0] == 0).all()
        r = np.nonzero(deg > 0)[0][:2000]
        assert np.array_equal(B[E[r], np.arange(N)[None, :]], out[r])


def test_gspmm_compute_ops(oracle, graphs):
    rowptr, col = graphs.random_csr(300, 200, 4000, 7, empty_frac=0.2)
    val = graphs.uniform(col.size, 1, 0.5, 1.5)
    B = graphs.uniform(200 * 8, 2, -1, 1).reshape(200, 8)
    row = np.repeat(np.arange(300), np.diff(rowptr))
    for cop, f in [("add", lambda a, b: a + b), ("sub", lambda a, b: b - a), ("mul", lambda a, b: a * b),
                   ("div", lambda a, b: b / a)]:     # src/gspmm-fp/gspmm.h:53-79: a = edge, b = feat
        x = f(val[:, None].astype(np.float32), B[col])
        ref = np.zeros((300, 8), np.float64)
        np.add.at(ref, row, x.astype(np.float64))
        assert np.allclose(oracle.spmm(rowptr, col, val, B, "sum", cop), ref, rtol=1e-5, atol=1e-5)
        mx = np.full((300, 8), -np.inf)
        np.maximum.at(mx, row, x)
        mx[np.diff(rowptr) == 0] = 0
        assert np.array_equal(oracle.spmm(rowptr, col, val, B, "max", cop), mx.astype(np.float32))
    cp = oracle.spmm(rowptr, col, None, B, "mean")
    ref = np.zeros((300, 8), np.float64)
    np.add.at(ref, row, B[col].astype(np.float64))
    deg = np.maximum(np.diff(rowptr), 1)[:, None]
    assert np.allclose(cp, ref / deg, rtol=1e-5, atol=1e-6)


def test_masked_backward_oracles(oracle, graphs):
    """spmm_mask / sddmm_csr_mask against a dense numpy statement of the max backward."""
    M = K = 60
    rowptr, col = graphs.random_csr(M, K, 500, 3, empty_frac=0.1)
    val = graphs.uniform(col.size, 1, 0.5, 1.5)
    B = graphs.uniform(K * 4, 2).reshape(K, 4)
    G = graphs.uniform(M * 4, 4).reshape(M, 4)
    out, E = oracle.spmm(rowptr, col, val, B, "max", with_arg=True)
    gB = np.zeros((K, 4)); gv = np.zeros(col.size)
    for r in range(M):
        for p in range(rowptr[r], rowptr[r + 1]):
            for v in range(4):
                if E[r, v] == col[p]:
                    gB[col[p], v] += val[p] * G[r, v]
                    gv[p] += G[r, v] * B[col[p], v]
    colptr, row, val_t, perm = oracle.csr2csc(rowptr, col, val, ncols=K)
    assert np.allclose(oracle.spmm_mask(colptr, row, val_t, G, E), gB, rtol=1e-5, atol=1e-6)
    assert np.allclose(oracle.sddmm_csr_mask(rowptr, col, G, B, E), gv, rtol=1e-5, atol=1e-6)


def test_oracle_equals_compiled_reference(oracle, graphs):
    if oracle.ref_lib() is None:
        pytest.skip("oracle/_ref not built (reference tree absent on this machine)")
    rowptr, col = graphs.random_csr(2000, 1500, 60000, 11, empty_frac=0.3, hub=2)
    val = graphs.uniform(col.size, 1)
    B = graphs.uniform(1500 * 24, 2).reshape(1500, 24)
    assert np.array_equal(oracle.spmm(rowptr, col, val, B), oracle.ref_spmm_host(rowptr, col, val, B))
    D1 = graphs.uniform(2000 * 24, 3).reshape(2000, 24)
    assert np.array_equal(oracle.sddmm_csr(rowptr, col, D1, B), oracle.ref_sddmm_host(rowptr, col, D1, B))
    out_t = oracle.ref_spmm_host_threads(rowptr, col, val, B, threads=3)
    assert np.array_equal(out_t, oracle.ref_spmm_host(rowptr, col, val, B))


def test_spconv_oracle_on_fixture(oracle):
    z = np.load(os.path.join(GOLDEN, "spconv_fp32_0.npz"))
    kpos, imap, omap = z["kpos"], z["imap"], z["omap"]
    in_nnz, out_nnz = int(z["in_nnz"]), int(z["out_nnz"])
    rng = np.random.default_rng(0)
    c_in, c_out = 4, 8
    x = rng.standard_normal((in_nnz, c_in)).astype(np.float32)
    W = rng.standard_normal((kpos.size - 1, c_in, c_out)).astype(np.float32)
    out = oracle.spconv(kpos, imap, omap, x, W, out_nnz)
    ref = np.zeros((out_nnz, c_out), np.float64)
    for k in range(kpos.size - 1):
        s, e = kpos[k], kpos[k + 1]
        np.add.at(ref, omap[s:e], x[imap[s:e]].astype(np.float64) @ W[k].astype(np.float64))
    assert np.allclose(out, ref, rtol=1e-4, atol=1e-4)


def test_spconv_oracle_pinned_to_reference_cpu_compute(oracle):
    """oracle_spconv against golden OUTPUTS of the reference's own cpu_compute (test/test_spconv.py:17-53), generated by
    tests/golden/make_fixtures.py (which imports and runs that function unmodified): bit for bit, both call forms."""
    g = np.load(os.path.join(GOLDEN, "spconv_cpu_compute.npz"))
    out_size = int(g["out_size"])
    for knnz, imap, omap, pre, want in ((g["knnz"], g["imap"], g["omap"], False, g["out_full"]),
                                        (g["knnz_mid"], g["imap_mid"], g["omap_mid"], True, g["out_separate_mid"])):
        kpos = np.concatenate([[0], np.cumsum(knnz)]).astype(np.int32)
        got = oracle.spconv(kpos, imap, omap, g["feats"], g["W"], out_size, precompute=pre)
        assert got.dtype == np.float32 and np.array_equal(got, want), float(np.abs(got - want).max())
