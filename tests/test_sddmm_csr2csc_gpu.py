"""GPU parity: SDDMM (CSR / COO / mean / masked) and the exact CSR->CSC transpose vs the oracle."""
import os

import numpy as np
import pytest
import torch

from util import assert_close_f32, sddmm_absref, spmm_absref

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def K():
    import dgsparse._kernels as k
    return k


@pytest.mark.parametrize("name", ["p2p-Gnutella31", "ca-CondMat"])
def test_sddmm_golden_k32(K, graphs, name):
    rowptr, col, (M, Kc) = graphs.load_fixture(name)
    g = np.load(os.path.join(GOLDEN, name + "_sddmm32.npz"))
    s1, s2 = (int(x) for x in g["seeds"])
    D1 = graphs.uniform(M * 32, s1).reshape(M, 32)
    D2 = graphs.uniform(Kc * 32, s2).reshape(Kc, 32)
    out = K.sddmm_csr(dev(rowptr), dev(col), dev(D1), dev(D2))
    assert tuple(out.shape) == (1, col.size)            # src/cuda/spmm_cuda.cu:342
    assert_close_f32(out.cpu().numpy()[0], g["out"], what=name)
    row = np.repeat(np.arange(M, dtype=np.int32), np.diff(rowptr))
    out2 = K.sddmm_coo(dev(row), dev(col), dev(D1), dev(D2))
    assert tuple(out2.shape) == (col.size,)
    assert_close_f32(out2.cpu().numpy(), g["out"], what=name + " coo")
    # legacy C symbols
    import dgsparse._lib as L
    d = [dev(rowptr), dev(col), dev(D1), dev(D2), dev(row)]
    o = torch.empty(col.size, device="cuda")
    torch.cuda.synchronize()
    L.lib.sddmm_cuda_csr(M, 32, col.size, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), o.data_ptr())
    torch.cuda.synchronize()
    assert_close_f32(o.cpu().numpy(), g["out"], what="sddmm_cuda_csr")
    o.zero_()
    L.lib.sddmm_cuda_coo(32, col.size, d[4].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), o.data_ptr())
    torch.cuda.synchronize()
    assert_close_f32(o.cpu().numpy(), g["out"], what="sddmm_cuda_coo")


@pytest.mark.parametrize("Kd", [1, 2, 7, 16, 36, 64, 100, 128, 132, 200, 256, 300, 512, 640, 1024, 1028])
def test_sddmm_widths(K, oracle, graphs, Kd):
    """Any K, including K%4==0 but K%32!=0 where the reference's vec4 CSR kernel drops the residue (q13)."""
    M, Kc = 2000, 1500
    rowptr, col = graphs.random_csr(M, Kc, 40000, 50 + Kd, empty_frac=0.3, hub=2)
    D1 = graphs.uniform(M * Kd, 1, -1, 1).reshape(M, Kd)
    D2 = graphs.uniform(Kc * Kd, 2, -1, 1).reshape(Kc, Kd)
    for mean in (False, True):
        out = K.sddmm_csr(dev(rowptr), dev(col), dev(D1), dev(D2), mean=mean).cpu().numpy()[0]
        assert_close_f32(out, oracle.sddmm_csr(rowptr, col, D1, D2, mean), oracle.sddmm_csr(rowptr, col, D1, D2, mean, f64=True),
                         what=f"K={Kd} mean={mean}", absref=sddmm_absref(oracle, rowptr, col, D1, D2, mean))
    row = np.repeat(np.arange(M, dtype=np.int32), np.diff(rowptr))
    perm = np.random.default_rng(1).permutation(col.size)       # COO needs no ordering
    out = K.sddmm_coo(dev(row[perm]), dev(col[perm]), dev(D1), dev(D2)).cpu().numpy()
    assert_close_f32(out, oracle.sddmm_coo(row[perm], col[perm], D1, D2), what=f"coo K={Kd}",
                     absref=sddmm_absref(oracle, rowptr, col, D1, D2)[perm])


@pytest.mark.parametrize("Kd", [64, 100, 128, 256, 384, 512])
@pytest.mark.parametrize("shape", ["deg1", "deg1to3", "mixed"])
def test_sddmm_ring_two_d1_slots(K, oracle, graphs, Kd, shape):
    """The ring kernel with TWO D1 slots per stage (what CSR inputs with >= 4 edges per row get): a batch's third / fourth distinct
    row is read from global by the consumer.  Matrices made of very short rows take that path on most batches; the result must be
    bit-identical to the four-slot ring (same dot, only the source of the D1 row differs) and match the oracle."""
    import ctypes

    import dgsparse._lib as L
    rng = np.random.default_rng(7)
    M, Kc = 6000, 3000
    if shape == "deg1":
        deg = np.ones(M, np.int64)
    elif shape == "deg1to3":
        deg = rng.integers(0, 4, M)
    else:   # runs of one-edge rows between hub rows and empty stretches
        deg = rng.integers(0, 3, M)
        deg[::97] = rng.integers(40, 300, deg[::97].size)
        deg[1000:1200] = 0
    rowptr = np.zeros(M + 1, np.int32)
    rowptr[1:] = np.cumsum(deg)
    col = rng.integers(0, Kc, int(rowptr[-1])).astype(np.int32)
    D1 = graphs.uniform(M * Kd, 1, -1, 1).reshape(M, Kd)
    D2 = graphs.uniform(Kc * Kd, 2, -1, 1).reshape(Kc, Kd)
    got = {}
    try:
        for nd in (2, 4):
            assert L.lib.dgs_set_option(b"sddmm_d1slots", nd) == 0
            assert L.lib.dgs_set_option(b"sddmm_no_ring", 0) == 0     # the ring also for K = 64 on a small input
            for mean in (False, True):
                got[nd, mean] = K.sddmm_csr(dev(rowptr), dev(col), dev(D1), dev(D2), mean=mean).cpu().numpy()[0]
                w = ctypes.c_int()
                L.lib.dgs_sddmm_last_geometry(ctypes.byref(w), None, None)
                assert w.value >= 1, "ring kernel expected"
    finally:
        L.lib.dgs_set_option(b"sddmm_d1slots", -1)
        L.lib.dgs_set_option(b"sddmm_no_ring", -1)
    for mean in (False, True):
        assert np.array_equal(got[2, mean], got[4, mean]), f"two vs four D1 slots differ (mean={mean})"
        assert_close_f32(got[2, mean], oracle.sddmm_csr(rowptr, col, D1, D2, mean), oracle.sddmm_csr(rowptr, col, D1, D2, mean, f64=True),
                         what=f"two D1 slots K={Kd} {shape} mean={mean}", absref=sddmm_absref(oracle, rowptr, col, D1, D2, mean))


@pytest.mark.parametrize("Kd", [8, 30, 64, 256])
@pytest.mark.parametrize("nnz_shape", ["1", "3", "7", "9", "hub0_tail", "hub0_33"])
def test_sddmm_tiny_and_hub_row0(K, oracle, graphs, Kd, nnz_shape):
    """the last (ragged) edge group of a matrix whose row 0 reaches it — rows [0,1,2,pad..] — must not be
    multiplied with D1[row 0] throughout.  nnz in {1,3,7,9} and a hub row 0 followed by short rows, every kernel family
    (register-staged K<64 / unaligned K, ring K>=64), plain / MEAN / masked."""
    M, Kc = 12, 17
    if nnz_shape.isdigit():
        rows = (np.arange(int(nnz_shape)) % M).astype(np.int32)     # rows 0, 1, 2, ...: row 0 first, other rows after it
    elif nnz_shape == "hub0_tail":
        rows = np.array([0] * 29 + [1, 2, 5], np.int32)               # last 8-group starts in row 0 and ends in row 5
    else:
        rows = np.array([0] * 33 + [3, 4], np.int32)
    col = (np.arange(rows.size, dtype=np.int32) * 5 + 3) % Kc
    rowptr = np.zeros(M + 1, np.int32)
    np.add.at(rowptr, rows + 1, 1)
    rowptr = np.cumsum(rowptr).astype(np.int32)
    order = np.lexsort((col, rows))
    col = col[order].astype(np.int32)
    D1 = graphs.uniform(M * Kd, 31, -1, 1).reshape(M, Kd)
    D2 = graphs.uniform(Kc * Kd, 32, -1, 1).reshape(Kc, Kd)
    for mean in (False, True):
        out = K.sddmm_csr(dev(rowptr), dev(col), dev(D1), dev(D2), mean=mean).cpu().numpy()[0]
        assert_close_f32(out, oracle.sddmm_csr(rowptr, col, D1, D2, mean), oracle.sddmm_csr(rowptr, col, D1, D2, mean, f64=True),
                         what=f"tiny {nnz_shape} K={Kd} mean={mean}", absref=sddmm_absref(oracle, rowptr, col, D1, D2, mean))
    E = ((np.arange(M * Kd, dtype=np.int64) * 7) % Kc).astype(np.int32).reshape(M, Kd)
    got = K.sddmm_csr(dev(rowptr), dev(col), dev(D1), dev(D2), E=dev(E)).cpu().numpy()[0]
    assert_close_f32(got, oracle.sddmm_csr_mask(rowptr, col, D1, D2, E), what=f"tiny masked {nnz_shape} K={Kd}",
                     absref=sddmm_absref(oracle, rowptr, col, D1, D2))


def test_sddmm_arxiv_like_k256(K, oracle, graphs):
    """BASELINE config 4 at full size: arxiv-like CSR, K=256."""
    rowptr, col = graphs.arxiv_like(1.0)
    M = rowptr.size - 1
    D1 = graphs.uniform(M * 256, 1).reshape(M, 256)
    D2 = graphs.uniform(M * 256, 2).reshape(M, 256)
    out = K.sddmm_csr(dev(rowptr), dev(col), dev(D1), dev(D2)).cpu().numpy()[0]
    ref64 = oracle.sddmm_csr(rowptr, col, D1, D2, f64=True)
    assert_close_f32(out, oracle.sddmm_csr(rowptr, col, D1, D2), ref64, what="arxiv", absref=ref64)   # inputs >= 0


def test_masked_kernels(K, oracle, graphs):
    M = Kc = 1500
    N = 64
    rowptr, col = graphs.random_csr(M, Kc, 30000, 21, empty_frac=0.2, hub=1)
    val = graphs.uniform(col.size, 1, 0.5, 1.5)
    B = graphs.uniform(Kc * N, 2, -1, 1).reshape(Kc, N)
    G = graphs.uniform(M * N, 3, -1, 1).reshape(M, N)
    _, E = oracle.spmm(rowptr, col, val, B, "max", with_arg=True)
    colptr, row, val_t, perm = oracle.csr2csc(rowptr, col, val, ncols=Kc)
    got = K.spmm_with_mask(dev(colptr), dev(row), dev(val_t), dev(G), dev(E)).cpu().numpy()
    assert_close_f32(got, oracle.spmm_mask(colptr, row, val_t, G, E), what="spmm_mask",
                     absref=spmm_absref(oracle, colptr, row, val_t, G))       # the unmasked sum of |terms| bounds the masked one
    got = K.sddmm_csr(dev(rowptr), dev(col), dev(G), dev(B), E=dev(E)).cpu().numpy()[0]
    assert_close_f32(got, oracle.sddmm_csr_mask(rowptr, col, G, B, E), what="sddmm_mask",
                     absref=sddmm_absref(oracle, rowptr, col, G, B))


@pytest.mark.parametrize("name", ["p2p-Gnutella31", "ca-CondMat"])
def test_csr2csc_bit_exact_vs_scipy_golden(K, graphs, name):
    """test/test_csr2csr.py:42-49: (colptr, row, values) equal scipy tocsc(); plus the exact permutation."""
    rowptr, col, (M, Kc) = graphs.load_fixture(name)
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    val = graphs.uniform(col.size, 5)
    colptr, row, val_t, perm = K.csr2csc(dev(rowptr), dev(col), dev(val), ncols=Kc)
    assert np.array_equal(colptr.cpu().numpy(), z["colptr"])
    assert np.array_equal(row.cpu().numpy(), z["row"])
    assert np.array_equal(perm.cpu().numpy(), z["perm"])
    assert np.array_equal(val_t.cpu().numpy(), val[z["perm"]])
    # the registered op, reference signature (square matrix)
    c2, r2, v2 = torch.ops.dgsparse_spmm.csr2csc(dev(rowptr), dev(col), dev(val))
    assert np.array_equal(c2.cpu().numpy(), z["colptr"]) and np.array_equal(r2.cpu().numpy(), z["row"])
    assert np.array_equal(v2.cpu().numpy(), val[z["perm"]])


# (20, 3000000, 50000): 22 key bits = three radix passes (first / middle / last kernels); 70000 columns = two; 50 = one
@pytest.mark.parametrize("shape", [(1, 1, 1), (50, 70000, 4000), (70000, 50, 300000), (5000, 5000, 0), (3000, 300, 200000),
                                   (20, 3000000, 50000), (2000, 600000, 150000)])
def test_csr2csc_shapes(K, oracle, graphs, shape):
    M, Kc, nnz = shape
    if nnz == 0:
        rowptr, col = np.zeros(M + 1, np.int32), np.zeros(0, np.int32)
    else:
        rowptr, col = graphs.random_csr(M, Kc, nnz, 77, empty_frac=0.1, hub=1)
    ref = oracle.csr2csc(rowptr, col, None, ncols=Kc)
    colptr, row, _, perm = K.csr2csc(dev(rowptr), dev(col), None, ncols=Kc)
    assert np.array_equal(colptr.cpu().numpy(), ref[0])
    assert np.array_equal(row.cpu().numpy(), ref[1])
    assert np.array_equal(perm.cpu().numpy(), ref[3])


@pytest.mark.parametrize("graph", ["reddit_like", "products_like"])
def test_csr2csc_full_size_above_2p24(K, graphs, graph):
    """reddit-like (114.6 M nnz, two radix passes) and products-like (123.7 M nnz, 2.45 M columns: three passes, a suffix-minimum
    over more than one block of tiles): > 2^24 nnz, where the reference's float32-arange permutation breaks (q10).
    Properties: perm is a permutation; col[perm] is sorted; ties keep CSR order; colptr = bincount."""
    rowptr, col = getattr(graphs, graph)(1.0)
    M, nnz = rowptr.size - 1, col.size
    rp, cc = dev(rowptr), dev(col)
    colptr, row, _, perm = K.csr2csc(rp, cc, None, ncols=M)
    key = cc.long()[perm.long()] * (1 << 31) + perm.long()
    assert bool((key[1:] > key[:-1]).all())                       # sorted by (col, CSR position) => stable
    counts = torch.bincount(cc.long(), minlength=M)
    assert torch.equal(colptr[1:].long(), torch.cumsum(counts, 0)) and int(colptr[0]) == 0
    chk = torch.zeros(nnz, dtype=torch.int32, device="cuda")
    chk[perm.long()] = 1
    assert int(chk.sum()) == nnz
    rows_ref = torch.repeat_interleave(torch.arange(M, device="cuda", dtype=torch.int32), (rp[1:] - rp[:-1]).long())
    assert torch.equal(row, rows_ref[perm.long()])


def test_edge_softmax(graphs):
    import dgsparse._kernels as K2
    rowptr, col = graphs.random_csr(500, 500, 8000, 3, empty_frac=0.2)
    v = graphs.uniform(col.size * 2, 1, -2, 2).reshape(-1, 2)
    out = K2.edge_softmax(dev(rowptr), dev(v), head=2).cpu().numpy()
    for r in range(0, 500, 37):
        s, e = rowptr[r], rowptr[r + 1]
        if e > s:
            x = np.exp(v[s:e] - v[s:e].max(0))
            assert np.allclose(out[s:e], x / x.sum(0), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("Kd", [16, 64, 256])
def test_edge_sharded_sddmm_single_rank_matches_csr(K, graphs, Kd):
    """dgsparse.distributed.EdgeShardedSDDMM at world 1 (the 2-rank run is tests/test_multigpu_gpu.py): the COO
    slice kernel must reproduce the CSR kernel bit for bit, empty rows and hub rows included."""
    from dgsparse.distributed import EdgeShardedSDDMM
    M = 3000
    rowptr, col = graphs.random_csr(M, M, 70000, 9, empty_frac=0.3, hub=2)
    D1 = dev(graphs.uniform(M * Kd, 3, -1, 1).reshape(M, Kd))
    D2 = dev(graphs.uniform(M * Kd, 4, -1, 1).reshape(M, Kd))
    rp, cc = dev(rowptr), dev(col)
    op = EdgeShardedSDDMM(rp, cc)
    got = op(D1, D2)
    assert tuple(got.shape) == (1, col.size)
    assert torch.equal(got, K.sddmm_csr(rp, cc, D1, D2))
