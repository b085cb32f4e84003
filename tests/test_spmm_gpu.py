"""GPU parity: CUDA SpMM / generalized SpMM (through the C ABI) vs the CPU oracle."""
import ctypes
import os

import numpy as np
import pytest
import torch

from util import assert_close_f32, spmm_absref

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RED = {"sum": 0, "max": 1, "min": 2, "mean": 3}
COMP = {"add": 0, "sub": 1, "mul": 2, "div": 3}


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def K():
    import dgsparse._kernels as k
    return k


@pytest.fixture
def knob():
    """dgs_set_option(name, value) for the duration of one test (include/dgsparse_b200.h), cleared afterwards."""
    import dgsparse._lib as L
    touched = []

    def set_(name, value):
        assert L.lib.dgs_set_option(name.encode(), int(value)) == 0, name
        touched.append(name)
    yield set_
    for name in touched:
        L.lib.dgs_set_option(name.encode(), -1)


@pytest.mark.parametrize("name", ["p2p-Gnutella31", "ca-CondMat"])
def test_config1_golden_feat32(K, oracle, graphs, name):
    """BASELINE config 1: example/data CSR, feat=32, against the reference host function's own output."""
    rowptr, col, (M, Kc) = graphs.load_fixture(name)
    g = np.load(os.path.join(GOLDEN, name + "_spmm32.npz"))
    sv, sb = (int(x) for x in g["seeds"])
    val = graphs.uniform(col.size, sv)
    B = graphs.uniform(Kc * 32, sb).reshape(Kc, 32)
    out = K.spmm(dev(rowptr), dev(col), dev(val), dev(B)).cpu().numpy()
    assert_close_f32(out[g["rows"]], g["out_rows"], what=f"{name} golden rows")
    assert np.allclose(out.astype(np.float64).sum(0), g["colsum"], rtol=1e-6)
    # legacy symbol spmm_cuda(m, k, rowptr, colind, values, dense, out) on stream 0
    import dgsparse._lib as L
    d = [dev(rowptr), dev(col), dev(val), dev(B)]
    o2 = torch.empty(M, 32, device="cuda")
    torch.cuda.synchronize()
    L.lib.spmm_cuda(M, 32, *[t.data_ptr() for t in d], o2.data_ptr())
    torch.cuda.synchronize()
    # same arithmetic; bit-identical only when both calls took the same kernel family (the row-parallel kernel sums a row
    # in one piece, the segment kernel folds rows cut by a segment boundary), which depends on what the library has seen of
    # these device pointers before
    assert_close_f32(o2.cpu().numpy()[g["rows"]], g["out_rows"], what=f"{name} spmm_cuda golden rows")
    assert np.allclose(o2.cpu().numpy(), out, rtol=1e-5, atol=1e-6)
    o3 = torch.empty(M, 32, device="cuda")
    L.lib.spmm_cuda_no_edge_value(M, 32, d[0].data_ptr(), d[1].data_ptr(), None, d[3].data_ptr(), o3.data_ptr())
    torch.cuda.synchronize()
    assert_close_f32(o3.cpu().numpy(), oracle.spmm(rowptr, col, None, B), what="no_edge_value")


@pytest.mark.parametrize("N", [1, 3, 4, 17, 32, 33, 64, 100, 128, 200, 256, 512])
@pytest.mark.parametrize("reduce", ["sum", "max", "min", "mean"])
def test_widths_and_reduces(K, oracle, graphs, N, reduce):
    M, Kc = 3000, 2500
    rowptr, col = graphs.random_csr(M, Kc, 90000, 100 + N, empty_frac=0.25, hub=2)
    val = graphs.uniform(col.size, 1, 0.5, 1.5)
    B = graphs.uniform(Kc * N, 2, -1.0, 1.0).reshape(Kc, N)
    with_arg = reduce in ("max", "min")
    got = K.spmm(dev(rowptr), dev(col), dev(val), dev(B), RED[reduce], COMP["mul"], with_arg=with_arg)
    if with_arg:
        ref, Eref = oracle.spmm(rowptr, col, val, B, reduce, "mul", with_arg=True)
        out, E = got[0].cpu().numpy(), got[1].cpu().numpy()
        assert np.array_equal(out, ref)         # max / min are exact
        assert np.array_equal(E, Eref)          # first extremum wins, -1 on empty rows
    else:
        ref = oracle.spmm(rowptr, col, val, B, reduce, "mul")
        ref64 = oracle.spmm_f64(rowptr, col, val, B, reduce, "mul")
        assert_close_f32(got.cpu().numpy(), ref, ref64, what=f"N={N} {reduce}",
                         absref=spmm_absref(oracle, rowptr, col, val, B, reduce))


@pytest.mark.parametrize("compute", ["add", "sub", "mul", "div"])
@pytest.mark.parametrize("reduce", ["sum", "max", "min", "mean"])
def test_gspmm_all_ops(oracle, graphs, compute, reduce):
    """gspmm-fp surface: u_<op>_e_<reduce> and copy_u_<reduce> (example/gspmm-fp/util.py:17-110)."""
    import dgsparse.gspmm as G
    rowptr, col, (M, Kc) = graphs.load_fixture("p2p-Gnutella31")
    N = 64
    val = graphs.uniform(col.size, 3, 0.5, 1.5)
    B = graphs.uniform(Kc * N, 4, -1.0, 1.0).reshape(Kc, N)
    fn = getattr(G, f"u_{compute}_e_{reduce}")
    out = fn(dev(rowptr), dev(col), dev(val).reshape(-1, 1), dev(B)).cpu().numpy()
    ref = oracle.spmm(rowptr, col, val, B, reduce, compute)
    if reduce in ("max", "min"):
        assert np.array_equal(out, ref)
    else:
        assert_close_f32(out, ref, oracle.spmm_f64(rowptr, col, val, B, reduce, compute), what=fn.__name__,
                         absref=spmm_absref(oracle, rowptr, col, val, B, reduce, compute))
    if compute == "add":
        cp = getattr(G, f"copy_u_{reduce}")(dev(rowptr), dev(col), dev(B)).cpu().numpy()
        refc = oracle.spmm(rowptr, col, None, B, reduce)
        if reduce in ("max", "min"):
            assert np.array_equal(cp, refc)
        else:
            assert_close_f32(cp, refc, oracle.spmm_f64(rowptr, col, None, B, reduce), what="copy_u",
                             absref=spmm_absref(oracle, rowptr, col, None, B, reduce))


def test_edge_cases(K, oracle, graphs):
    # all rows empty
    rowptr = np.zeros(11, np.int32)
    col = np.zeros(0, np.int32)
    B = graphs.uniform(5 * 8, 1).reshape(5, 8)
    out, E = K.spmm(dev(rowptr), dev(col), None, dev(B), RED["max"], with_arg=True)
    assert (out.cpu().numpy() == 0).all() and (E.cpu().numpy() == -1).all()
    # one giant row (split over many segments) + trailing empty rows, every reduce
    Kc, N = 70000, 64
    cols = np.sort(np.random.default_rng(0).choice(Kc, 60000, replace=False)).astype(np.int32)
    rowptr = np.array([0, 0, 60000, 60000, 60000], np.int32)
    val = graphs.uniform(cols.size, 2, 0.5, 1.5)
    B = graphs.uniform(Kc * N, 3, -1, 1).reshape(Kc, N)
    for reduce in ["sum", "mean", "max", "min"]:
        wa = reduce in ("max", "min")
        got = K.spmm(dev(rowptr), dev(cols), dev(val), dev(B), RED[reduce], COMP["mul"], with_arg=wa)
        if wa:
            ref, Eref = oracle.spmm(rowptr, cols, val, B, reduce, with_arg=True)
            assert np.array_equal(got[0].cpu().numpy(), ref) and np.array_equal(got[1].cpu().numpy(), Eref)
        else:
            assert_close_f32(got.cpu().numpy(), oracle.spmm(rowptr, cols, val, B, reduce),
                             oracle.spmm_f64(rowptr, cols, val, B, reduce), what="giant row " + reduce,
                             absref=spmm_absref(oracle, rowptr, cols, val, B, reduce))
    # ties: all-equal features -> arg must be the FIRST nonzero's column (strict compare, spmm_cuda.cuh:38-42)
    Bc = np.ones((Kc, 8), np.float32)
    out, E = K.spmm(dev(rowptr), dev(cols), None, dev(Bc), RED["max"], with_arg=True)
    assert (E.cpu().numpy()[1] == cols[0]).all()
    # strided dense operand (column panel of a wider matrix): ldb != N
    wide = graphs.uniform(Kc * 96, 5).reshape(Kc, 96)
    panel = dev(wide)[:, 32:64]
    assert not panel.is_contiguous()
    import dgsparse._lib as L
    outp = torch.empty(4, 32, device="cuda")
    ws = torch.empty(L.lib.dgs_spmm_workspace_bytes(32, cols.size, 0), dtype=torch.uint8, device="cuda")
    rp, cc, vv = dev(rowptr), dev(cols), dev(val)
    L.check(L.lib.dgs_spmm_csr(4, 32, cols.size, rp.data_ptr(), cc.data_ptr(), vv.data_ptr(), panel.data_ptr(), 96,
                               outp.data_ptr(), 32, None, 0, 0, 2, ws.data_ptr(), ws.numel(), None), "strided")
    torch.cuda.synchronize()
    assert_close_f32(outp.cpu().numpy(), oracle.spmm(rowptr, cols, val, wide[:, 32:64]),
                     oracle.spmm_f64(rowptr, cols, val, wide[:, 32:64]), what="strided panel",
                     absref=spmm_absref(oracle, rowptr, cols, val, wide[:, 32:64]))


def test_gespmm_descr_api_and_colmajor(oracle, graphs):
    """gespmmCsrSpMM(SpMatCsrDescr_t, B, N, C, transpose_BC, alg) — src/ge-spmm/gespmm.h:9-33."""
    import dgsparse._lib as L
    M, Kc, N = 500, 400, 48
    rowptr, col = graphs.random_csr(M, Kc, 6000, 5, empty_frac=0.1)
    val = graphs.uniform(col.size, 1)
    B = graphs.uniform(Kc * N, 2).reshape(Kc, N)
    ref = oracle.spmm(rowptr, col, val, B)
    rp, cc, vv = dev(rowptr), dev(col), dev(val)
    d = L.SpMatCsrDescr_t(M, Kc, col.size, rp.data_ptr(), cc.data_ptr(), vv.data_ptr())
    for alg in (0, 8, 10):
        C = torch.empty(M, N, device="cuda")
        L.lib.gespmmCsrSpMM(d, dev(B).data_ptr(), N, C.data_ptr(), True, alg)
        torch.cuda.synchronize()
        assert_close_f32(C.cpu().numpy(), ref, what=f"alg {alg}")
    Bt = dev(np.ascontiguousarray(B.T))          # column-major B = row-major B^T
    Ct = torch.empty(N, M, device="cuda")
    L.lib.gespmmCsrSpMM(d, Bt.data_ptr(), N, Ct.data_ptr(), False, 4)
    torch.cuda.synchronize()
    assert_close_f32(Ct.cpu().numpy().T, ref, what="column-major")


@pytest.mark.parametrize("N", [7, 33, 64])
def test_gespmm_colmajor_both_paths_agree(oracle, graphs, N):
    """transpose_BC = false (src/ge-spmm/csrspmm_non_transpose.cu): the thread-per-element kernel and the row-major kernel
    between two tiled transposes (what larger products take) give the same column-major C; ragged tile edges (M, K, N not
    multiples of 32), empty rows and a hub row included."""
    import dgsparse._lib as L
    M, Kc = 3001, 2777
    rowptr, col = graphs.random_csr(M, Kc, 90000, 11, empty_frac=0.15, hub=1)
    val = graphs.uniform(col.size, 1, 0.5, 1.5)
    B = graphs.uniform(Kc * N, 2, -1.0, 1.0).reshape(Kc, N)
    want, want64, T = oracle.spmm(rowptr, col, val, B), oracle.spmm_f64(rowptr, col, val, B), spmm_absref(oracle, rowptr, col, val, B)
    rp, cc, vv = dev(rowptr), dev(col), dev(val)
    d = L.SpMatCsrDescr_t(M, Kc, col.size, rp.data_ptr(), cc.data_ptr(), vv.data_ptr())
    Bt = dev(np.ascontiguousarray(B.T))
    got = {}
    try:
        for path in (0, 1):
            assert L.lib.dgs_set_option(b"spmm_colmajor", path) == 0
            Ct = torch.full((N, M), float("nan"), device="cuda")
            L.lib.gespmmCsrSpMM(d, Bt.data_ptr(), N, Ct.data_ptr(), False, 0)
            torch.cuda.synchronize()
            got[path] = Ct.cpu().numpy().T
            assert_close_f32(got[path], want, want64, what=f"column-major path {path} N={N}", absref=T)
    finally:
        L.lib.dgs_set_option(b"spmm_colmajor", -1)
    # B is left untouched by the transposing path
    assert np.array_equal(Bt.cpu().numpy(), np.ascontiguousarray(B.T))
    # the legacy scratch (segment partials + the transposed copies) can be handed back and is re-allocated on demand
    free0 = torch.cuda.mem_get_info()[0]
    assert L.lib.dgs_legacy_scratch_release() == 0
    assert torch.cuda.mem_get_info()[0] >= free0
    assert L.lib.dgs_legacy_scratch_release() == 0          # nothing held: still fine
    Ct = torch.full((N, M), float("nan"), device="cuda")
    L.lib.gespmmCsrSpMM(d, Bt.data_ptr(), N, Ct.data_ptr(), False, 0)
    torch.cuda.synchronize()
    assert np.array_equal(Ct.cpu().numpy().T, got[1])


def test_host_buffer_entry(oracle, graphs):
    """dgs_spmm_csr_host: the HOST-pointer entry bench.py's e2e leg times."""
    import dgsparse._lib as L
    M, Kc, N = 4000, 3000, 64
    rowptr, col = graphs.random_csr(M, Kc, 150000, 9, empty_frac=0.2, hub=1)
    val = graphs.uniform(col.size, 1)
    B = graphs.uniform(Kc * N, 2).reshape(Kc, N)
    C = np.empty((M, N), np.float32)
    E = np.empty((M, N), np.int32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    L.check(L.lib.dgs_spmm_csr_host(M, Kc, N, col.size, p(rowptr), p(col), p(val), p(B), p(C), None, 0, 2), "host sum")
    ref64 = oracle.spmm_f64(rowptr, col, val, B)
    assert_close_f32(C, oracle.spmm(rowptr, col, val, B), ref64, what="host sum", absref=ref64)   # inputs >= 0
    L.check(L.lib.dgs_spmm_csr_host(M, Kc, N, col.size, p(rowptr), p(col), p(val), p(B), p(C), p(E), 1, 2), "host max")
    ref, Eref = oracle.spmm(rowptr, col, val, B, "max", with_arg=True)
    assert np.array_equal(C, ref) and np.array_equal(E, Eref)


def test_host_buffer_entry_pipelined_row_blocks(oracle, graphs, monkeypatch):
    """The host entry cuts the CSR into row blocks of ~equal nnz and overlaps H2D / SpMM / D2H: force many blocks
    (skewed rows, empty rows at block boundaries) and require the same answer as the one-block path."""
    import dgsparse._lib as L
    rowptr, col = graphs.reddit_like(1 / 32)
    M = rowptr.size - 1
    N = 64
    val = graphs.uniform(col.size, 1, 0.5, 1.5)
    B = graphs.uniform(M * N, 2, -1, 1).reshape(M, N)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    outs = []
    for blk in (str(1 << 40), str(max(1000, col.size // 11))):
        monkeypatch.setenv("DGS_HOST_BLOCK_NNZ", blk)
        C = np.full((M, N), np.nan, np.float32)
        E = np.full((M, N), -7, np.int32)
        L.check(L.lib.dgs_spmm_csr_host(M, M, N, col.size, p(rowptr), p(col), p(val), p(B), p(C), p(E), 1, 2), "host max")
        Cs = np.full((M, N), np.nan, np.float32)
        L.check(L.lib.dgs_spmm_csr_host(M, M, N, col.size, p(rowptr), p(col), p(val), p(B), p(Cs), None, 0, 2), "host sum")
        outs.append((C, E, Cs))
    ref, Eref = oracle.spmm(rowptr, col, val, B, "max", with_arg=True)
    for C, E, Cs in outs:
        assert np.array_equal(C, ref) and np.array_equal(E, Eref)
        assert_close_f32(Cs, oracle.spmm(rowptr, col, val, B), oracle.spmm_f64(rowptr, col, val, B), what="host sum blocks",
                         absref=spmm_absref(oracle, rowptr, col, val, B))


def test_reddit_like_scaled_parity(K, oracle, graphs):
    """Same generator as the bench workload (config 2) at 1/16 scale: hub rows of thousands of nnz."""
    rowptr, col = graphs.reddit_like(1 / 16)
    M = rowptr.size - 1
    val = graphs.uniform(col.size, 1)
    B = graphs.uniform(M * 64, 2).reshape(M, 64)
    out = K.spmm(dev(rowptr), dev(col), dev(val), dev(B)).cpu().numpy()
    ref64 = oracle.spmm_f64(rowptr, col, val, B)
    assert_close_f32(out, oracle.spmm(rowptr, col, val, B), ref64, what="reddit/16", absref=ref64)   # inputs >= 0


def test_full_size_properties(K, graphs):
    """BASELINE config 2 at full size (232 965 rows, 114.6 M nnz, feat 64): size-independent checks.
    (1) column-sum identity  1^T (A B) = (A^T 1)^T B  in fp64;  (2) linearity in B;  (3) a sampled set
    of rows recomputed with torch in fp64;  (4) determinism: two runs give identical bytes."""
    rowptr, col = graphs.reddit_like(1.0)
    M, nnz = rowptr.size - 1, col.size
    rp, cc = dev(rowptr), dev(col)
    val = torch.rand(nnz, device="cuda", generator=torch.Generator("cuda").manual_seed(1))
    B = torch.rand(M, 64, device="cuda", generator=torch.Generator("cuda").manual_seed(2))
    C = K.spmm(rp, cc, val, B)
    w = torch.zeros(M, dtype=torch.float64, device="cuda").index_add_(0, cc.long(), val.double())
    lhs = C.double().sum(0)
    rhs = (w[:, None] * B.double()).sum(0)
    assert torch.allclose(lhs, rhs, rtol=1e-6), (lhs - rhs).abs().max()
    B2 = torch.rand(M, 64, device="cuda", generator=torch.Generator("cuda").manual_seed(3))
    C2 = K.spmm(rp, cc, val, B2)
    C12 = K.spmm(rp, cc, val, B + 2 * B2)
    assert torch.allclose(C12, C + 2 * C2, rtol=2e-5, atol=0.0)   # three fp32 results combined, inputs >= 0
    rows = np.unique(np.concatenate([np.arange(0, M, 2111), np.argsort(np.diff(rowptr))[-8:]]))
    for r in rows:
        s, e = int(rowptr[r]), int(rowptr[r + 1])
        ref = (val[s:e].double()[:, None] * B[cc[s:e].long()].double()).sum(0)
        assert torch.allclose(C[r].double(), ref, rtol=1e-5, atol=0.0), r   # inputs >= 0: purely relative
    assert torch.equal(C, K.spmm(rp, cc, val, B))


@pytest.mark.parametrize("panel", ["32"])
@pytest.mark.parametrize("N", [16, 64, 100, 128, 136, 256])
def test_narrow_panels_same_as_wide(K, oracle, graphs, knob, panel, N):
    """Column panels narrower than 64 (option spmm_panel = 32: the 8-lane row-segment geometry on matrices wider than its panel;
    the 8 / 16-column kernels were a measured dead end, csrc/spmm.cu pick_panel).  Every reduce, with and without edge
    values; max / min and their arg index bit-exact, sum / mean within the stated tolerance (segment boundaries, hence the
    places where a long row's partial sums are folded, differ with the group width, so sums are not bit-identical across
    panel widths)."""
    M, Kc = 4000, 3500
    rowptr, col = graphs.random_csr(M, Kc, 130000, 300 + N, empty_frac=0.25, hub=2)
    val = graphs.uniform(col.size, 1, 0.5, 1.5)
    B = graphs.uniform(Kc * N, 2, -1.0, 1.0).reshape(Kc, N)
    d = [dev(rowptr), dev(col), dev(val), dev(B)]
    for reduce in ("sum", "mean", "max", "min"):
        for v in (d[2], None):
            wa = reduce in ("max", "min")
            knob("spmm_panel", panel)
            got = K.spmm(d[0], d[1], v, d[3], RED[reduce], COMP["mul"], with_arg=wa)
            hv = None if v is None else val
            if wa:
                ref, Eref = oracle.spmm(rowptr, col, hv, B, reduce, "mul", with_arg=True)
                assert np.array_equal(got[0].cpu().numpy(), ref) and np.array_equal(got[1].cpu().numpy(), Eref)
            else:
                assert_close_f32(got.cpu().numpy(), oracle.spmm(rowptr, col, hv, B, reduce), oracle.spmm_f64(rowptr, col, hv, B, reduce),
                                 what=f"panel {panel} N={N} {reduce}", absref=spmm_absref(oracle, rowptr, col, hv, B, reduce))


def test_products_like_full_size(K, oracle, graphs):
    """BASELINE config 3 at full size: products-like CSR (2 449 029 rows, 123.7 M nnz), feat 128, the gspmm-fp ops
    copy_u_max / copy_u_mean / u_mul_e_max / u_mul_e_mean (example/gspmm-fp/util.py); B (1.25 GB) is 10x the L2.  Checks: (1) identity with the reference's own gspmm-fp module where it
    was built (max bit-identical, mean 1e-5), (2) ~200 sampled rows incl. the 8 longest against fp64, (3) determinism."""
    import dgsparse.gspmm as G
    rowptr, col = graphs.products_like(1.0)
    M, nnz, N = rowptr.size - 1, col.size, 128
    rp, cc = dev(rowptr), dev(col)
    val = torch.rand(nnz, device="cuda", generator=torch.Generator("cuda").manual_seed(1)) + 0.5
    B = torch.rand(M, N, device="cuda", generator=torch.Generator("cuda").manual_seed(2)) * 2 - 1
    REF = oracle.ref_gspmm_module()
    rows = np.unique(np.concatenate([np.arange(0, M, M // 190), np.argsort(np.diff(rowptr))[-8:]]))
    for name in ("copy_u_max", "copy_u_mean", "u_mul_e_max", "u_mul_e_mean"):
        red = name.rsplit("_", 1)[1]
        has_val = name.startswith("u_mul")
        out = getattr(G, name)(rp, cc, val.reshape(-1, 1), B) if has_val else getattr(G, name)(rp, cc, B)
        assert tuple(out.shape) == (M, N)
        if REF is not None:
            theirs = (REF.GSpMM_u_e(rp, cc, val, B, getattr(REF.REDUCEOP, red.upper()), REF.COMPUTEOP.MUL) if has_val
                      else REF.GSpMM_u(rp, cc, B, getattr(REF.REDUCEOP, red.upper())))
            if red == "max":
                assert torch.equal(out, theirs), name
            else:   # two fp32 means of ~50 signed terms: compare both to fp64 below, and each other loosely relative to |terms|
                assert torch.allclose(out, theirs, rtol=1e-5, atol=1e-5), name
            del theirs
        for r in rows:
            s, e = int(rowptr[r]), int(rowptr[r + 1])
            if e == s:
                assert (out[r] == 0).all()
                continue
            t32 = B[cc[s:e].long()]
            if has_val:
                t32 = t32 * val[s:e][:, None]                      # one correctly rounded fp32 product per term, as the kernel
            terms = t32.double()
            if red == "max":
                assert torch.equal(out[r], t32.max(0).values), (name, int(r))
            else:
                want, mag = terms.mean(0), terms.abs().mean(0)
                assert ((out[r].double() - want).abs() <= 1e-5 * want.abs() + 1e-6 * mag).all(), (name, int(r))
        again = getattr(G, name)(rp, cc, val.reshape(-1, 1), B) if has_val else getattr(G, name)(rp, cc, B)
        assert torch.equal(out, again), name
        del out, again


def test_legacy_spmm_cuda_without_host_sync_survives_in_place_rewrites(oracle, graphs):
    """spmm_cuda(m, k, rowptr, ...) has no nnz argument.  Ours reads rowptr[m] with a blocking copy only the first time it
    sees a (rowptr, m) pair; afterwards the grid is sized from a hint and the kernels take the true nnz from the device.
    Rewrite the CSR IN PLACE under the same pointers — more nonzeros, fewer, back — and require the right answer on every
    call, including the first one after each rewrite (stale hint: longer segments / idle groups, never a wrong result)."""
    import dgsparse._lib as L
    M, Kc, N = 3000, 2500, 32
    mats = [graphs.random_csr(M, Kc, n, s, empty_frac=0.2, hub=h) for n, s, h in ((40000, 1, 1), (140000, 2, 2), (9000, 3, 0), (400, 4, 0))]
    cap = max(c.size for _, c in mats)
    rp = torch.zeros(M + 1, dtype=torch.int32, device="cuda")
    cc = torch.zeros(cap, dtype=torch.int32, device="cuda")
    vv = torch.zeros(cap, dtype=torch.float32, device="cuda")
    B = graphs.uniform(Kc * N, 9, -1, 1).reshape(Kc, N)
    dB = dev(B)
    out = torch.empty(M, N, device="cuda")
    for which in (0, 0, 1, 1, 1, 2, 2, 3, 3, 0, 1):
        rowptr, col = mats[which]
        val = graphs.uniform(col.size, 10 + which, 0.5, 1.5)
        rp.copy_(torch.from_numpy(rowptr)); cc[:col.size].copy_(torch.from_numpy(col)); vv[:col.size].copy_(torch.from_numpy(val))
        out.fill_(float("nan"))
        torch.cuda.synchronize()
        L.lib.spmm_cuda(M, N, rp.data_ptr(), cc.data_ptr(), vv.data_ptr(), dB.data_ptr(), out.data_ptr())
        torch.cuda.synchronize()
        assert_close_f32(out.cpu().numpy(), oracle.spmm(rowptr, col, val, B), oracle.spmm_f64(rowptr, col, val, B),
                         what=f"in-place matrix {which}", absref=spmm_absref(oracle, rowptr, col, val, B))


@pytest.mark.parametrize("N", [4, 16, 32, 64, 100, 128, 256])
def test_row_parallel_kernel_forced(K, oracle, graphs, knob, N):
    """spmm_rowpar_kernel (single launch, a lane group per row) forced on with option spmm_rowpar = 1: every reduce, with and
    without edge values, empty rows, and rows far longer than it would ever be chosen for (hub = 2) — it must be correct
    for ANY matrix, its selection is only a performance decision."""
    import dgsparse._lib as L
    knob("spmm_rowpar", 1)
    M, Kc = 3000, 2500
    rowptr, col = graphs.random_csr(M, Kc, 60000, 500 + N, empty_frac=0.3, hub=2)
    val = graphs.uniform(col.size, 1, 0.5, 1.5)
    B = graphs.uniform(Kc * N, 2, -1.0, 1.0).reshape(Kc, N)
    d = [dev(rowptr), dev(col), dev(val), dev(B)]
    for reduce in ("sum", "mean", "max", "min"):
        for v, hv in ((d[2], val), (None, None)):
            wa = reduce in ("max", "min")
            got = K.spmm(d[0], d[1], v, d[3], RED[reduce], COMP["mul"], with_arg=wa)
            assert L.lib.dgs_spmm_last_path() == 1
            if wa:
                ref, Eref = oracle.spmm(rowptr, col, hv, B, reduce, "mul", with_arg=True)
                assert np.array_equal(got[0].cpu().numpy(), ref) and np.array_equal(got[1].cpu().numpy(), Eref)
            else:
                assert_close_f32(got.cpu().numpy(), oracle.spmm(rowptr, col, hv, B, reduce), oracle.spmm_f64(rowptr, col, hv, B, reduce),
                                 what=f"rowpar N={N} {reduce}", absref=spmm_absref(oracle, rowptr, col, hv, B, reduce))


def test_spmm_inside_cuda_graph_capture(K, oracle, graphs):
    """The library may be captured into a CUDA graph: while a stream is capturing it must not allocate, record or query
    anything (graph notes are skipped, the segment path is taken), and replays must follow new operand VALUES."""
    import dgsparse._lib as L
    rowptr, col, (M, Kc) = graphs.load_fixture("p2p-Gnutella31")
    N = 32
    val = graphs.uniform(col.size, 1)
    rp, cc, vv = dev(rowptr), dev(col), dev(val)
    L.lib.dgs_spmm_forget_graph_notes()
    B0 = graphs.uniform(Kc * N, 2).reshape(Kc, N)
    dB = dev(B0)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):                      # warm-up outside the capture (lazy module load), as torch asks
        K.spmm(rp, cc, vv, dB)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = K.spmm(rp, cc, vv, dB)
        assert L.lib.dgs_spmm_last_path() == 0        # never the note-driven kernel inside a capture
    for seed in (2, 7):
        Bn = graphs.uniform(Kc * N, seed).reshape(Kc, N)
        dB.copy_(torch.from_numpy(Bn))
        g.replay()
        torch.cuda.synchronize()
        assert_close_f32(out.cpu().numpy(), oracle.spmm(rowptr, col, val, Bn), what=f"graph replay {seed}")


def test_row_parallel_selected_by_graph_note_and_healed(K, oracle, graphs):
    """The library learns from its own fix-up scan that a matrix has short rows only (p2p-Gnutella31: longest row 78) and
    switches later calls to the single-launch kernel; a matrix with a long row never switches; and a CSR rewritten IN PLACE
    with a long row under the same pointer sends the calls back to the segment path after the row-parallel kernel met it."""
    import dgsparse._lib as L
    rowptr, col, (M, Kc) = graphs.load_fixture("p2p-Gnutella31")
    N = 32
    val = graphs.uniform(col.size, 1)
    B = graphs.uniform(Kc * N, 2).reshape(Kc, N)
    ref = oracle.spmm(rowptr, col, val, B)
    rp, cc, vv, dB = dev(rowptr), dev(col), dev(val), dev(B)
    L.lib.dgs_spmm_forget_graph_notes()       # notes are keyed on device pointers, which the allocator recycles across tests
    paths = []
    for _ in range(6):
        out = K.spmm(rp, cc, vv, dB)
        paths.append(L.lib.dgs_spmm_last_path())
        torch.cuda.synchronize()
        assert_close_f32(out.cpu().numpy(), ref, what="gnutella")
    assert paths[0] == 0 and paths[-1] == 1, paths
    # a hub matrix stays on the segment path
    hr, hc = graphs.random_csr(4000, 4000, 100000, 3, hub=1)
    hrp, hcc, hB = dev(hr), dev(hc), dev(graphs.uniform(4000 * N, 4).reshape(4000, N))
    for _ in range(4):
        K.spmm(hrp, hcc, None, hB)
        torch.cuda.synchronize()
        assert L.lib.dgs_spmm_last_path() == 0
    # rewrite the Gnutella CSR in place: one row of 5000 nonzeros (same rowptr pointer and M)
    r2 = rowptr.copy()
    big = int(np.argmax(np.diff(rowptr)))
    grow = 5000 - int(rowptr[big + 1] - rowptr[big])
    r2[big + 1:] += grow
    c2 = np.concatenate([col[:rowptr[big]], np.sort(np.random.default_rng(0).choice(Kc, 5000, replace=False)).astype(np.int32),
                         col[rowptr[big + 1]:]])
    v2 = graphs.uniform(c2.size, 5)
    cc2, vv2 = dev(c2), dev(v2)
    rp.copy_(torch.from_numpy(r2))
    ref2 = oracle.spmm(r2, c2, v2, B)
    paths = []
    for _ in range(4):
        out = K.spmm(rp, cc2, vv2, dB)
        paths.append(L.lib.dgs_spmm_last_path())
        torch.cuda.synchronize()
        assert_close_f32(out.cpu().numpy(), ref2, oracle.spmm_f64(r2, c2, v2, B), what="rewritten", absref=oracle.spmm_f64(r2, c2, v2, B))
    assert paths[0] == 1 and paths[-1] == 0, paths



def test_resident_csr_host_entry(oracle, graphs):
    """dgs_csr_upload + dgs_spmm_csr_resident_host: the CSR crosses PCIe once, every call moves only B in and C (and E) out.
    Several B's, two feature widths (the device scratch regrows), sum and max + arg, a matrix large enough for several row
    blocks with empty rows at block boundaries; results must equal the per-call host entry's."""
    import dgsparse._lib as L
    rowptr, col = graphs.reddit_like(1 / 4)          # 28.6 M nnz -> 3 row blocks
    M = rowptr.size - 1
    val = graphs.uniform(col.size, 1, 0.5, 1.5)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    h = ctypes.c_void_p()
    L.check(L.lib.dgs_csr_upload(M, M, col.size, p(rowptr), p(col), p(val), ctypes.byref(h)), "dgs_csr_upload")
    try:
        for N, seed in ((64, 2), (32, 3), (64, 4)):
            B = graphs.uniform(M * N, seed, -1, 1).reshape(M, N)
            C = np.full((M, N), np.nan, np.float32)
            L.check(L.lib.dgs_spmm_csr_resident_host(h, N, p(B), p(C), None, 0, 2), "resident sum")
            want = np.empty((M, N), np.float32)
            L.check(L.lib.dgs_spmm_csr_host(M, M, N, col.size, p(rowptr), p(col), p(val), p(B), p(want), None, 0, 2), "host sum")
            assert np.array_equal(C, want)                       # same kernels, same row blocks
            rows = np.unique(np.concatenate([np.arange(0, M, 997), np.argsort(np.diff(rowptr))[-4:]]))
            sub_rp = np.concatenate([[0], np.cumsum(np.diff(rowptr)[rows])]).astype(np.int32)
            sub_idx = np.concatenate([np.arange(rowptr[r], rowptr[r + 1]) for r in rows])
            ref64 = oracle.spmm_f64(sub_rp, col[sub_idx], val[sub_idx], B)
            assert_close_f32(C[rows], oracle.spmm(sub_rp, col[sub_idx], val[sub_idx], B), ref64, what=f"resident N={N}",
                             absref=spmm_absref(oracle, sub_rp, col[sub_idx], val[sub_idx], B))
        N = 64
        B = graphs.uniform(M * N, 9, -1, 1).reshape(M, N)
        C, E = np.empty((M, N), np.float32), np.empty((M, N), np.int32)
        L.check(L.lib.dgs_spmm_csr_resident_host(h, N, p(B), p(C), p(E), 1, 2), "resident max")
        ref, Eref = oracle.spmm(rowptr, col, val, B, "max", with_arg=True)
        assert np.array_equal(C, ref) and np.array_equal(E, Eref)
    finally:
        L.lib.dgs_csr_free(h)

