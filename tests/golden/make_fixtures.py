"""Regenerates tests/golden/*.npz.  Run in the BUILD container only (needs /root/reference):

    python tests/golden/make_fixtures.py

1. <name>.npz            CSR (rowptr, col, shape) of /root/reference/example/data/<name>.mtx, loaded the
                         way the reference's own test does (test/test_csr2csr.py:20-28: scipy
                         mmread().tocsr()); plus the scipy tocsc() transpose (colptr, row, perm) that
                         test pins csr2csc against (test/test_csr2csr.py:42-49).
2. <name>_spmm32.npz     golden SpMM output at feat=32 produced by the UNMODIFIED reference host
                         function spmm_reference_host<int,float> (example/util/sp_util.hpp:62-84, built
                         as oracle/_ref/libref_host.so): a strided row sample, the row ids, and fp64
                         column sums of the whole output as a checksum.  Inputs are regenerated from
                         seeds by tools/graphs.uniform().
3. <name>_sddmm32.npz    golden SDDMM output (sddmm_reference_host, sp_util.hpp:87-112), K=32.
4. spconv_*.npz          the kernel maps of example/data/sample-data/fp32/minkunet-semantickitti/*.pth
                         (kpos, imap, omap, sizes) re-saved without pickle, for the spconv tier.
5. spconv_cpu_compute.npz  golden OUTPUTS of the reference's own `cpu_compute` (test/test_spconv.py:17-53, imported
                         from /root/reference/test/test_spconv.py and run unmodified) on a sub-layer cut out of the
                         c_in = 4 MinkUNet fixture: the pairs among ~500 voxels (a closed sub-layer, in_nnz == out_nnz), renumbered so that
                         the Python quadruple loop finishes in about a minute.  Both call forms: all taps in the maps
                         (precompute = False) and the centre tap taken out and added by `precompute = True`.
                         tests/test_oracle_golden.py pins oracle_spconv to these bit for bit.
"""
import os
import sys

import numpy as np
from scipy.io import mmread

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from tools import graphs   # noqa: E402

REF = "/root/reference/example/data"
OUT = os.path.dirname(os.path.abspath(__file__))

SEED_VAL, SEED_B, SEED_D1, SEED_D2 = 11, 12, 13, 14


def main():
    for name in ["p2p-Gnutella31", "ca-CondMat"]:
        A = mmread(os.path.join(REF, name + ".mtx")).tocsr().astype(np.float32)
        A.sort_indices()
        rowptr, col = A.indptr.astype(np.int32), A.indices.astype(np.int32)
        M, K = A.shape
        # scipy transpose with an exact permutation carried as float64 payload
        P = A.copy()
        P.data = np.arange(P.nnz, dtype=np.float64)
        C = P.tocsc()
        np.savez_compressed(os.path.join(OUT, name + ".npz"), rowptr=rowptr, col=col,
                            shape=np.array(A.shape, np.int64), colptr=C.indptr.astype(np.int32),
                            row=C.indices.astype(np.int32), perm=C.data.astype(np.int32))
        N = 32
        val = graphs.uniform(col.size, SEED_VAL)
        B = graphs.uniform(K * N, SEED_B).reshape(K, N)
        out = oracle.ref_spmm_host(rowptr, col, val, B)
        rows = np.unique(np.concatenate([np.arange(0, M, 97), np.argsort(np.diff(rowptr))[-64:],
                                         np.arange(min(M, 64))])).astype(np.int32)
        np.savez_compressed(os.path.join(OUT, name + "_spmm32.npz"), rows=rows, out_rows=out[rows],
                            colsum=out.astype(np.float64).sum(0),
                            seeds=np.array([SEED_VAL, SEED_B], np.int64))
        D1 = graphs.uniform(M * N, SEED_D1).reshape(M, N)
        D2 = graphs.uniform(K * N, SEED_D2).reshape(K, N)
        sd = oracle.ref_sddmm_host(rowptr, col, D1, D2)
        np.savez_compressed(os.path.join(OUT, name + "_sddmm32.npz"), out=sd.astype(np.float32),
                            seeds=np.array([SEED_D1, SEED_D2], np.int64))
        print(name, A.shape, A.nnz, "empty rows", int((np.diff(rowptr) == 0).sum()))

    try:
        import torch
        base = os.path.join(REF, "sample-data/fp32/minkunet-semantickitti")
        for i, f in enumerate(sorted(os.listdir(base))):
            d = torch.load(os.path.join(base, f), weights_only=False, map_location="cpu")
            keep = {}
            for k, v in d.items():
                if torch.is_tensor(v):
                    keep[k] = v.cpu().numpy()
                elif isinstance(v, (int, float)):
                    keep[k] = np.array(v)
            np.savez_compressed(os.path.join(OUT, f"spconv_fp32_{i}.npz"), **keep)
            print("spconv", f, {k: (v.shape, v.dtype) for k, v in keep.items()})
    except Exception as e:  # pragma: no cover
        print("spconv fixtures skipped:", e)
    spconv_cpu_compute_golden()


def spconv_cpu_compute_golden(n_out=500):
    """Run the reference's own cpu_compute on a sub-layer of the c_in = 4 fixture and store inputs + outputs."""
    import importlib.util
    import torch
    spec = importlib.util.spec_from_file_location("ref_test_spconv", "/root/reference/test/test_spconv.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)                      # defines kpos_quantized / cpu_compute / remove_mid / test_spconv only
    g = np.load(os.path.join(OUT, "spconv_fp32_0.npz"))
    k_vol, c_in, c_out = int(g["k_vol"]), int(g["c_in"]), int(g["c_out"])
    kpos, imap, omap = g["kpos"], g["imap"], g["omap"]
    assert c_in == 4 and int(g["in_nnz"]) == int(g["out_nnz"])
    rng = np.random.default_rng(2024)
    # ~n_out output voxels: a contiguous run (neighbours of neighbours are likely inside) plus scattered ones
    start = int(rng.integers(0, int(g["out_nnz"]) - n_out))
    keep_out = np.unique(np.concatenate([np.arange(start, start + n_out - 40), rng.choice(int(g["out_nnz"]), 40, replace=False)]))
    # a CLOSED sub-layer: pairs whose input and output voxel are both kept, so in_nnz == out_nnz == |kept| and the
    # separate_mid form (which the op only accepts for in_nnz == out_nnz, src/cuda/spconv_cuda.cu:61-82) can be driven too
    sel = np.isin(omap, keep_out) & np.isin(imap, keep_out)
    tap = np.repeat(np.arange(k_vol), np.diff(kpos))
    order = keep_out
    newid = -np.ones(int(g["in_nnz"]), np.int64)
    newid[order] = np.arange(order.size)
    imap_s, omap_s, tap_s = newid[imap[sel]].astype(np.int32), newid[omap[sel]].astype(np.int32), tap[sel]
    knnz_s = np.bincount(tap_s, minlength=k_vol).astype(np.int32)
    kpos_s = np.concatenate([[0], np.cumsum(knnz_s)]).astype(np.int32)
    feats = rng.uniform(-1, 1, (order.size, c_in)).astype(np.float32)
    W = rng.uniform(-1, 1, (k_vol, c_in, c_out)).astype(np.float32)
    out_size = keep_out.size
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    full = ref.cpu_compute(t(feats), t(W), out_size, t(knnz_s), t(imap_s), t(omap_s), False).numpy()
    # the separate_mid form: centre tap removed from the maps, added by precompute=True
    mid = k_vol // 2
    s, e = int(kpos_s[mid]), int(kpos_s[mid + 1])
    assert np.array_equal(imap_s[s:e], omap_s[s:e]) and e - s == out_size     # the centre tap is the identity map
    imap_m, omap_m = np.delete(imap_s, np.s_[s:e]), np.delete(omap_s, np.s_[s:e])
    knnz_m = knnz_s.copy()
    knnz_m[mid] = 0
    sep = ref.cpu_compute(t(feats), t(W), out_size, t(knnz_m), t(imap_m), t(omap_m), True).numpy()
    np.savez_compressed(os.path.join(OUT, "spconv_cpu_compute.npz"), feats=feats, W=W, out_size=np.array(out_size),
                        knnz=knnz_s, imap=imap_s, omap=omap_s, out_full=full,
                        knnz_mid=knnz_m, imap_mid=imap_m, omap_mid=omap_m, out_separate_mid=sep)
    print("spconv cpu_compute golden:", out_size, "outputs,", order.size, "inputs,", int(knnz_s.sum()), "pairs,",
          "max |full - separate_mid| =", float(np.abs(full - sep).max()))


if __name__ == "__main__":
    main()
