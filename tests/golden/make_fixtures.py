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


if __name__ == "__main__":
    main()
