#!/usr/bin/env python
"""SpMM on the arxiv-like graph (1.17 M nnz: between the latency and the bandwidth regime), a few calls per width — for ncu launch lists."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dgsparse-lib_b200"))


def main():
    import dgsparse._lib as L
    from tools import graphs
    graphs.build()
    rowptr, col = graphs.arxiv_like(1.0)
    M, nnz = rowptr.size - 1, int(col.size)
    import numpy as np
    print("empty rows:", int((np.diff(rowptr) == 0).sum()), "of", M, flush=True)
    rp, cc = torch.from_numpy(rowptr).cuda(), torch.from_numpy(col).cuda()
    vv = torch.rand(nnz, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for N in (32, 128):
        B = torch.rand(M, N, device="cuda")
        C = torch.empty(M, N, device="cuda")
        ws = torch.empty(L.lib.dgs_spmm_workspace_bytes(N, nnz, 0), dtype=torch.uint8, device="cuda")
        for _ in range(4):
            L.lib.dgs_spmm_csr(M, N, nnz, rp.data_ptr(), cc.data_ptr(), vv.data_ptr(), B.data_ptr(), N, C.data_ptr(), N,
                               None, 0, L.SUM, L.MUL, ws.data_ptr(), ws.numel(), st)
        torch.cuda.synchronize()


if __name__ == "__main__":
    main()
