#!/usr/bin/env python
"""Experiment: SDDMM on arxiv-like K=256 as feature-axis passes (K/p columns per pass, leading dimension 256) — does keeping
a narrower D2 slice L2-resident pay for walking the edge list p times?  Prints ms per pass and per full product."""
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
    M, nnz, K = rowptr.size - 1, int(col.size), 256
    rp, cc = torch.from_numpy(rowptr).cuda(), torch.from_numpy(col).cuda()
    D1, D2 = torch.rand(M, K, device="cuda"), torch.rand(M, K, device="cuda")
    out = torch.empty(nnz, device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def run(p):
        w = K // p
        for i in range(p):
            L.check(L.lib.dgs_sddmm_csr(M, w, nnz, rp.data_ptr(), cc.data_ptr(), D1.data_ptr() + 4 * w * i, K,
                                        D2.data_ptr() + 4 * w * i, K, None, 0, out.data_ptr(), st), "sddmm")
    for p in (1, 2, 4):
        for _ in range(5):
            run(p)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50):
            run(p)
        b.record()
        torch.cuda.synchronize()
        print(f"passes={p} width={K // p}: {a.elapsed_time(b) / 50:.4f} ms per full product", flush=True)


if __name__ == "__main__":
    main()
