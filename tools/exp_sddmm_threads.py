#!/usr/bin/env python
"""Register-staged SDDMM kernel (K < 64, unaligned K, masked): 256- vs 64-thread CTAs (dgs_set_option sddmm_threads)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dgsparse-lib_b200")]
import dgsparse._lib as L  # noqa: E402
from tools import graphs  # noqa: E402
from tools.bench_vs_ref import timeit  # noqa: E402

graphs.build()
st = torch.cuda.current_stream().cuda_stream
cases = [("arxiv-like", graphs.arxiv_like(1.0), 50), ("ca-CondMat (example/data)", graphs.load_fixture("ca-CondMat")[:2], 200),
         ("reddit-like x0.1", graphs.reddit_like(0.1), 20)]
for gname, (rowptr, col), reps in cases:
    M, nnz = rowptr.size - 1, int(col.size)
    rp, cc = torch.from_numpy(rowptr).cuda(), torch.from_numpy(col).cuda()
    for Kd in (16, 32, 48, 100):
        D1, D2 = torch.rand(M, Kd, device="cuda"), torch.rand(M, Kd, device="cuda")
        out = torch.empty(nnz, device="cuda")
        run = lambda: L.lib.dgs_sddmm_csr(M, Kd, nnz, rp.data_ptr(), cc.data_ptr(), D1.data_ptr(), Kd, D2.data_ptr(), Kd, None, 0,
                                          out.data_ptr(), st)
        rec = {"op": "sddmm_csr (register kernel)", "graph": gname, "nnz": nnz, "K": Kd}
        keep = {}
        for th in (256, 64, 256, 64):
            L.lib.dgs_set_option(b"sddmm_threads", th)
            t = timeit(run, reps)
            rec.setdefault("threads%d_ms" % th, []).append(round(t, 5))
            keep[th] = out.clone()
        L.lib.dgs_set_option(b"sddmm_threads", -1)
        rec["bit_identical"] = bool(torch.equal(keep[256], keep[64]))
        print(json.dumps(rec), flush=True)
