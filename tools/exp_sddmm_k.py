#!/usr/bin/env python
"""SDDMM on the arxiv-like graph across K, ours vs the reference's kernel (the SDDMM half of tools/bench_vs_ref.py alone)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dgsparse-lib_b200")]
import dgsparse._lib as L  # noqa: E402
from oracle import oracle  # noqa: E402
from tools import graphs  # noqa: E402
from tools.bench_vs_ref import timeit  # noqa: E402

R = oracle.ref_cuda_lib()
rowptr, col = graphs.arxiv_like(1.0)
M, nnz = rowptr.size - 1, int(col.size)
rp, cc = torch.from_numpy(rowptr).cuda(), torch.from_numpy(col).cuda()
for Kd in (64, 128, 256, 512):
    D1, D2 = torch.rand(M, Kd, device="cuda"), torch.rand(M, Kd, device="cuda")
    ours, ref = torch.empty(nnz, device="cuda"), torch.zeros(nnz, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    t_o = timeit(lambda: L.lib.dgs_sddmm_csr(M, Kd, nnz, rp.data_ptr(), cc.data_ptr(), D1.data_ptr(), Kd, D2.data_ptr(), Kd, None, 0,
                                             ours.data_ptr(), st), 50)
    t_r = timeit(lambda: R.sddmm_cuda_csr(M, Kd, nnz, rp.data_ptr(), cc.data_ptr(), D1.data_ptr(), D2.data_ptr(), ref.data_ptr()), 20) if R else float("nan")
    print("    arxiv-like", Kd, "ours %.4f ms ref %.4f ms x%.2f" % (t_o, t_r, t_r / t_o), flush=True)
