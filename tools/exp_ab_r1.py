#!/usr/bin/env python
"""A/B on ONE box: the round-1 library (tools/_r1lib/libdgsparse_b200_r1.so, built from commit 729da3c; not tracked) against the
current one (build it with: git archive 729da3c dgsparse-lib_b200 include | tar -x -C /tmp/r1; python /tmp/r1/dgsparse-lib_b200/build.py; copy the .so), same buffers, same timing loop, dgs_spmm_csr at several widths."""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dgsparse-lib_b200")]
import dgsparse._lib as L  # noqa: E402
from tools import graphs  # noqa: E402
from tools.bench_vs_ref import timeit  # noqa: E402

old = ctypes.CDLL(os.path.join(ROOT, "tools", "_r1lib", "libdgsparse_b200_r1.so"))
sig = L.SIGNATURES["dgs_spmm_csr"]
old.dgs_spmm_csr.restype, old.dgs_spmm_csr.argtypes = sig
old.dgs_spmm_workspace_bytes.restype, old.dgs_spmm_workspace_bytes.argtypes = L.SIGNATURES["dgs_spmm_workspace_bytes"]
for gname, gen in (("reddit", graphs.reddit_like), ("products", graphs.products_like)):
    rowptr, col = gen(1.0)
    M, nnz = rowptr.size - 1, int(col.size)
    rp, cc = torch.from_numpy(rowptr).cuda(), torch.from_numpy(col).cuda()
    vv = torch.rand(nnz, device="cuda")
    for N in (32, 64, 128, 256):
        B = torch.rand(M, N, device="cuda")
        out = torch.empty(M, N, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        res = {}
        for name, lib in (("r1", old), ("now", L.lib), ("r1_again", old), ("now_again", L.lib)):
            ws = torch.empty(lib.dgs_spmm_workspace_bytes(N, nnz, 0), dtype=torch.uint8, device="cuda")
            res[name] = timeit(lambda: lib.dgs_spmm_csr(M, N, nnz, rp.data_ptr(), cc.data_ptr(), vv.data_ptr(), B.data_ptr(), N,
                                                       out.data_ptr(), N, None, 0, 0, 2, ws.data_ptr(), ws.numel(), st), 20)
        print(json.dumps({"graph": gname, "N": N, **{k: round(v, 4) for k, v in res.items()}}), flush=True)
