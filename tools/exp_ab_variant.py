#!/usr/bin/env python
"""A/B in ONE process: the shipped library against a variant built by tools/build_variant.py (same buffers, same timing
loop, interleaved), dgs_spmm_csr at several widths and segment counts per resident lane group.

    python tools/exp_ab_variant.py t64 [--products]
"""
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

tag = sys.argv[1]
var = ctypes.CDLL(os.path.join(ROOT, "tools", "_build", "variant_" + tag, "libdgsparse_b200_%s.so" % tag))
for name in ("dgs_spmm_csr", "dgs_spmm_workspace_bytes", "dgs_set_option"):
    getattr(var, name).restype, getattr(var, name).argtypes = L.SIGNATURES[name]
graphs.build()
cases = [("reddit", graphs.reddit_like, (64, 128, 256))]
if "--products" in sys.argv:
    cases.append(("products", graphs.products_like, (128,)))
for gname, gen, widths in cases:
    rowptr, col = gen(1.0)
    M, nnz = rowptr.size - 1, int(col.size)
    rp, cc = torch.from_numpy(rowptr).cuda(), torch.from_numpy(col).cuda()
    vv = torch.rand(nnz, device="cuda")
    for N in widths:
        B = torch.rand(M, N, device="cuda")
        out = torch.empty(M, N, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        for segs in (-1, 1, 2, 4, 8):
            res, keep = {}, {}
            for name, lib in (("shipped", L.lib), (tag, var), ("shipped_again", L.lib), (tag + "_again", var)):
                lib.dgs_set_option(b"spmm_segs", segs)
                ws = torch.empty(lib.dgs_spmm_workspace_bytes(N, nnz, 0), dtype=torch.uint8, device="cuda")
                res[name] = timeit(lambda: lib.dgs_spmm_csr(M, N, nnz, rp.data_ptr(), cc.data_ptr(), vv.data_ptr(), B.data_ptr(), N,
                                                           out.data_ptr(), N, None, 0, 0, 2, ws.data_ptr(), ws.numel(), st), 20)
                keep[name] = out.clone()
                lib.dgs_set_option(b"spmm_segs", -1)
            print(json.dumps({"graph": gname, "N": N, "spmm_segs": segs if segs > 0 else "library default",
                              **{k: round(v, 4) for k, v in res.items()},
                              "bit_identical": bool(torch.equal(keep["shipped"], keep[tag]))}), flush=True)
        del B, out
    del rp, cc, vv
    torch.cuda.empty_cache()
