#!/usr/bin/env python
"""Ours vs the reference's own CUDA kernels (oracle/_ref/libref_cuda.so, built unmodified for sm_100a) across
feature widths, on the same device buffers with the same timing loop.  One JSON line per case.

    python tools/bench_vs_ref.py [--reps 20] [--scale 1.0]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dgsparse-lib_b200"))


def timeit(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--small", action="store_true", help="only the reference's own small fixtures + the arxiv-like graph (latency regime)")
    args = ap.parse_args()
    import dgsparse._kernels as K
    import dgsparse._lib as L
    from oracle import oracle
    from tools import graphs
    graphs.build()
    R = oracle.ref_cuda_lib()
    assert R is not None, "oracle/_ref/libref_cuda.so missing"
    fixture = lambda name: (lambda scale: graphs.load_fixture(name)[:2])
    cases = [("p2p-Gnutella31 (example/data)", fixture("p2p-Gnutella31"), (32, 64, 128)),
             ("ca-CondMat (example/data)", fixture("ca-CondMat"), (32, 64, 128)),
             ("arxiv-like", graphs.arxiv_like, (32, 64, 128, 256))]
    if not args.small:
        cases = [("reddit-like", graphs.reddit_like, (32, 64, 128, 256)),
                 ("products-like", graphs.products_like, (32, 64, 128, 256))]
    for gname, gen, widths in cases:
        rowptr, col = gen(args.scale)
        M, nnz = rowptr.size - 1, int(col.size)
        rp, cc = torch.from_numpy(rowptr).cuda(), torch.from_numpy(col).cuda()
        vv = torch.rand(nnz, device="cuda")
        for N in widths:
            B = torch.rand(M, N, device="cuda")
            ours = torch.empty(M, N, device="cuda")
            ref = torch.zeros(M, N, device="cuda")
            ws = torch.empty(L.lib.dgs_spmm_workspace_bytes(N, nnz, 0), dtype=torch.uint8, device="cuda")
            st = torch.cuda.current_stream().cuda_stream
            t_o = timeit(lambda: L.lib.dgs_spmm_csr(M, N, nnz, rp.data_ptr(), cc.data_ptr(), vv.data_ptr(), B.data_ptr(), N,
                                                    ours.data_ptr(), N, None, 0, L.SUM, L.MUL, ws.data_ptr(), ws.numel(), st),
                         args.reps)
            t_r = timeit(lambda: R.spmm_cuda(M, N, rp.data_ptr(), cc.data_ptr(), vv.data_ptr(), B.data_ptr(), ref.data_ptr()),
                         max(3, args.reps // 4))
            err = float(((ours - ref).abs() / ref.abs().clamp_min(1e-6)).max())
            print(json.dumps({"op": "spmm_sum", "graph": gname, "M": M, "nnz": nnz, "N": N, "ours_ms": t_o, "reference_cuda_ms": t_r,
                              "speedup": t_r / t_o, "ours_gflops": 2.0 * nnz * N / t_o / 1e6, "max_rel_diff": err}), flush=True)
            del B, ours, ref
        del rp, cc, vv
        torch.cuda.empty_cache()
    sd_cases = [("arxiv-like", graphs.arxiv_like(args.scale))]
    if args.small:
        sd_cases = [("p2p-Gnutella31 (example/data)", graphs.load_fixture("p2p-Gnutella31")[:2]),
                    ("ca-CondMat (example/data)", graphs.load_fixture("ca-CondMat")[:2])]
    for gname, (rowptr, col) in sd_cases:
      M, nnz = rowptr.size - 1, int(col.size)
      rp, cc = torch.from_numpy(rowptr).cuda(), torch.from_numpy(col).cuda()
      for Kd in (32, 64, 128, 256, 512):
          D1, D2 = torch.rand(M, Kd, device="cuda"), torch.rand(M, Kd, device="cuda")
          ours = torch.empty(nnz, device="cuda")
          ref = torch.zeros(nnz, device="cuda")
          t_o = timeit(lambda: L.lib.dgs_sddmm_csr(M, Kd, nnz, rp.data_ptr(), cc.data_ptr(), D1.data_ptr(), Kd, D2.data_ptr(), Kd,
                                                    None, 0, ours.data_ptr(), torch.cuda.current_stream().cuda_stream), args.reps)
          t_r = timeit(lambda: R.sddmm_cuda_csr(M, Kd, nnz, rp.data_ptr(), cc.data_ptr(), D1.data_ptr(), D2.data_ptr(), ref.data_ptr()),
                       args.reps)
          err = float(((ours - ref).abs() / ref.abs().clamp_min(1e-6)).max())
          print(json.dumps({"op": "sddmm_csr", "graph": gname, "M": M, "nnz": nnz, "K": Kd, "ours_ms": t_o, "reference_cuda_ms": t_r,
                            "speedup": t_r / t_o, "max_rel_diff": err}), flush=True)


if __name__ == "__main__":
    main()
