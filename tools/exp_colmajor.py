#!/usr/bin/env python
"""gespmmCsrSpMM(transpose_BC = false): column-major operands.  Times the thread-per-element kernel (spmm_colmajor = 0), the
row-major kernel between two tiled transposes (= 1), the library's own choice (unset) and the reference's CUDA
(oracle/_ref/libref_cuda.so, src/ge-spmm/csrspmm_non_transpose.cu) on the same buffers.  One JSON line per case.

    python tools/exp_colmajor.py [--reps 20] [--scale 0.25]
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
    ap.add_argument("--scale", type=float, default=0.25, help="scale of the reddit-like case")
    args = ap.parse_args()
    import dgsparse._lib as L
    from oracle import oracle
    from tools import graphs
    graphs.build()
    R = oracle.ref_cuda_lib()
    cases = [("p2p-Gnutella31 (example/data)", graphs.load_fixture("p2p-Gnutella31")[:2], (32, 128)),
             ("arxiv-like", graphs.arxiv_like(1.0), (32, 128)),
             ("reddit-like x%g" % args.scale, graphs.reddit_like(args.scale), (64,))]
    for gname, (rowptr, col), widths in cases:
        M, nnz = rowptr.size - 1, int(col.size)
        rp, cc = torch.from_numpy(rowptr).cuda(), torch.from_numpy(col).cuda()
        vv = torch.rand(nnz, device="cuda")
        d = L.SpMatCsrDescr_t(M, M, nnz, rp.data_ptr(), cc.data_ptr(), vv.data_ptr())
        for N in widths:
            Bt = torch.rand(N, M, device="cuda")          # column-major B[M, N]
            rec = {"op": "gespmmCsrSpMM column-major", "graph": gname, "M": M, "nnz": nnz, "N": N}
            outs = {}
            for name, opt in (("naive", 0), ("transposed", 1), ("chosen", -1)):
                L.lib.dgs_set_option(b"spmm_colmajor", opt)
                Ct = torch.empty(N, M, device="cuda")
                reps = args.reps if (opt != 0 or nnz * N < 2e9) else 3
                rec[name + "_ms"] = timeit(lambda: L.lib.gespmmCsrSpMM(d, Bt.data_ptr(), N, Ct.data_ptr(), False, 0), reps)
                outs[name] = Ct
            L.lib.dgs_set_option(b"spmm_colmajor", -1)
            if R is not None:
                Cr = torch.zeros(N, M, device="cuda")
                rd = oracle.SpMatCsrDescr(M, M, nnz, rp.data_ptr(), cc.data_ptr(), vv.data_ptr())
                # 10 = GESPMM_ALG_DEFAULT (-> PARREDUCE_ROWBALANCE_NON_TRANSPOSE), 4 = SEQREDUCE_ROWBALANCE_NON_TRANSPOSE: the two
                # non-transposed algorithms that overwrite C (src/ge-spmm/gespmm.cc:92-108; a transposed code exits the process)
                for alg in (10, 4):
                    rec["reference_cuda_alg%d_ms" % alg] = timeit(
                        lambda: R.gespmmCsrSpMM(rd, Bt.data_ptr(), N, Cr.data_ptr(), False, alg), max(3, args.reps // 4))
                rec["reference_cuda_ms"] = min(rec["reference_cuda_alg10_ms"], rec["reference_cuda_alg4_ms"])
                rec["max_rel_diff_vs_reference"] = float(((outs["chosen"] - Cr).abs() / Cr.abs().clamp_min(1e-6)).max())
                rec["speedup_chosen"] = rec["reference_cuda_ms"] / rec["chosen_ms"]
            rec["max_rel_diff_naive_vs_transposed"] = float(((outs["naive"] - outs["transposed"]).abs()
                                                            / outs["naive"].abs().clamp_min(1e-6)).max())
            print(json.dumps(rec), flush=True)
            del Bt, outs
        del rp, cc, vv
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
