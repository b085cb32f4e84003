#!/usr/bin/env python
"""csr2csc: our radix transpose against the reference's own op (cuSPARSE Csr2cscEx2 ALG1 behind
torch.ops.dgsparse_spmm.csr2csc, oracle/_ref/_spmm_cuda.so built unmodified for sm_100a), same CSR, same timing loop.

    python tools/bench_csr2csc.py [--reps 10] [--scale 1.0]            # ours, then the reference in a process of its own
    python tools/bench_csr2csc.py --impl reference --graph reddit-like  # (what the parent spawns)

The two register the same torch.ops namespace, so the reference runs in a child process.  One JSON line per (graph, impl).
Algorithmic bytes = rowptr + col + val in, colptr + row + val_t out (+ the int32 permutation for ours: the reference gets
it by transposing a float arange, exact only below 2^24 nnz, dgsparse/storage.py:164-169).
"""
import argparse
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dgsparse-lib_b200"))

GRAPHS = ("reddit-like", "products-like", "arxiv-like")


def timeit(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def load(graph, scale):
    from tools import graphs
    graphs.build()
    gen = {"reddit-like": graphs.reddit_like, "products-like": graphs.products_like, "arxiv-like": graphs.arxiv_like}[graph]
    rowptr, col = gen(scale)
    return torch.from_numpy(rowptr).cuda(), torch.from_numpy(col).cuda()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--impl", default="both", choices=["both", "ours", "reference"])
    ap.add_argument("--graph", default=None, choices=GRAPHS)
    args = ap.parse_args()
    graphs_ = (args.graph,) if args.graph else GRAPHS

    if args.impl == "reference":
        so = os.path.join(ROOT, "oracle", "_ref", "_spmm_cuda.so")
        if not os.path.exists(so):
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/_spmm_cuda.so not built"}))
            return 0
        torch.ops.load_library(so)
        for g in graphs_:
            rp, cc = load(g, args.scale)
            M, nnz = rp.numel() - 1, cc.numel()
            val = torch.rand(nnz, device="cuda")
            ms = timeit(lambda: torch.ops.dgsparse_spmm.csr2csc(rp, cc, val), args.reps)
            alg = 4 * (M + 1) * 2 + 16 * nnz
            print(json.dumps({"op": "csr2csc", "impl": "reference (cusparseCsr2cscEx2 ALG1, include/cuda/csr2csc.cuh:8-26)",
                              "graph": g, "M": M, "nnz": nnz, "ms": ms, "algorithmic_bytes": alg,
                              "algorithmic_gbs": alg / ms / 1e6}), flush=True)
        return 0

    if args.impl in ("both", "ours"):
        import dgsparse._kernels as K
        for g in graphs_:
            rp, cc = load(g, args.scale)
            M, nnz = rp.numel() - 1, cc.numel()
            val = torch.rand(nnz, device="cuda")
            ms = timeit(lambda: K.csr2csc(rp, cc, val, want_perm=True), args.reps)
            ms_np = timeit(lambda: K.csr2csc(rp, cc, val, want_perm=False), args.reps)
            alg = 4 * (M + 1) * 2 + 16 * nnz + 4 * nnz
            print(json.dumps({"op": "csr2csc", "impl": "ours (dgs_csr2csc, stable radix transpose + exact int32 permutation)",
                              "graph": g, "M": M, "nnz": nnz, "ms": ms, "ms_without_perm_output": ms_np,
                              "algorithmic_bytes": alg, "algorithmic_gbs": alg / ms / 1e6}), flush=True)
            del rp, cc, val
            torch.cuda.empty_cache()
    if args.impl == "both":
        cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--reps", str(args.reps), "--scale", str(args.scale)]
        if args.graph:
            cmd += ["--graph", args.graph]
        subprocess.run(cmd, check=False)
    return 0


if __name__ == "__main__":
    sys.exit(main())
