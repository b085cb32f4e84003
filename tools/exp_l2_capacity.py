#!/usr/bin/env python
"""Experiment (GPU): how large may the gathered B panel be before the row gather of the SpMM falls out of the L2?
reddit-like graphs scaled to K = 233k .. 700k rows (same degree law, uniform columns), feat 64: B = 60 .. 180 MB.
Prints ms and ns per nonzero; run under `ncu --metrics dram__bytes_read.sum -k regex:spmm_rowseg` for the DRAM bytes.
    python tools/exp_l2_capacity.py [scale ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dgsparse-lib_b200")]
import torch  # noqa: E402

import dgsparse._kernels as K  # noqa: E402
import dgsparse._lib as L  # noqa: E402
from tools import graphs  # noqa: E402


def main():
    scales = [float(x) for x in sys.argv[1:]] or [1.0, 1.25, 1.5, 1.75, 2.0, 2.5, 3.0]
    dev = torch.device("cuda", 0)
    for sc in scales:
        rowptr, col = graphs.reddit_like(sc)
        M, nnz = rowptr.size - 1, col.size
        rp, cc = torch.from_numpy(rowptr).to(dev), torch.from_numpy(col).to(dev)
        val = torch.rand(nnz, device=dev)
        B = torch.rand(M, 64, device=dev)
        out = torch.empty(M, 64, device=dev)
        for _ in range(2):
            K.spmm(rp, cc, val, B, L.SUM, L.MUL, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            K.spmm(rp, cc, val, B, L.SUM, L.MUL, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(json.dumps({"scale": sc, "K": M, "nnz": nnz, "B_MB": M * 256 / 1e6, "ms": ms, "ns_per_nnz": ms * 1e6 / nnz,
                          "gather_TBps": nnz * 256 / (ms * 1e-3) / 1e12}), flush=True)
        del rp, cc, val, B, out
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
