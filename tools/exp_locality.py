#!/usr/bin/env python
"""Experiment (GPU): what the row-segment SpMM does with LOCALITY.  The headline generator draws columns uniformly — the worst
case for L1 and L2 and the only case round 1 measured.  Same M / nnz / degree law, columns 85 % inside the row's community:
    python tools/exp_locality.py            # ms per call, feat 64, for uniform and for communities of 1024 / 4096 / 16384 nodes
Run under `ncu --metrics l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,dram__bytes_read.sum -k regex:spmm_rowseg`
for the hit rates."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dgsparse-lib_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

import dgsparse._kernels as K  # noqa: E402
import dgsparse._lib as L  # noqa: E402
from tools import graphs  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    cases = [("uniform", lambda: graphs.reddit_like(1.0))] + [
        (f"communities_{s}", (lambda s=s: graphs.reddit_like_communities(1.0, comm_size=s))) for s in (1024, 4096, 16384)]
    for name, gen in cases:
        rowptr, col = gen()
        M, nnz = rowptr.size - 1, col.size
        row = np.repeat(np.arange(M), np.diff(rowptr))
        rp, cc = torch.from_numpy(rowptr).to(dev), torch.from_numpy(col).to(dev)
        val = torch.rand(nnz, device=dev)
        B = torch.rand(M, N, device=dev)
        out = torch.empty(M, N, device=dev)
        S = int(name.split("_")[1]) if "_" in name else 0
        intra = float(np.mean((row // S) == (col // S))) if S else None
        for affine in (0,):        # (round 2 also ran an SM-affine block order here: no gain, tools/dead_ends/README.md)
            for _ in range(2):
                K.spmm(rp, cc, val, B, L.SUM, L.MUL, out=out)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                K.spmm(rp, cc, val, B, L.SUM, L.MUL, out=out)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(json.dumps({"graph": name, "N": N, "nnz": nnz, "intra_community_fraction": intra, "sm_affine": affine, "ms": ms,
                              "gflops": 2.0 * nnz * N / ms / 1e6, "gather_TBps": nnz * N * 4 / (ms * 1e-3) / 1e12}), flush=True)
        del rp, cc, val, B, out
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
