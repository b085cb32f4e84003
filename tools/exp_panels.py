#!/usr/bin/env python
"""Experiment (GPU, round 2 — its 8 / 16-column kernels now live under tools/dead_ends/, so only widths 64 and 32 still run
from this tree; the recorded results are profiles/r02_exp_panels_*.jsonl): SpMM time against the column-panel width
(DGS_SPMM_PANEL, read at library load) on the products-like and reddit-like matrices.   python tools/exp_panels.py [products|reddit] [N] [--once PANEL]   (--once: a single call, for ncu)"""
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
    which = sys.argv[1] if len(sys.argv) > 1 else "products"
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    once = sys.argv[sys.argv.index("--once") + 1] if "--once" in sys.argv else None
    rowptr, col = graphs.products_like(1.0) if which == "products" else graphs.reddit_like(1.0)
    M, nnz = rowptr.size - 1, col.size
    dev = torch.device("cuda", 0)
    rp, cc = torch.from_numpy(rowptr).to(dev), torch.from_numpy(col).to(dev)
    val = torch.rand(nnz, device=dev) + 0.5
    B = torch.rand(M, N, device=dev)
    out = torch.empty(M, N, device=dev)
    cases = [("sum_val", L.SUM, val, False), ("max_val_arg", L.MAX, val, True), ("max_noval", L.MAX, None, False),
             ("mean_noval", L.MEAN, None, False)]
    if once:
        os.environ["DGS_SPMM_PANEL"] = once
        for name, red, v, wa in cases[:2]:
            K.spmm(rp, cc, v, B, red, L.MUL, with_arg=wa)
        torch.cuda.synchronize()
        return
    ref = {}
    for panel in ("64", "32", "16", "8", "auto"):
        if panel == "auto":
            os.environ.pop("DGS_SPMM_PANEL", None)
        else:
            os.environ["DGS_SPMM_PANEL"] = panel
        for name, red, v, wa in cases:
            for _ in range(2):
                r = K.spmm(rp, cc, v, B, red, L.MUL, with_arg=wa, out=out)
            torch.cuda.synchronize()
            L.lib.dgs_profile_enable(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                r = K.spmm(rp, cc, v, B, red, L.MUL, with_arg=wa, out=out)
            e1.record()
            torch.cuda.synchronize()
            import ctypes
            ids, ms = (ctypes.c_int * 64)(), (ctypes.c_float * 64)()
            n = L.lib.dgs_profile_collect(64, ids, ms)
            L.lib.dgs_profile_enable(0)
            k = [ms[i] for i in range(n) if ids[i] == 1]
            f = [ms[i] for i in range(n) if ids[i] == 2]
            o = r[0] if wa else r
            chk = float(o.double().sum().item())
            key = name
            same = None
            if key in ref:
                same = bool(torch.equal(o, ref[key])) if red == L.MAX else float((o - ref[key]).abs().max().item())
            else:
                ref[key] = o.clone()
            print(json.dumps({"graph": which, "N": N, "panel": panel, "op": name, "ms_per_call": e0.elapsed_time(e1) / 5,
                              "kernel_ms": sum(k) / max(1, len(k)), "fixup_ms": sum(f) / max(1, len(f)), "checksum": chk,
                              "vs_panel64": same}), flush=True)


if __name__ == "__main__":
    main()
