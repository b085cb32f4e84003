#!/usr/bin/env python
"""Times the spconv gather-GEMM-scatter on the reference's MinkUNet kernel maps (tests/golden/spconv_fp32_*.npz).

    python tools/bench_spconv.py [--reps 50]

Prints one JSON line per (layer, precision): forward / dX / dW device time (CUDA events, L2 flushed between
repetitions), the algorithmic bytes (maps + each touched input row once + output once + weights) and the
gather/scatter bytes that actually move through L2 (pairs * (c_in + c_out) * 4), and GFLOP/s = 2*pairs*c_in*c_out/t.
Beside every tensor-core line: the REFERENCE'S OWN spconv_fwd_fused / spconv_bwd_fused (src/cuda/spconv_cuda.cu:18-253,
compiled unmodified for sm_100a into oracle/_ref/_ref_spconv.so by oracle/build_ref_spconv.sh) on the same maps and
tensors, same timing loop (`reference_*_ms`; its backward computes dX and dW in one call).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dgsparse-lib_b200"))


def timeit(fn, reps, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--channels", default="", help="extra c_in,c_out pairs on the layer-1 maps, e.g. 128,128;256,256")
    args = ap.parse_args()
    import dgsparse.spconv as S
    from oracle import oracle
    REF = oracle.ref_spconv_module()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    cases = []
    for idx in (0, 1):
        g = np.load(os.path.join(ROOT, "tests", "golden", f"spconv_fp32_{idx}.npz"))
        cases.append((f"minkunet_layer{idx}", g, int(g["c_in"]), int(g["c_out"])))
    for pair in filter(None, args.channels.split(";")):
        ci, co = (int(v) for v in pair.split(","))
        cases.append((f"minkunet_maps_{ci}x{co}", cases[1][1], ci, co))
    for name, g, c_in, c_out in cases:
        in_nnz, out_nnz, k_vol = int(g["in_nnz"]), int(g["out_nnz"]), int(g["k_vol"])
        knnz = torch.from_numpy(g["knnz"].astype(np.int64)).cuda()
        kpos, qkpos, sum_nnz = S.quantize_kpos(knnz)
        imap, omap = torch.from_numpy(g["imap"]).cuda(), torch.from_numpy(g["omap"]).cuda()
        pairs = int(g["imap"].size)
        x = torch.rand(in_nnz, c_in, device="cuda")
        w = torch.rand(k_vol, c_in, c_out, device="cuda")
        go = torch.rand(out_nnz, c_out, device="cuda")
        flop = 2.0 * pairs * c_in * c_out
        alg = 8 * pairs + 4 * in_nnz * c_in + 4 * out_nnz * c_out + 4 * k_vol * c_in * c_out
        for prec in ("fp32", "tf32", "bf16"):
            fwd = timeit(lambda: S.spconv_fwd_fused(x, w, kpos, qkpos, imap, omap, out_nnz, sum_nnz, False, True, precision=prec),
                         args.reps, flush)
            dx = timeit(lambda: S.spconv_bwd_fused(go, x, w, kpos, qkpos, imap, omap, sum_nnz, False, True, need_kernel=False,
                                                   precision=prec), args.reps, flush)
            dw = timeit(lambda: S.spconv_bwd_fused(go, x, w, kpos, qkpos, imap, omap, sum_nnz, False, True, need_in=False,
                                                   precision=prec), args.reps, flush)
            line = {"case": name, "precision": prec, "pairs": pairs, "c_in": c_in, "c_out": c_out,
                    "fwd_ms": fwd, "dx_ms": dx, "dw_ms": dw, "fwd_gflops": flop / fwd / 1e6,
                    "fwd_algorithmic_gbs": alg / fwd / 1e6,
                    "fwd_gather_scatter_gbs": 4.0 * pairs * (c_in + c_out) / fwd / 1e6}
            if REF is not None and prec in ("fp32", "tf32"):
                arch80 = prec == "tf32"
                line["reference_fwd_ms"] = timeit(lambda: REF.spconv_fwd_fused(x, w, kpos, qkpos, imap, omap, out_nnz, sum_nnz, False, arch80),
                                                  args.reps, flush)
                if arch80:   # the reference's backward is tensor-core only (tf32 kernels whatever arch80 says)
                    line["reference_bwd_dx_plus_dw_ms"] = timeit(
                        lambda: REF.spconv_bwd_fused(go, x, w, kpos, qkpos, imap, omap, sum_nnz, False, True), args.reps, flush)
                    line["ours_bwd_dx_plus_dw_ms"] = dx + dw
                line["fwd_speedup_vs_reference"] = line["reference_fwd_ms"] / fwd
            print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
