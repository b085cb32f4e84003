#!/usr/bin/env python
"""Kernel-map construction: dgs_kmap_build behind dgsparse.sparse_mapping.build_kernel_map against the
reference's own sparse_mapping (src/cuda/sparse_mapping.cu:20-161, built unmodified into oracle/_ref/_ref_spconv.so), same
voxel set, same timing loop.  One JSON line per case.

    python tools/bench_kmap.py [--reps 20] [--voxels 100000]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dgsparse-lib_b200")]


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


def surface_voxels(n, seed):
    """A LiDAR-like occupancy: points on a few random planes / shells inside a 400^3 grid, deduplicated (about 10 % tap
    occupancy for a 3^3 kernel, like the MinkUNet fixtures)."""
    rng = np.random.default_rng(seed)
    pts = []
    while sum(p.shape[0] for p in pts) < n * 2:
        o, u, v = rng.uniform(50, 350, 3), rng.normal(size=3), rng.normal(size=3)
        u /= np.linalg.norm(u); v -= u * (u @ v); v /= np.linalg.norm(v)
        st = rng.uniform(-120, 120, (n // 4, 2))
        pts.append(o + st[:, :1] * u + st[:, 1:] * v + rng.normal(scale=0.4, size=(n // 4, 3)))
    p = np.clip(np.rint(np.concatenate(pts)), 0, 399).astype(np.int32)
    c = np.unique(np.concatenate([np.zeros((p.shape[0], 1), np.int32), p], 1), axis=0)
    return c[rng.permutation(c.shape[0])[:n]].copy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--voxels", type=int, default=100000)
    args = ap.parse_args()
    from dgsparse.sparse_mapping import build_kernel_map
    from oracle import oracle
    REF = oracle.ref_spconv_module()
    in_c = surface_voxels(args.voxels, 1)
    d_in = torch.from_numpy(in_c).cuda()
    n = in_c.shape[0]
    for ks, stride, sep in ((3, 1, True), (3, 1, False), (2, 2, False), (5, 1, True)):
        k_vol = ks ** 3
        km = build_kernel_map(d_in, ks, stride, separate_mid=sep)
        ours = timeit(lambda: build_kernel_map(d_in, ks, stride, separate_mid=sep), args.reps)
        line = {"op": "kernel_map", "voxels": n, "kernel": ks, "stride": stride, "separate_mid": sep,
                "pairs": int(km.kpos[-1].item()), "out_voxels": int(km.out_nnz), "ours_ms": ours,
                "ours": "dgsparse.sparse_mapping.build_kernel_map: output voxels, maps, counts AND the device->host read of the pair count"}
        if stride == 1:
            # like for like with the reference call below: caller-side allocations + the C-ABI call, nothing read back
            from dgsparse._lib import lib, ptr, stream_of

            def core():
                maps = torch.empty(2, k_vol * n, dtype=torch.int32, device="cuda")
                counts = torch.zeros(3 * k_vol + 2, dtype=torch.int32, device="cuda")
                ws = torch.empty(lib.dgs_kmap_workspace_bytes(n, n, k_vol), dtype=torch.uint8, device="cuda")
                lib.dgs_kmap_build_ex(n, ptr(d_in), n, ptr(d_in), ks, ks, ks, 1, 1, 1, 0, 0, 0, 1, 128, int(sep), ptr(maps[0]), ptr(maps[1]),
                                      ptr(counts[:k_vol]), ptr(counts[k_vol:2 * k_vol + 1]), ptr(counts[2 * k_vol + 1:]), ptr(ws),
                                      ws.numel(), stream_of(d_in))
            line["ours_c_abi_ms"] = timeit(core, args.reps)
        if REF is not None:
            i3 = lambda v: torch.tensor([v, v, v], dtype=torch.int32, device="cuda")
            pad, lo, hi = i3(0), i3(0), i3(400)

            def ref():
                m = torch.full((k_vol * n,), -1, dtype=torch.int32, device="cuda")
                knnz = torch.zeros(k_vol, dtype=torch.int32, device="cuda")
                kpos = torch.zeros(k_vol + 1, dtype=torch.int32, device="cuda")
                qkpos = torch.zeros(k_vol + 1, dtype=torch.int32, device="cuda")
                return REF.sparse_mapping(d_in, 1, ks, ks, ks, k_vol, 4, 4, stride, stride, stride, 1, 1, 1, pad, lo, hi, m, knnz, kpos,
                                          qkpos, sep)
            try:
                line["reference_ms"] = timeit(ref, args.reps)
                line["speedup"] = line["reference_ms"] / line.get("ours_c_abi_ms", ours)
                line["reference"] = "sparse_mapping (src/cuda/sparse_mapping.cu:20-161), incl. its caller-side map / counter allocation"
            except Exception as ex:   # the reference asserts on some configurations
                line["reference_error"] = repr(ex)[:200]
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
