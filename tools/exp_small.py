#!/usr/bin/env python
"""Latency regime: SpMM / SDDMM on the reference's own small fixture (p2p-Gnutella31: 62 586 rows, 147 892 nnz, 74 % empty rows)
through the C ABI.  Run under `ncu --metrics gpu__time_duration.sum` for the per-kernel split, or alone for the call time."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dgsparse-lib_b200"))


def main():
    import dgsparse._lib as L
    from tools import graphs
    name = sys.argv[1] if len(sys.argv) > 1 else "p2p-Gnutella31"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    rowptr, col, _ = graphs.load_fixture(name)
    M, nnz = rowptr.size - 1, int(col.size)
    rp, cc = torch.from_numpy(rowptr).cuda(), torch.from_numpy(col).cuda()
    vv = torch.rand(nnz, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for N in (32, 64, 128):
        B = torch.rand(M, N, device="cuda")
        C = torch.empty(M, N, device="cuda")
        ws = torch.empty(L.lib.dgs_spmm_workspace_bytes(N, nnz, 0), dtype=torch.uint8, device="cuda")
        f = lambda: L.lib.dgs_spmm_csr(M, N, nnz, rp.data_ptr(), cc.data_ptr(), vv.data_ptr(), B.data_ptr(), N, C.data_ptr(), N,
                                       None, 0, L.SUM, L.MUL, ws.data_ptr(), ws.numel(), st)
        for _ in range(reps):
            f()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps * 20):
            f()
        b.record()
        torch.cuda.synchronize()
        print(f"{name} spmm N={N}: {a.elapsed_time(b) / (reps * 20) * 1e3:.1f} us per call", flush=True)
    for Kd in (32, 64, 128, 256, 512):
        D1, D2 = torch.rand(M, Kd, device="cuda"), torch.rand(M, Kd, device="cuda")
        out = torch.empty(nnz, device="cuda")
        f = lambda: L.lib.dgs_sddmm_csr(M, Kd, nnz, rp.data_ptr(), cc.data_ptr(), D1.data_ptr(), Kd, D2.data_ptr(), Kd, None, 0,
                                        out.data_ptr(), st)
        for _ in range(reps):
            f()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps * 20):
            f()
        b.record()
        torch.cuda.synchronize()
        print(f"{name} sddmm K={Kd}: {a.elapsed_time(b) / (reps * 20) * 1e3:.1f} us per call", flush=True)


if __name__ == "__main__":
    main()
