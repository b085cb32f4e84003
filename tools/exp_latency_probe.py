#!/usr/bin/env python
"""Where do the microseconds of one torch.ops.dgsparse_spmm.spmm_sum call on a Cora-sized graph go?  (round-2 probe)"""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, ROOT + "/dgsparse-lib_b200"]
from tools import graphs
import dgsparse
import dgsparse._lib as L
rowptr, col = graphs.random_csr(2708, 2708, 10556, 7, empty_frac=0.0)
dev = torch.device("cuda", 0)
rp, cc = torch.from_numpy(rowptr).to(dev), torch.from_numpy(col).to(dev)
nnz = cc.numel()
colptr, row, perm = torch.ops.dgsparse_spmm.csr2csc_perm(rp, cc, 2708)


def wall(body, n=200):
    for _ in range(20):
        body()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); a.record()
    for _ in range(n):
        body()
    b.record(); t_issue = time.time() - t0
    torch.cuda.synchronize()
    return {"wall_us": (time.time() - t0) / n * 1e6, "issue_us": t_issue / n * 1e6, "gpu_us": a.elapsed_time(b) / n * 1e3}


for feat in (16, 64, 128):
    val = torch.rand(nnz, device=dev, requires_grad=True)
    X = torch.rand(2708, feat, device=dev, requires_grad=True)
    vd, Xd = val.detach(), X.detach()
    out = torch.empty(2708, feat, device=dev)
    ws = torch.empty(L.lib.dgs_spmm_workspace_bytes(feat, nnz, 0), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    res = {"feat": feat}
    res["op_grad"] = wall(lambda: torch.ops.dgsparse_spmm.spmm_sum(rp, cc, val, colptr, row, perm, X, True, 0))
    res["op_nograd"] = wall(lambda: torch.ops.dgsparse_spmm.spmm_sum(rp, cc, vd, colptr, row, perm, Xd, True, 0))
    for mode in (-1, 0, 1):
        L.lib.dgs_set_option(b"spmm_rowpar", mode)
        res[f"cabi_rowpar{mode}"] = wall(lambda: L.lib.dgs_spmm_csr_k(2708, 2708, feat, nnz, rp.data_ptr(), cc.data_ptr(), vd.data_ptr(), Xd.data_ptr(), feat,
                                                                        out.data_ptr(), feat, None, 0, 0, 2, ws.data_ptr(), ws.numel(), st))
        res[f"cabi_rowpar{mode}"]["path"] = L.lib.dgs_spmm_last_path()
    L.lib.dgs_set_option(b"spmm_rowpar", -1)
    res["empty_alloc"] = wall(lambda: torch.empty(2708, feat, device=dev))
    print(json.dumps({k: ({kk: round(vv, 2) for kk, vv in v.items()} if isinstance(v, dict) else v) for k, v in res.items()}), flush=True)
