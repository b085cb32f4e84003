#!/usr/bin/env python
"""Per-call latency of torch.ops.dgsparse_spmm.spmm_sum on a Cora-sized graph (2 708 nodes, 10 556 edges), measured the
way the reference's own benchmark does (benchmark/bench_spmm_time.py:58-67: 10 warm-up calls, 100 calls, wall clock
around a synchronize), forward and forward+backward, feat 16 / 64 / 128, for three implementations of the SAME op
boundary, each in a process of its own (they register the same TORCH_LIBRARY namespace):

  compiled   dgsparse/_spmm_cuda.so  (csrc/torch_ops.cpp over the C ABI)            — the product
  python     dgsparse/_ops.py registration over ctypes (DGSPARSE_PY_OPS=1)          — the fallback
  reference  oracle/_ref/_spmm_cuda.so: the reference's own extension, unmodified, built for sm_100a

    python tools/bench_call_latency.py            # prints one JSON line per (impl, feat)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import json, os, sys, time
import numpy as np, torch
root, impl = sys.argv[1], sys.argv[2]
sys.path[:0] = [root, root + "/dgsparse-lib_b200"]
from tools import graphs
rowptr, col = graphs.random_csr(2708, 2708, 10556, 7, empty_frac=0.0)
dev = torch.device("cuda", 0)
rp, cc = torch.from_numpy(rowptr).to(dev), torch.from_numpy(col).to(dev)
nnz = cc.numel()
if impl == "reference":
    torch.ops.load_library(os.path.join(root, "oracle", "_ref", "_spmm_cuda.so"))
    idx = torch.arange(nnz, device=dev, dtype=torch.float)
    colptr, row, perm = torch.ops.dgsparse_spmm.csr2csc(rp, cc, idx)
    perm = perm.to(torch.int)
else:
    import dgsparse
    assert dgsparse.ops_backend.startswith(impl), dgsparse.ops_backend
    colptr, row, perm = torch.ops.dgsparse_spmm.csr2csc_perm(rp, cc, 2708)
for feat in (16, 64, 128):
    val = torch.rand(nnz, device=dev, requires_grad=True)
    X = torch.rand(2708, feat, device=dev, requires_grad=True)
    f = lambda: torch.ops.dgsparse_spmm.spmm_sum(rp, cc, val, colptr, row, perm, X, True, 0)
    res = {"impl": impl, "feat": feat}
    for name, body in (("forward_us", lambda: f()), ("forward_backward_us", lambda: f().sum().backward())):
        for _ in range(10):
            body()
        torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(100):
            body()
        torch.cuda.synchronize()
        res[name] = (time.time() - t0) / 100 * 1e6
    print(json.dumps(res), flush=True)
"""


def main():
    for impl, env in (("compiled", {}), ("python", {"DGSPARSE_PY_OPS": "1"}), ("reference", {})):
        if impl == "reference" and not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "_spmm_cuda.so")):
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/_spmm_cuda.so not built"}))
            continue
        e = dict(os.environ)
        e.pop("DGSPARSE_PY_OPS", None)
        e.update(env)
        r = subprocess.run([sys.executable, "-c", WORKER, ROOT, impl], capture_output=True, text=True, env=e, timeout=600)
        sys.stdout.write(r.stdout)
        if r.returncode != 0:
            print(json.dumps({"impl": impl, "error": r.stderr[-500:]}))


if __name__ == "__main__":
    main()
