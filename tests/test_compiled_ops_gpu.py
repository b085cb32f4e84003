"""GPU: the compiled op library (dgsparse/_spmm_cuda.so, csrc/torch_ops.cpp) against the Python registration of the same
ops (dgsparse/_ops.py, DGSPARSE_PY_OPS=1) — same kernels underneath, so forward AND both gradients must be bit-identical
for all four reduce ops; plus csr2csc / csr2csc_perm / sddmm ops."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import sys
import numpy as np, torch
sys.path[:0] = [%(root)r, %(root)r + "/dgsparse-lib_b200"]
import dgsparse
from dgsparse import SparseTensor
from tools import graphs
assert dgsparse.ops_backend.startswith(%(want)r), dgsparse.ops_backend
rowptr, col, (M, Kc) = graphs.load_fixture("p2p-Gnutella31")
N = 48
val = graphs.uniform(col.size, 21, 0.5, 1.5)
B = graphs.uniform(Kc * N, 22, -1.0, 1.0).reshape(Kc, N)
gout = graphs.uniform(M * N, 23, -1.0, 1.0).reshape(M, N)
out = {}
for has_value in (True, False):
    for op in ("sum", "max", "min", "mean"):
        v = torch.from_numpy(val).cuda().requires_grad_()
        X = torch.from_numpy(B).cuda().requires_grad_()
        st = SparseTensor(rowptr=torch.from_numpy(rowptr).cuda(), col=torch.from_numpy(col).cuda(), values=v, has_value=has_value)
        y = getattr(dgsparse, "spmm_" + op)(st, X, 0)
        y.backward(torch.from_numpy(gout).cuda())
        k = f"{op}_{int(has_value)}"
        out[k + "_y"], out[k + "_gx"] = y.detach().cpu().numpy(), X.grad.cpu().numpy()
        if has_value:
            out[k + "_gv"] = v.grad.reshape(-1).cpu().numpy()
rp, cc = torch.from_numpy(rowptr).cuda(), torch.from_numpy(col).cuda()
a = torch.ops.dgsparse_spmm.csr2csc(rp, cc, torch.from_numpy(val).cuda())
b = torch.ops.dgsparse_spmm.csr2csc_perm(rp, cc, M)
for i, t in enumerate(a): out[f"csc_{i}"] = t.cpu().numpy()
for i, t in enumerate(b): out[f"cscp_{i}"] = t.cpu().numpy()
X = torch.from_numpy(B).cuda()
out["sddmm_csr"] = torch.ops.dgsparse_spmm.sddmm_csr(rp, cc, X, X).cpu().numpy()
np.savez(sys.argv[1], **out)
"""


def run(tmp, name, env_extra, want):
    dst = str(tmp / (name + ".npz"))
    env = dict(os.environ)
    env.pop("DGSPARSE_PY_OPS", None)
    env.update(env_extra)
    r = subprocess.run([sys.executable, "-c", WORKER % {"root": ROOT, "want": want}, dst], capture_output=True, text=True,
                       timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return dict(np.load(dst))


def test_compiled_ops_identical_to_python_ops(tmp_path):
    c = run(tmp_path, "compiled", {}, "compiled")
    p = run(tmp_path, "python", {"DGSPARSE_PY_OPS": "1"}, "python")
    assert set(c) == set(p)
    for k in sorted(c):
        assert c[k].shape == p[k].shape and np.array_equal(c[k], p[k]), k
    assert np.isfinite(c["sum_1_gv"]).all() and np.abs(c["mean_1_gx"]).sum() > 0
