#!/usr/bin/env python
"""Runs the REFERENCE'S OWN torch ops (oracle/_ref/_spmm_cuda.so = src/spmm.cpp + src/cuda/spmm_cuda.cu compiled unmodified
for sm_100a by oracle/build_ref_torch_face.sh) on saved inputs, in a process of its own: the reference registers
TORCH_LIBRARY(dgsparse_spmm) (src/spmm.cpp:264), the same namespace our package registers, so the two cannot share a
process.  TEST INFRASTRUCTURE ONLY (tests/test_vs_reference_torch_face_gpu.py).

    python oracle/run_ref_torch_face.py in.npz out.npz

in.npz:  rowptr, col (int32), val, B, gout (float32).  out.npz: for op in sum/max/min/mean: <op>_out, <op>_gval,
<op>_gdense (forward and autograd through torch.ops.dgsparse_spmm.spmm_<op>, driven the way dgsparse/spmm.py:5-106 and
dgsparse/storage.py:159-174 of the reference drive it), plus csc_colptr / csc_row / csc_perm of its csr2csc op.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    src, dst = sys.argv[1], sys.argv[2]
    torch.ops.load_library(os.path.join(HERE, "_ref", "_spmm_cuda.so"))
    d = np.load(src)
    dev = torch.device("cuda", 0)
    rowptr, col = torch.from_numpy(d["rowptr"]).to(dev), torch.from_numpy(d["col"]).to(dev)
    gout = torch.from_numpy(d["gout"]).to(dev)
    out = {}
    # dgsparse/storage.py:159-174: CSC arrays + permutation from the csr2csc op fed with arange as float
    idx = torch.arange(col.numel(), device=dev, dtype=torch.float)
    colptr, row, perm = torch.ops.dgsparse_spmm.csr2csc(rowptr, col, idx)
    perm = perm.to(torch.int)
    out["csc_colptr"], out["csc_row"], out["csc_perm"] = colptr.cpu().numpy(), row.cpu().numpy(), perm.cpu().numpy()
    for op in ("sum", "max", "min", "mean"):
        val = torch.from_numpy(d["val"]).to(dev).requires_grad_()
        B = torch.from_numpy(d["B"]).to(dev).requires_grad_()
        fn = getattr(torch.ops.dgsparse_spmm, "spmm_" + op)
        y = fn(rowptr, col, val, colptr, row, perm, B, True, 0)
        out[op + "_out"] = y.detach().cpu().numpy()
        try:
            y.backward(gout)
            torch.cuda.synchronize()
            out[op + "_gval"] = val.grad.detach().reshape(-1).cpu().numpy()
            out[op + "_gdense"] = B.grad.detach().cpu().numpy()
        except Exception as e:  # the reference's max/min/mean backward is not exercised by every build
            out[op + "_bwd_error"] = np.array(str(e))
    np.savez(dst, **out)


if __name__ == "__main__":
    main()
