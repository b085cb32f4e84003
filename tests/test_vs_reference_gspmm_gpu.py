"""GPU parity against the REFERENCE'S OWN gspmm-fp module (oracle/_ref/spmm.so = src/gspmm-fp/gspmm.{cu,cc} compiled
unmodified for sm_100a by oracle/build_ref_gspmm.sh): all 16 u_<compute>_e_<reduce> ops and the 4 copy_u_<reduce> ops,
ours (dgsparse.gspmm, the same entry-point names) vs theirs on the same CUDA tensors.

max / min are exact; sum / mean within 1e-5 relative (+1e-6 of the output scale: summation order differs).
Feature widths are multiples of 32: the reference dispatches two kernels for k < 32 (missing `else`).
Skipped when the reference build is absent (it cannot be rebuilt on the GPU box)."""
import numpy as np
import pytest
import torch

from util import assert_close_f32, spmm_absref

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def REF(oracle):
    m = oracle.ref_gspmm_module()
    if m is None:
        pytest.skip("oracle/_ref/spmm.so not built (oracle/build_ref_gspmm.sh)")
    return m


@pytest.mark.parametrize("N", [32, 64, 128])
@pytest.mark.parametrize("compute", ["add", "sub", "mul", "div"])
@pytest.mark.parametrize("reduce", ["sum", "max", "min", "mean"])
def test_u_op_e_reduce_same_as_reference(REF, oracle, graphs, compute, reduce, N):
    import dgsparse.gspmm as G
    rowptr, col, (M, Kc) = graphs.load_fixture("p2p-Gnutella31")     # 46 199 empty rows, max degree 78
    val = graphs.uniform(col.size, 3, 0.5, 1.5)
    B = graphs.uniform(Kc * N, 4, -1.0, 1.0).reshape(Kc, N)
    rp, cc, vv, Bd = dev(rowptr), dev(col), dev(val), dev(B)
    ours = getattr(G, f"u_{compute}_e_{reduce}")(rp, cc, vv, Bd)
    theirs = REF.GSpMM_u_e(rp, cc, vv, Bd, getattr(REF.REDUCEOP, reduce.upper()), getattr(REF.COMPUTEOP, compute.upper()))
    torch.cuda.synchronize()
    if reduce in ("max", "min"):
        assert torch.equal(ours, theirs)
    else:
        assert_close_f32(ours.cpu().numpy(), theirs.cpu().numpy(), what=f"u_{compute}_e_{reduce} N={N}",
                         absref=spmm_absref(oracle, rowptr, col, val, B, reduce, compute))


@pytest.mark.parametrize("reduce", ["sum", "max", "min", "mean"])
def test_copy_u_same_as_reference(REF, oracle, graphs, reduce):
    import dgsparse.gspmm as G
    M, Kc, N = 6000, 5000, 64
    rowptr, col = graphs.random_csr(M, Kc, 200000, 17, empty_frac=0.3, hub=2)
    B = graphs.uniform(Kc * N, 5, -1.0, 1.0).reshape(Kc, N)
    rp, cc, Bd = dev(rowptr), dev(col), dev(B)
    ours = getattr(G, f"copy_u_{reduce}")(rp, cc, Bd)
    theirs = REF.GSpMM_u(rp, cc, Bd, getattr(REF.REDUCEOP, reduce.upper()))
    torch.cuda.synchronize()
    if reduce in ("max", "min"):
        assert torch.equal(ours, theirs)
    else:
        assert_close_f32(ours.cpu().numpy(), theirs.cpu().numpy(), what=f"copy_u_{reduce}",
                         absref=spmm_absref(oracle, rowptr, col, None, B, reduce))
