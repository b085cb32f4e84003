"""GPU: the PyTorch operator boundary — dgsparse.spmm_{sum,max,min,mean}(SparseTensor, dense, algorithm)
forward + backward, modelled on the reference's test/test_spmm.py:8-203 (forward vs torch.sparse.mm,
backward vs autograd through a dense statement of the same op)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _graph(graphs, M, nnz, seed):
    rowptr, col = graphs.random_csr(M, M, nnz, seed, empty_frac=0.15, hub=1)
    return rowptr, col


def _dense_ref(rowptr, col, val, X, reduce):
    """Differentiable dense statement in fp64 (edge-list form)."""
    M = rowptr.numel() - 1
    row = torch.repeat_interleave(torch.arange(M, device=X.device), (rowptr[1:] - rowptr[:-1]).long())
    msg = val[:, None] * X[col.long()]
    out = torch.zeros(M, X.size(1), dtype=X.dtype, device=X.device)
    if reduce in ("sum", "mean"):
        out = out.index_add(0, row, msg)
        if reduce == "mean":
            deg = (rowptr[1:] - rowptr[:-1]).clamp(min=1).to(X.dtype)
            out = out / deg[:, None]
        return out
    idx = row[:, None].expand_as(msg)
    return out.scatter_reduce(0, idx, msg, "amax" if reduce == "max" else "amin", include_self=False)


@pytest.mark.parametrize("reduce", ["sum", "max", "min", "mean"])
@pytest.mark.parametrize("N", [32, 64, 128])
@pytest.mark.parametrize("has_value", [True, False])
def test_forward_backward(graphs, reduce, N, has_value):
    import dgsparse
    from dgsparse import SparseTensor
    M = 2000
    rowptr, col = _graph(graphs, M, 40000, 5)
    rp = torch.from_numpy(rowptr).cuda()
    cc = torch.from_numpy(col).cuda()
    g = torch.Generator("cuda").manual_seed(N)
    val = (torch.rand(col.size, device="cuda", generator=g) + 0.5) if has_value else torch.ones(col.size, device="cuda")
    X = torch.rand(M, N, device="cuda", generator=g)
    W = torch.rand(M, N, device="cuda", generator=g)     # random cotangent instead of .sum()

    tcsr = torch.sparse_csr_tensor(rp.long(), cc.long(), val.clone(), size=(M, M))
    dcsr = SparseTensor.from_torch_sparse_csr_tensor(tcsr.detach(), has_value=has_value, requires_grad=has_value)
    Xd = X.clone().requires_grad_()
    fn = getattr(dgsparse, f"spmm_{reduce}")
    out = fn(dcsr, Xd, 0)
    (out * W).sum().backward()

    v64 = val.double().requires_grad_()
    X64 = X.double().requires_grad_()
    ref = _dense_ref(rp, cc, v64, X64, reduce)
    (ref * W.double()).sum().backward()

    assert torch.allclose(out.double(), ref, rtol=1e-5, atol=1e-6)
    assert torch.allclose(Xd.grad.double(), X64.grad, rtol=1e-5, atol=1e-5)
    if has_value:
        gv = dcsr.storage._values.grad
        assert gv is not None and gv.shape == val.shape
        assert torch.allclose(gv.double(), v64.grad, rtol=1e-5, atol=1e-5)
    # forward against the reference tests' own oracle for sum (test/test_spmm.py:25)
    if reduce == "sum":
        assert torch.allclose(out, torch.sparse.mm(tcsr, X), rtol=1e-5, atol=1e-5)


def test_storage_builds_exact_csc(graphs, oracle):
    from dgsparse import SparseTensor, csr2csc
    rowptr, col, (M, K) = graphs.load_fixture("ca-CondMat")
    st = SparseTensor(rowptr=torch.from_numpy(rowptr).cuda(), col=torch.from_numpy(col).cuda(), has_value=False)
    ref = oracle.csr2csc(rowptr, col, None, ncols=M)
    assert np.array_equal(st.storage.colptr().cpu().numpy(), ref[0])
    assert np.array_equal(st.storage.row().cpu().numpy(), ref[1])
    assert np.array_equal(st.storage.csr2csc().cpu().numpy(), ref[3])
    assert st.storage.csr2csc().dtype == torch.int32
    colptr, row, vals = csr2csc(st)
    assert np.array_equal(colptr.cpu().numpy(), ref[0])


def test_algorithm_argument_is_accepted(graphs):
    import dgsparse
    from dgsparse import SparseTensor
    rowptr, col = _graph(graphs, 300, 3000, 1)
    st = SparseTensor(rowptr=torch.from_numpy(rowptr).cuda(), col=torch.from_numpy(col).cuda(), has_value=False)
    X = torch.rand(300, 16, device="cuda")
    base = dgsparse.spmm_sum(st, X, 0)
    for alg in (1, 2, 3):
        assert torch.equal(dgsparse.spmm_sum(st, X, alg), base)


def test_non_default_stream_and_device_guard(graphs):
    import dgsparse._kernels as K
    rowptr, col = _graph(graphs, 1000, 20000, 2)
    rp, cc = torch.from_numpy(rowptr).cuda(), torch.from_numpy(col).cuda()
    X = torch.rand(1000, 64, device="cuda")
    ref = K.spmm(rp, cc, None, X)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        out = K.spmm(rp, cc, None, X)
    s.synchronize()
    assert torch.equal(out, ref)
