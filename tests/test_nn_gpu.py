"""GPU: dgsparse.nn GCN / GIN layers (reference dgsparse/nn, test/test_dgl.py:52-88) against dense torch math,
forward and backward, plus a short training run (the caller side of the SpMM path)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def random_edges(n, e, seed):
    g = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, n, (2, e), generator=g)
    ei = torch.unique(ei, dim=1)                  # the dense reference below cannot represent duplicate edges
    return ei.cuda()


def dense_gcn_adj(ei, n):
    A = torch.zeros(n, n, dtype=torch.float64, device="cuda")
    A[ei[0], ei[1]] = 1.0
    A.fill_diagonal_(1.0)
    d = A.sum(1).pow(-0.5)
    d[torch.isinf(d)] = 0
    return d[:, None] * A * d[None, :]


def test_gcn_matches_dense_forward_backward():
    from dgsparse.nn import GCN, get_gcn_dcsr_from_edge_index
    n, fin, hid, out = 700, 48, 64, 16
    ei = random_edges(n, 6000, 1)
    torch.manual_seed(0)
    model = GCN(fin, out, hid).cuda()
    x = torch.randn(n, fin, device="cuda", requires_grad=True)
    dcsr = get_gcn_dcsr_from_edge_index(ei, n)
    y = model(dcsr, x)
    y.square().sum().backward()
    A = dense_gcn_adj(ei, n)
    W1, W2 = model.conv1.W.weight.detach().double(), model.conv2.W.weight.detach().double()
    x64 = x.detach().double().requires_grad_()
    W1.requires_grad_(); W2.requires_grad_()
    y64 = A @ (F.relu(A @ (x64 @ W1.T)) @ W2.T)
    y64.square().sum().backward()
    assert torch.allclose(y.double(), y64, rtol=1e-4, atol=1e-5)
    assert torch.allclose(x.grad.double(), x64.grad, rtol=1e-3, atol=1e-5)
    assert torch.allclose(model.conv1.W.weight.grad.double(), W1.grad, rtol=1e-3, atol=1e-4)
    assert torch.allclose(model.conv2.W.weight.grad.double(), W2.grad, rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("agg", ["sum", "mean", "max"])
def test_gin_matches_dense(agg):
    from dgsparse.nn import GIN
    n, fin, hid, out = 500, 32, 64, 8
    ei = random_edges(n, 4000, 2)
    torch.manual_seed(1)
    model = GIN(fin, out, hid, aggregator_type=agg, init_eps=0.1, cached=True).cuda()
    x = torch.randn(n, fin, device="cuda")
    y = model(ei, x, n)
    A = torch.zeros(n, n, dtype=torch.float64, device="cuda")
    A[ei[0], ei[1]] = 1.0

    def aggregate(h):
        if agg == "sum":
            return A @ h
        if agg == "mean":
            return (A @ h) / A.sum(1).clamp(min=1)[:, None]
        big = torch.where(A[:, :, None] > 0, h[None, :, :], torch.full_like(h[None, :, :], -1e30))
        m = big.max(1).values
        return torch.where(A.sum(1)[:, None] > 0, m, torch.zeros_like(m))     # empty rows -> 0

    h = x.double()
    for conv in (model.conv1, model.conv2):
        h = (1 + 0.1) * h + aggregate(h)
        h = F.relu(h @ conv.apply_func.weight.detach().double().T + conv.apply_func.bias.detach().double())
    assert torch.allclose(y.double(), h, rtol=1e-4, atol=1e-4)


def test_gcn_trains():
    """Node classification on a planted-partition graph: the loss must fall and accuracy beat chance by far."""
    from dgsparse.nn import GCN, get_gcn_dcsr_from_edge_index
    rng = np.random.default_rng(0)
    n, classes = 1200, 4
    label = rng.integers(0, classes, n)
    src = rng.integers(0, n, 20000)
    same = rng.random(20000) < 0.85
    # intra-class edges with probability 0.85: pick a random node of the same class
    by_class = [np.flatnonzero(label == c) for c in range(classes)]
    dst = np.array([rng.choice(by_class[label[s]]) if k else rng.integers(0, n) for s, k in zip(src, same)])
    ei = torch.tensor(np.stack([np.concatenate([src, dst]), np.concatenate([dst, src])]), device="cuda")
    x = torch.randn(n, 32, device="cuda")
    y = torch.tensor(label, device="cuda")
    torch.manual_seed(0)
    model = GCN(32, classes, 64).cuda()
    dcsr = get_gcn_dcsr_from_edge_index(ei, n)
    opt = torch.optim.Adam(model.parameters(), lr=0.02)
    losses = []
    for _ in range(60):
        opt.zero_grad()
        loss = F.cross_entropy(model(dcsr, x), y)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    acc = float((model(dcsr, x).argmax(1) == y).float().mean())
    assert losses[-1] < 0.5 * losses[0] and acc > 0.8, (losses[0], losses[-1], acc)
