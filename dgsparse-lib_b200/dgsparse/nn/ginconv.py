"""GIN layers — mirror of dgsparse/nn/ginconv.py:9-112 of the reference (same constructor arguments and
`forward(edge_index, X, num_nodes)`), without torch_sparse; `cached=True` really caches the CSR."""
import torch
import torch.nn.functional as F

from .. import SparseTensor, spmm_max, spmm_mean, spmm_sum
from .graph import csr_from_edge_index


class GINConv(torch.nn.Module):

    def __init__(self, apply_func=None, aggregator_type="sum", init_eps=0, learn_eps=False, activation=None,
                 cached=False):
        super().__init__()
        self.apply_func = apply_func
        self._aggregator_type = aggregator_type
        self.activation = activation
        self.cached = cached
        self._cached_dcsr = None
        if learn_eps:
            self.eps = torch.nn.Parameter(torch.FloatTensor([init_eps]))
        else:
            self.register_buffer("eps", torch.FloatTensor([init_eps]))

    def forward(self, edge_index, X, num_nodes):
        neigh = self.aggregate_neigh(edge_index, X, num_nodes, 0)
        rst = (1 + self.eps) * X + neigh
        if self.apply_func is not None:
            rst = self.apply_func(rst)
        if self.activation is not None:
            rst = self.activation(rst)
        return rst

    def _dcsr(self, edge_index, num_nodes):
        if self.cached and self._cached_dcsr is not None:
            return self._cached_dcsr
        rowptr, col, value = csr_from_edge_index(edge_index, num_nodes)
        dcsr = SparseTensor(row=None, rowptr=rowptr, col=col, values=value.requires_grad_(), has_value=True)
        if self.cached:
            self._cached_dcsr = dcsr
        return dcsr

    def aggregate_neigh(self, edge_index, X, num_nodes, algorithm):
        dcsr = self._dcsr(edge_index, num_nodes)
        if self._aggregator_type == "max":
            return spmm_max(dcsr, X, algorithm)
        if self._aggregator_type == "mean":
            return spmm_mean(dcsr, X, algorithm)
        return spmm_sum(dcsr, X, algorithm)    # 'sum' and anything else, dgsparse/nn/ginconv.py:61-68


class GIN(torch.nn.Module):

    def __init__(self, in_size, out_size, hidden_size, aggregator_type="sum", init_eps=0, learn_eps=False,
                 activation=F.relu, cached=False):
        super().__init__()
        self.conv1 = GINConv(torch.nn.Linear(in_size, hidden_size), aggregator_type, init_eps, learn_eps, activation,
                             cached)
        self.conv2 = GINConv(torch.nn.Linear(hidden_size, out_size), aggregator_type, init_eps, learn_eps, activation,
                             cached)

    def forward(self, edge_index, X, num_nodes):
        X = self.conv1(edge_index, X, num_nodes)
        X = self.conv2(edge_index, X, num_nodes)
        return X

    @property
    def eps(self):
        return self.conv1.eps
