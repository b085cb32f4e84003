"""GCN layers — mirror of dgsparse/nn/gcnconv.py:10-70 of the reference (same classes, same forward signature
`forward(dcsr, x)`), with the symmetric normalisation done in torch instead of torch_sparse."""
import torch
import torch.nn.functional as F

from .. import SparseTensor, spmm_sum
from .graph import csr_from_edge_index


class GCNConv(torch.nn.Module):  # dgsparse/nn/gcnconv.py:10-19

    def __init__(self, in_size, out_size):
        super().__init__()
        self.W = torch.nn.Linear(in_size, out_size, bias=False)

    def forward(self, dcsr, x):
        x = self.W(x)
        x = spmm_sum(dcsr, x, 0)
        return x


class GCN(torch.nn.Module):  # dgsparse/nn/gcnconv.py:22-34

    def __init__(self, in_size, out_size, hidden_size):
        super().__init__()
        self.conv1 = GCNConv(in_size, hidden_size)
        self.conv2 = GCNConv(hidden_size, out_size)

    def forward(self, dcsr, x):
        x = self.conv1(dcsr, x)
        x = F.relu(x)
        x = self.conv2(dcsr, x)
        return x


def gcn_norm_from_edge_index(edge_index, num_nodes, add_self_loops=True):
    """D^-1/2 (A + I) D^-1/2 as (rowptr, col, value) — dgsparse/nn/gcnconv.py:37-49 (fill_diag replaces any
    existing diagonal entry by one of weight 1; deg is the row sum)."""
    row, col = edge_index[0].long(), edge_index[1].long()
    value = torch.ones(row.numel(), dtype=torch.float32, device=row.device)
    if add_self_loops:
        keep = row != col
        loops = torch.arange(num_nodes, device=row.device)
        row, col = torch.cat([row[keep], loops]), torch.cat([col[keep], loops])
        value = torch.cat([value[keep], torch.ones(num_nodes, dtype=torch.float32, device=row.device)])
    deg = torch.zeros(num_nodes, dtype=torch.float32, device=row.device).index_add_(0, row, value)
    dis = deg.pow(-0.5)
    dis.masked_fill_(dis == float("inf"), 0.0)
    value = value * dis[row] * dis[col]
    return csr_from_edge_index(torch.stack([row, col]), num_nodes, value)


def get_gcn_dcsr_from_edge_index(edge_index, num_nodes):  # dgsparse/nn/gcnconv.py:52-70
    rowptr, col, value = gcn_norm_from_edge_index(edge_index, num_nodes)
    return SparseTensor(row=None, rowptr=rowptr, col=col, values=value.requires_grad_(), has_value=True)
