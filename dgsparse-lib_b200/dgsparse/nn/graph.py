"""edge_index -> CSR helpers in plain torch ops (the reference uses torch_sparse.SparseTensor for this,
dgsparse/nn/gcnconv.py:37-49, ginconv.py:43-48; that dependency is dropped)."""
import torch


def csr_from_edge_index(edge_index: torch.Tensor, num_nodes: int, value: torch.Tensor = None):
    """(rowptr int32[num_nodes+1], col int32[nnz], value float32[nnz]) sorted by (row, col), duplicates kept —
    the order torch_sparse.SparseTensor(row=, col=).csr() yields.  value defaults to ones."""
    row, col = edge_index[0].long(), edge_index[1].long()
    if value is None:
        value = torch.ones(row.numel(), dtype=torch.float32, device=row.device)
    perm = torch.argsort(row * num_nodes + col, stable=True)
    row, col, value = row[perm], col[perm], value[perm].to(torch.float32)
    rowptr = torch.zeros(num_nodes + 1, dtype=torch.int64, device=row.device)
    rowptr[1:] = torch.cumsum(torch.bincount(row, minlength=num_nodes), 0)
    return rowptr.to(torch.int32), col.to(torch.int32), value
