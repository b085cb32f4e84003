"""dgsparse.nn — GNN layers over the dgsparse SpMM ops (the reference's dgsparse/nn, whose package file is
misnamed `__init.py` so that it exports nothing; fixed here)."""
from .gcnconv import GCN, GCNConv, gcn_norm_from_edge_index, get_gcn_dcsr_from_edge_index
from .ginconv import GIN, GINConv
from .graph import csr_from_edge_index

__all__ = ["GCN", "GCNConv", "GIN", "GINConv", "gcn_norm_from_edge_index", "get_gcn_dcsr_from_edge_index",
           "csr_from_edge_index"]
