"""Raw node / edge feature lookup (reference feature_getter.py:25-106).

Tables are always device-resident (`register_buffer=False`, the reference's host-table mode, would
put a PCIe round trip on every gather; the tables are moved to `device` instead).  A missing table
yields zeros of width `dim`; the fused kernels take that as a NULL pointer and skip the read."""
from typing import Optional

import torch
from torch import Tensor, nn
from torch.nn import functional as F

from www2023tiger_b200 import ops
from ._native import f32c, use_kernel


class FeatureGetter(nn.Module):
    n_nodes: int
    n_edges: int
    nfeat_dim: int
    efeat_dim: int
    out_dim: int
    device: torch.device

    def get_node_embeddings(self, nids: Tensor) -> Tensor:
        raise NotImplementedError

    def get_edge_embeddings(self, eids: Tensor) -> Tensor:
        raise NotImplementedError


class NumericalFeature(FeatureGetter):
    def __init__(self, nfeats: Optional[Tensor], efeats: Optional[Tensor], dim: int, *, use_tsfm: bool = False,
                 register_buffer: bool = True, device: torch.device = None):
        super().__init__()
        self.pin_mem = register_buffer
        self.device = device
        self.use_tsfm = use_tsfm
        self.n_nodes = self.n_edges = None
        self.nfeat_dim = self.efeat_dim = None
        for name, table in (('nfeats', nfeats), ('efeats', efeats)):
            if table is not None:
                table = table.float().contiguous()
                if not register_buffer and device is not None:
                    table = table.to(device)
            self.register_buffer(name, table, persistent=False)
        if nfeats is not None:
            self.n_nodes, self.nfeat_dim = nfeats.shape
        if efeats is not None:
            self.n_edges, self.efeat_dim = efeats.shape
        self.out_dim = dim
        if use_tsfm:
            if nfeats is not None:
                self.node_linear = nn.Linear(self.nfeat_dim, dim)
            if efeats is not None:
                self.edge_linear = nn.Linear(self.efeat_dim, dim)
        self.nfeat_dim = self.nfeat_dim if self.nfeat_dim else dim
        self.efeat_dim = self.efeat_dim if self.efeat_dim else dim

    def _lookup(self, table: Optional[Tensor], ids: Tensor, linear) -> Tensor:
        if table is None:
            return torch.zeros(ids.shape, device=ids.device).unsqueeze(-1).expand(*ids.shape, self.out_dim)
        if use_kernel() and table.is_cuda:
            flat = ids.reshape(-1).to(table.device, torch.int64).contiguous()
            x = ops.gather_rows(table, flat)[0].reshape(*ids.shape, table.shape[1])
        else:
            x = F.embedding(ids.to(table.device), table)
        return linear(x) if linear is not None else x

    def get_node_embeddings(self, nids: Tensor) -> Tensor:
        return self._lookup(self.nfeats, nids, self.node_linear if self.use_tsfm and self.nfeats is not None else None)

    def get_edge_embeddings(self, eids: Tensor) -> Tensor:
        return self._lookup(self.efeats, eids, self.edge_linear if self.use_tsfm and self.efeats is not None else None)
