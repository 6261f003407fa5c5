"""MergeLayer: fc2(relu(fc1([x1 | x2]))) with xavier-normal weights (reference basic_modules.py:5-19)."""
import torch
from torch import nn

from www2023tiger_b200 import ops
from ._native import f32c, use_kernel


class MergeLayer(nn.Module):
    def __init__(self, dim1, dim2, hidden_size, out_size, dropout=0.):
        super().__init__()
        self.fc1 = nn.Linear(dim1 + dim2, hidden_size)
        self.fc2 = nn.Linear(hidden_size, out_size)
        self.dropout = nn.Dropout(dropout)
        self.act = nn.ReLU()
        nn.init.xavier_normal_(self.fc1.weight)
        nn.init.xavier_normal_(self.fc2.weight)

    def forward(self, x1, x2):
        x = torch.cat([x1, x2], dim=-1)
        if use_kernel(self) and x.is_cuda and x.dim() == 2:
            x = f32c(x)
            hid = torch.empty(x.shape[0], self.fc1.out_features, device=x.device)
            out = torch.empty(x.shape[0], self.fc2.out_features, device=x.device)
            ops.sgemm_nt(x, f32c(self.fc1.weight), f32c(self.fc1.bias), hid, relu=True)   # dropout is identity here
            ops.sgemm_nt(hid, f32c(self.fc2.weight), f32c(self.fc2.bias), out)
            return out
        return self.fc2(self.dropout(self.act(self.fc1(x))))


class MLP(nn.Module):
    """Node-classification decoder (basic_modules.py:22-33); outside the memory path, kept for import parity."""

    def __init__(self, dim, dropout=0.3):
        super().__init__()
        self.fn = nn.Sequential(nn.Linear(dim, 80), nn.ReLU(), nn.Dropout(dropout),
                                nn.Linear(80, 10), nn.ReLU(), nn.Dropout(dropout), nn.Linear(10, 1))

    def forward(self, x):
        return self.fn(x).squeeze(dim=-1)
