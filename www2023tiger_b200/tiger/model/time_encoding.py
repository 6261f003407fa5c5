"""TGAT harmonic time encoding cos(t*w + b) (reference time_encoding.py:6-27)."""
import numpy as np
import torch
from torch import Tensor, nn

from www2023tiger_b200 import ops
from ._native import f32c, use_kernel


class TimeEncode(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.dim = dim
        self.basis_freq = nn.Parameter(torch.from_numpy(1 / 10 ** np.linspace(0, 9, dim)).float())
        self.phase = nn.Parameter(torch.zeros(dim).float())

    def forward(self, ts: Tensor) -> Tensor:
        """ts [n] or [n, len] -> [n, dim] or [n, len, dim]."""
        if use_kernel() and ts.is_cuda:
            return ops.time_encode(f32c(ts), f32c(self.basis_freq), f32c(self.phase))
        return torch.cos(ts.unsqueeze(-1) * self.basis_freq + self.phase)
