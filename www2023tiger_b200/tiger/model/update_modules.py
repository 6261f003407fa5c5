"""Memory updaters (reference tiger/model/update_modules.py): GRUCell (default) and MergeLayer."""
import torch
from torch import Tensor, nn

from www2023tiger_b200 import ops
from ._native import f32c, use_kernel
from .basic_modules import MergeLayer


class UpdateModule(nn.Module):
    def __init__(self, msg_dim, memory_dim):
        super().__init__()
        self.msg_dim = msg_dim
        self.memory_dim = memory_dim

    def forward(self, mem: Tensor, msg: Tensor, delta_ts: Tensor) -> Tensor:
        raise NotImplementedError


class GRUUpdater(UpdateModule):
    """h' = GRUCell(msg, mem); the tensor-core kernel reads the gate weights pre-split into tf32 head / tail
    planes (ops.GruPack), re-packed whenever a parameter's version counter moves (optimizer steps,
    load_state_dict)."""

    def __init__(self, msg_dim, memory_dim):
        super().__init__(msg_dim, memory_dim)
        self.cell = nn.GRUCell(input_size=self.msg_dim, hidden_size=self.memory_dim)
        self._pack, self._pack_key = None, None

    def packed(self) -> ops.GruPack:
        c = self.cell
        params = (c.weight_ih, c.weight_hh, c.bias_ih, c.bias_hh)
        key = tuple((p.data_ptr(), p._version) for p in params)
        if self._pack is None or key != self._pack_key:
            args = [f32c(p) for p in params]
            if self._pack is None or self._pack.wpack.device != args[0].device:
                self._pack = ops.GruPack(*args)
            else:
                self._pack.refresh(*args)
            self._pack_key = key
        return self._pack

    def forward(self, mem: Tensor, msg: Tensor, delta_ts: Tensor) -> Tensor:
        if use_kernel() and msg.is_cuda:
            return ops.gru_update(self.packed(), node_ids=None, x_table=f32c(msg), h_table=f32c(mem),
                                  n_rows=msg.shape[0])
        return self.cell(msg, mem)


class MergeUpdater(UpdateModule):
    def __init__(self, msg_dim, memory_dim):
        super().__init__(msg_dim, memory_dim)
        self.fn = MergeLayer(msg_dim, memory_dim, memory_dim, memory_dim)

    def forward(self, mem: Tensor, msg: Tensor, delta_ts: Tensor) -> Tensor:
        return self.fn(msg, mem)


class IdentityUpdater(UpdateModule):
    def forward(self, mem: Tensor, msg: Tensor, delta_ts: Tensor) -> Tensor:
        return mem
