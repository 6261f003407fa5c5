"""Message transform + last-message aggregation (reference tiger/model/message_modules.py)."""
from typing import Optional, Tuple

import torch
from torch import Tensor, nn

from www2023tiger_b200 import ops
from ._native import ErrFlags, f32c, require_cuda, use_kernel


class MessageFunction(nn.Module):
    def __init__(self, raw_msg_dim: int, out_msg_dim: Optional[int] = None):
        super().__init__()
        self.input_size = raw_msg_dim
        self.output_size = out_msg_dim

    def forward(self, raw_messages: Tensor) -> Tensor:
        raise NotImplementedError


class IdentityMessageFunction(MessageFunction):
    def __init__(self, raw_msg_dim: int, *args, **kwargs):
        super().__init__(raw_msg_dim, raw_msg_dim)

    def forward(self, raw_messages: Tensor) -> Tensor:
        return raw_messages


def _linear(layer: nn.Linear, x: Tensor, relu: bool = False) -> Tensor:
    if use_kernel() and x.is_cuda:
        out = torch.empty(x.shape[0], layer.out_features, device=x.device)
        return ops.sgemm_nt(f32c(x), f32c(layer.weight), f32c(layer.bias), out, relu=relu)
    y = layer(x)
    return torch.relu(y) if relu else y


class LinearMessageFunction(MessageFunction):
    def __init__(self, raw_msg_dim: int, out_msg_dim: Optional[int] = None, dropout: float = 0.0):
        out_msg_dim = raw_msg_dim if out_msg_dim is None else out_msg_dim
        super().__init__(raw_msg_dim, out_msg_dim)
        self.fn = nn.Sequential(nn.Dropout(dropout), nn.Linear(raw_msg_dim, out_msg_dim))

    def forward(self, raw_messages: Tensor) -> Tensor:
        return _linear(self.fn[1], self.fn[0](raw_messages))


class MLPMessageFunction(MessageFunction):
    def __init__(self, raw_msg_dim: int, out_msg_dim: Optional[int] = None, dropout: float = 0.0):
        out_msg_dim = raw_msg_dim if out_msg_dim is None else out_msg_dim
        super().__init__(raw_msg_dim, out_msg_dim)
        self.hidden_size = self.output_size // 2
        self.fn = nn.Sequential(nn.Dropout(dropout), nn.Linear(raw_msg_dim, self.hidden_size), nn.ReLU(),
                                nn.Dropout(dropout), nn.Linear(self.hidden_size, self.output_size))

    def forward(self, raw_messages: Tensor) -> Tensor:
        h = _linear(self.fn[1], self.fn[0](raw_messages), relu=True)
        return _linear(self.fn[4], self.fn[3](h))


class MessageAggregatorNoGrad(nn.Module):
    def __init__(self, raw_feat_getter, time_encoder):
        super().__init__()
        self.raw_feat_getter = raw_feat_getter
        self.time_encoder = time_encoder


class LastMessageAggregatorNoGradLastOnly(MessageAggregatorNoGrad):
    """The stored row already is the aggregated (last) message: gather rows + timestamps and check that
    no message precedes the memory state it is applied to (message_modules.py:150-160)."""

    def forward(self, node_ids: Tensor, prev_ts: Tensor, node_msg: Tuple[Tensor, Tensor]) -> Tuple[Tensor, Tensor]:
        vals, ts_table = node_msg
        require_cuda(vals, 'LastMessageAggregatorNoGradLastOnly')
        msgs, ts = ops.gather_rows(vals, node_ids.contiguous(), ts_table)
        if bool((prev_ts > ts).any()):
            raise ValueError('Messages happened later than memory updating.')
        return msgs, ts


class LastMessageAggregator(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError('only the LastOnly aggregator (msg_last_only=True) is implemented')


class LastMessageAggregatorNoGrad(LastMessageAggregator):
    pass
