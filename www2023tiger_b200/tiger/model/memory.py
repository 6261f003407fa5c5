"""Node memory tables and the last-message store (reference tiger/model/memory.py:12-138).

Buffers keep the reference's names, shapes and persistence, so checkpoints interchange.  The
Python set `nodes_with_messages` of the reference becomes a uint8 flag table `has_msg` on the
device (a `nodes_with_messages` property still materialises the set for code that reads it)."""
import copy
from typing import Optional, Tuple, Union

import numpy as np
import torch
from torch import Tensor, nn

from www2023tiger_b200 import ops
from ._native import ErrFlags, f32c, require_cuda


class Memory(nn.Module):
    def __init__(self, n, dim):
        super().__init__()
        self.n, self.dim = n, dim
        self.register_buffer('vals', torch.zeros(n, dim), persistent=True)
        self.register_buffer('update_ts', torch.zeros(n), persistent=True)
        self.register_buffer('active_mask', torch.zeros(n).bool(), persistent=True)
        self._err = ErrFlags()

    def clone(self):
        # like the reference (memory.py:21-25) the activity flags are not carried over (they start at zero); unlike
        # it, every buffer of the copy lives on the source's device - the kernels store through all three
        other = Memory(self.n, self.dim).to(self.vals.device)
        other.vals.data = self.vals.data.clone()
        other.update_ts.data = self.update_ts.data.clone()
        return other

    @property
    def device(self):
        return self.vals.device

    def clear(self):
        self.vals.zero_()
        self.update_ts.zero_()
        self.active_mask.zero_()

    def get(self, ids: Tensor) -> Tuple[Tensor, Tensor]:
        require_cuda(self.vals, 'Memory.get')
        return ops.gather_rows(self.vals, ids.contiguous(), self.update_ts)

    @torch.no_grad()
    def set(self, ids: Tensor, vals: Tensor, ts: Tensor, skip_check=False):
        require_cuda(self.vals, 'Memory.set')
        ids = ids.contiguous()
        err = None
        if not skip_check:
            if ids.numel() > 1:
                scratch = ops.SelectScratch(self.n, ids.device) if ids.numel() > 2048 else None
                *_, count = ops.select_latest(ids, f32c(ts), scratch, want_unique=False)
                if int(count) != ids.numel():
                    raise ValueError('Duplicate node ids are not allowed.')
            err = self._err.get(self.device)
        ops.scatter_rows(self.vals, ids, f32c(vals), ts_table=self.update_ts, ts=f32c(ts), active=self.active_mask,
                         check=not skip_check, err_flags=err)
        if err is not None:
            self._err.check()                     # 'You are not allowed to modify past memory.'


class MessageStoreNoGradLastOnly(nn.Module):
    def __init__(self, n, dim):
        super().__init__()
        self.n, self.dim = n, dim
        self.register_buffer('node_msg_vals', torch.zeros((n, dim)).float(), persistent=False)
        self.register_buffer('node_msg_ts', torch.zeros(n).float(), persistent=False)
        self.register_buffer('has_msg', torch.zeros(n, dtype=torch.uint8), persistent=False)
        self._err = ErrFlags()

    @property
    def nodes_with_messages(self) -> set:
        return set(torch.nonzero(self.has_msg).flatten().cpu().numpy().tolist())

    @property
    def node_messages(self) -> Tuple[Tensor, Tensor]:
        return self.node_msg_vals, self.node_msg_ts

    def clone(self):
        return copy.deepcopy(self)

    @torch.no_grad()
    def store_events(self, src_ids: Tensor, dst_ids: Tensor, src_prev_ts: Tensor, dst_prev_ts: Tensor,
                     src_vals: Tensor, dst_vals: Tensor, eids: Tensor, ts: Tensor, emb_getter, time_encoder):
        """Build the raw messages [self+nf | other+nf | edge | cos(dt)] of both endpoints, keep the latest
        per node and store them (one fused kernel after the argmax-by-timestamp selection)."""
        require_cuda(self.node_msg_vals, 'MessageStore.store_events')
        dev = self.node_msg_vals.device
        pos = torch.cat([src_ids, dst_ids]).contiguous()
        ts = f32c(ts)
        winner, *_ = ops.select_latest(pos, ts, ops.SelectScratch(self.n, dev) if pos.numel() > 2048 else None,
                                       want_unique=False)
        d = src_vals.shape[1]
        err = self._err.get(dev)
        ops.store_messages_dense(src_ids.contiguous(), dst_ids.contiguous(), eids.contiguous(), ts, winner,
                                 f32c(src_vals), f32c(dst_vals), f32c(src_prev_ts), f32c(dst_prev_ts),
                                 emb_getter.nfeats, emb_getter.efeats, d, self.dim - 3 * d,
                                 f32c(time_encoder.basis_freq), f32c(time_encoder.phase), self.node_msg_vals,
                                 self.node_msg_ts, self.has_msg, err)
        self._err.check()                         # 'Node has unused messages.'

    def get_outdated_node_ids(self, node_ids: Union[Tensor, np.ndarray, None]) -> Tensor:
        """Ids (ascending, LongTensor on the CPU like the reference) of nodes holding a pending message,
        restricted to `node_ids` when given."""
        if node_ids is None:
            return torch.nonzero(self.has_msg).flatten().cpu()
        ids = torch.as_tensor(node_ids).to(self.has_msg.device, torch.int64)
        ids = torch.unique(ids)
        return ids[self.has_msg[ids].bool()].cpu()

    def clear(self, nids: Optional[Tensor] = None):
        """Drop pending-message flags (the reference never zeroes the table rows either: its
        `vals[nids].fill_(0)` writes to a copy, memory.py:137-138)."""
        if nids is None:
            self.has_msg.zero_()
        else:
            self.has_msg[nids.to(self.has_msg.device)] = 0


# Dict-of-lists stores of the reference (memory.py:141-237) are unreachable from its CLI
# (init_utils.py:166 hard-codes msg_last_only=True); the names stay importable.
class MessageStore(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError('only MessageStoreNoGradLastOnly (msg_last_only=True) is implemented')


class MessageStoreNoGrad(MessageStore):
    pass
