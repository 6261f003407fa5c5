"""Batch containers handed from the collator to the model (reference: tiger/data/data_classes.py).

Field names are what the model reads (`layers`, `computation_graph_nodes`,
`np_computation_graph_nodes`, `local_index`, `restart_data.*`, `hit_data`).  The collator of this
package fills them with tensors that already live on the device, so `.to(device)` and
`.pin_memory()` are cheap no-ops in the normal flow; host tensors are still accepted and moved.
"""
from dataclasses import dataclass, fields
from typing import List, Optional, Tuple

import numpy as np
import torch
from torch import Tensor


class RestartData:
    def _tensor_fields(self):
        return [f.name for f in fields(self)]

    def __iter__(self):
        for name in self._tensor_fields():
            yield getattr(self, name)

    def to(self, device: torch.device):
        for name in self._tensor_fields():
            setattr(self, name, getattr(self, name).to(device))

    def pin_memory(self):
        for name in self._tensor_fields():
            t = getattr(self, name)
            if not t.is_cuda:
                setattr(self, name, t.pin_memory())


@dataclass
class SeqRestartData(RestartData):
    index: Tensor            # positions of the selected (latest) occurrence of each unique positive
    nids: Tensor             # unique positive node ids, ascending
    ts: Tensor               # float32 event times of the selected positions
    hist_nids: Tensor        # [n, hist_len]
    anonymized_ids: Tensor   # [n, hist_len]
    hist_eids: Tensor
    hist_ts: Tensor
    hist_dirs: Tensor


@dataclass
class StaticRestartData(RestartData):
    index: Tensor
    nids: Tensor
    ts: Tensor
    prev_ts: Tensor          # [n, 1] float32: time of the last event before ts (0 if none)


@dataclass
class HitData(RestartData):
    src_hits: Tensor
    dst_hits: Tensor
    neg_src_hits: Tensor
    neg_dst_hits: Tensor


class ComputationGraph:
    def __init__(self, tige_data: Tuple[List[Tuple], object], restart_data: Optional[RestartData],
                 hit_data: Optional[HitData], n_nodes: int, local_index: Optional[Tensor] = None):
        """tige_data = [layers, unique involved ids]; `layers[0] = (batch nids, None, None)`,
        `layers[l] = (neigh_nids, neigh_eids, neigh_ts)`.  The involved ids may be a numpy array (as
        in the reference) or a (numpy, device tensor) pair prepared by the device collator."""
        self.n_nodes = n_nodes
        self.layers = tige_data[0]
        nodes = tige_data[1]
        if isinstance(nodes, tuple):
            self.np_computation_graph_nodes, self.computation_graph_nodes = nodes
        else:
            self.np_computation_graph_nodes = np.asarray(nodes)
            self.computation_graph_nodes = torch.from_numpy(self.np_computation_graph_nodes)
        self.restart_data = restart_data
        self.hit_data = hit_data
        if local_index is None:                                  # data_classes.py:163-165
            nodes_t = self.computation_graph_nodes
            local_index = torch.zeros(n_nodes, dtype=torch.long, device=nodes_t.device)
            local_index[nodes_t] = torch.arange(len(nodes_t), device=nodes_t.device)
        self.local_index = local_index

    @property
    def device(self):
        return self.computation_graph_nodes.device

    def to(self, device: torch.device):
        self.layers = [tuple(None if t is None else t.to(device) for t in layer) for layer in self.layers]
        self.computation_graph_nodes = self.computation_graph_nodes.to(device)
        self.local_index = self.local_index.to(device)
        if self.restart_data is not None:
            self.restart_data.to(device)
        if self.hit_data is not None:
            self.hit_data.to(device)
        return self

    def pin_memory(self):
        pin = lambda t: t if (t is None or t.is_cuda) else t.pin_memory()
        self.layers = [tuple(pin(t) for t in layer) for layer in self.layers]
        self.computation_graph_nodes = pin(self.computation_graph_nodes)
        self.local_index = pin(self.local_index)
        if self.restart_data is not None:
            self.restart_data.pin_memory()
        if self.hit_data is not None:
            self.hit_data.pin_memory()
        return self
