"""Builds the drop-in model the way the reference's init_utils.init_model does (init_utils.py:126-168),
but from this repository's `tiger` package."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'www2023tiger_b200')
if PKG not in sys.path:
    sys.path.insert(0, PKG)          # makes `import tiger` resolve to www2023tiger_b200/tiger

from tiger.model.feature_getter import NumericalFeature      # noqa: E402
from tiger.model.restarters import SeqRestarter, StaticRestarter  # noqa: E402
from tiger.model.tiger import TIGER                          # noqa: E402


def init_model(nfeats, efeats, train_graph, full_graph_num_node, n_edges, device, *, dim, n_layers, n_heads,
               n_neighbors, hit_type, dropout, restarter_type, hist_len, msg_src, upd_src, msg_tsfm_type='id',
               mem_update_type='gru'):
    if nfeats is not None:
        nfeats = torch.as_tensor(nfeats).float()
        dim = nfeats.shape[1] if dim is None else dim
    if efeats is not None:
        efeats = torch.as_tensor(efeats).float()
        dim = efeats.shape[1] if dim is None else dim
    getter = NumericalFeature(nfeats, efeats, dim=dim, register_buffer=True, device=device)
    getter.n_nodes = full_graph_num_node
    getter.n_edges = n_edges
    if restarter_type == 'seq':
        restarter = SeqRestarter(raw_feat_getter=getter, graph=train_graph, hist_len=hist_len, n_head=n_heads,
                                 dropout=dropout)
    elif restarter_type == 'static':
        restarter = StaticRestarter(raw_feat_getter=getter, graph=train_graph)
    else:
        raise NotImplementedError
    model = TIGER(raw_feat_getter=getter, graph=train_graph, restarter=restarter, n_neighbors=n_neighbors,
                  hit_type=hit_type, n_layers=n_layers, n_head=n_heads, dropout=dropout, msg_src=msg_src,
                  upd_src=upd_src, msg_tsfm_type=msg_tsfm_type, mem_update_type=mem_update_type, tgn_mode=True,
                  msg_last_only=True)
    return model.to(device) if device is not None else model


def model_from_golden(g, graph, device, dropout=0.1):
    m = init_model(g.nfeats, g.efeats, graph, g.N, len(g.src), device, dim=g.dim,
                   n_layers=g.n_layers, n_heads=g.n_heads, n_neighbors=g.K, hit_type=g.hit_type, dropout=dropout,
                   restarter_type=g.restarter, hist_len=g.hist_len, msg_src=g.msg_src, upd_src=g.upd_src,
                   msg_tsfm_type=g.tsfm, mem_update_type=g.upd)
    return m


STATE_BUFFER_SUFFIXES = ('memory.vals', 'memory.update_ts', 'memory.active_mask')


def load_golden_weights(model, g):
    """The fixtures hold the reference model's state_dict minus the memory buffers."""
    res = model.load_state_dict({k: v for k, v in g.W.items()}, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    assert all(k.endswith(STATE_BUFFER_SUFFIXES) for k in res.missing_keys), res.missing_keys
    return model
