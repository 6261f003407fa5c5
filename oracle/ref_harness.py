"""Drives the UNMODIFIED reference (vendored by tools/vendor_ref.py into the git-ignored oracle/_ref/) on the
host CPU: the `cpu_baseline` / `--impl reference` arm of bench.py with `kind: "reference"`, and the in-bench
parity check.

TEST / MEASUREMENT INFRASTRUCTURE ONLY - nothing under www2023tiger_b200/ imports this module.

The recipe is SURVEY.md Appendix C.  Everything that computes is the reference's own code:
`InteractionData`, `Graph.from_data`, `GraphCollator`, `init_utils.init_model`, `TIGER.restart`,
`TIGER.contrast_learning`, `TIGER.contrast_and_mutual_learning`, `torch.optim.Adam`.  The two documented
workarounds are the `torch_scatter.scatter_max` shim (the package is neither installed nor vendored by the
reference; CPU tie rule) and by-passing `load_jodie_data` (needs the JODIE csv files, and
`random.sample(set)` fails on Python >= 3.11) by constructing `InteractionData` directly.

The eval step is the loop body of `eval_edge_prediction` (tiger/eval_utils.py:29-46), the training step the
loop body of `train_self_supervised.py:143-174`.
"""
import hashlib
import json
import os
import sys
import time
from typing import Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, '_ref')


def available() -> Optional[str]:
    """None when oracle/_ref holds an intact vendored reference, else the reason it cannot be used."""
    mf = os.path.join(REF_DIR, 'MANIFEST.json')
    if not os.path.exists(mf):
        return 'oracle/_ref is missing (run tools/vendor_ref.py where /root/reference exists)'
    files = json.load(open(mf))['files']
    for rel, digest in files.items():
        p = os.path.join(REF_DIR, rel)
        if not os.path.exists(p) or hashlib.sha256(open(p, 'rb').read()).hexdigest() != digest:
            return f'oracle/_ref/{rel} does not match its manifest entry'
    return None


_mods = None


def import_reference():
    """Imports the vendored reference as the top-level packages it expects to be (`tiger`, `init_utils`)."""
    global _mods
    if _mods is not None:
        return _mods
    why = available()
    if why:
        raise RuntimeError(why)
    if 'tiger' in sys.modules and not os.path.abspath(sys.modules['tiger'].__file__).startswith(REF_DIR):
        raise RuntimeError('another top-level `tiger` package is already imported')
    sys.path.insert(0, REF_DIR)
    import init_utils                                               # noqa: E402
    from tiger.data.data_loader import GraphCollator, InteractionData   # noqa: E402
    from tiger.data.graph import Graph                              # noqa: E402
    _mods = dict(init_utils=init_utils, GraphCollator=GraphCollator, InteractionData=InteractionData, Graph=Graph)
    return _mods


class ReferenceRunner:
    """The reference's model + collator on a prefix `[0, n_graph_events)` of a synthetic stream.

    `n_graph_events` bounds the cost of `Graph.from_data` (Python loops: ~15 us per event); queries only look
    backwards in time, so a graph over the events up to the end of the replayed window gives exactly the results of
    the full graph."""

    def __init__(self, st, neg, *, restarter: str, msg_src: str, upd_src: str, n_graph_events: Optional[int] = None,
                 n_neighbors: int = 10, n_heads: int = 2, hist_len: int = 40, batch: int = 200, seed: int = 0,
                 threads: Optional[int] = None, dropout: float = 0.1, hit_type: str = 'bin', n_layers: int = 1,
                 msg_tsfm_type: str = 'id', mem_update_type: str = 'gru'):
        import torch
        m = import_reference()
        self.torch = torch
        self.cores = threads or (os.cpu_count() or 1)
        torch.set_num_threads(self.cores)
        E = st.n_events if n_graph_events is None else min(n_graph_events, st.n_events)
        self.E, self.B, self.st, self.restarter = E, batch, st, restarter
        t0 = time.perf_counter()
        self.data = m['InteractionData'](st.src[:E], st.dst[:E], st.ts[:E], st.eids[:E], st.labels[:E], seed=seed,
                                         eval=True, neg_dst=np.asarray(neg[:E]))
        self.graph = m['Graph'].from_data(self.data, strategy='recent_edges', seed=seed, max_node_id=st.n_nodes - 1)
        self.graph_build_s = time.perf_counter() - t0
        self.collator = m['GraphCollator'](self.graph, n_neighbors, n_layers, restarter=restarter, hist_len=hist_len)
        torch.manual_seed(seed)
        np.random.seed(seed)
        efeats = st.efeats[:E + 1] if st.efeats is not None else None
        # the model's row counts follow the FULL stream (n_nodes rows per table, n_edges = all events)
        self.model = m['init_utils'].init_model(
            st.nfeats, efeats, self.graph, self.graph, _Len(st.n_events), torch.device('cpu'), feature_as_buffer=True,
            dim=st.shape.dim, n_layers=n_layers, n_heads=n_heads, n_neighbors=n_neighbors, hit_type=hit_type,
            dropout=dropout, restarter_type=restarter, hist_len=hist_len, msg_src=msg_src, upd_src=upd_src,
            msg_tsfm_type=msg_tsfm_type, mem_update_type=mem_update_type)
        self.uptodate = set()
        self.optimizer = None
        self.reset()

    # ---- state ----
    def reset(self, train: bool = False):
        self.model.train(train)
        self.model.reset()
        self.uptodate = set()

    def state_dict(self):
        """Parameters only (memories / feature buffers excluded), detached clones under the reference's names."""
        skip = ('memory.vals', 'memory.update_ts', 'memory.active_mask', 'raw_feat_getter')
        return {k: v.detach().clone() for k, v in self.model.state_dict().items()
                if not any(s in k for s in skip)}

    # ---- one batch ----
    def collate(self, lo: int):
        batch = [self.data[i] for i in range(lo, lo + self.B)]
        return self.collator(batch)

    def _lazy_restart(self, ts, cg):
        torch = self.torch
        involved = cg.np_computation_graph_nodes
        restart_nodes = set(involved) - set(self.uptodate)
        r_nids = torch.tensor(list(restart_nodes)).long()
        self.model.restart(r_nids, torch.full((len(r_nids),), ts.min().item()))
        self.uptodate.update(restart_nodes)
        return r_nids

    def eval_step(self, lo: int, lazy_restart: bool = True):
        """Loop body of eval_edge_prediction (tiger/eval_utils.py:29-46) on events [lo, lo + B)."""
        torch = self.torch
        src, dst, neg, ts, eids, _, cg = self.collate(lo)
        with torch.no_grad():
            if lazy_restart:
                self._lazy_restart(ts, cg)
            out = self.model.contrast_learning(src, dst, neg, ts, eids, cg)
        return out, cg

    def train_step(self, lo: int, *, lr: float = 1e-4, mutual_coef: float = 1.0, lazy_restart: bool = True):
        """Loop body of train_self_supervised.py:143-174 (lazy-restart mode on, as the DDP driver always runs)."""
        torch = self.torch
        if self.optimizer is None:
            self.optimizer = torch.optim.Adam(self.model.parameters(), lr=lr)
        src, dst, neg, ts, eids, _, cg = self.collate(lo)
        self.optimizer.zero_grad()
        if lazy_restart:
            self._lazy_restart(ts, cg)
        contrast_loss, mutual_loss = self.model.contrast_and_mutual_learning(src, dst, neg, ts, eids, cg)
        loss = contrast_loss + mutual_coef * mutual_loss
        loss.backward()
        self.optimizer.step()
        return float(contrast_loss.item()), float(mutual_loss.item())


class _Len:
    """`init_model` only takes `len(full_data)` (init_utils.py:142)."""

    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n
