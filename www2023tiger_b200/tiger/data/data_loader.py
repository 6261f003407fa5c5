"""Batch collation on the device + dataset plumbing (reference: tiger/data/data_loader.py).

`GraphCollator.__call__` keeps the reference's contract - a list of
`(src, dst, neg_dst, ts, eid, label)` tuples in, `(src, dst, neg, ts, eids, labels,
ComputationGraph)` out - but the work it does per batch (3B neighbor lookups, the sorted-unique
involved-node set, `local_index`, the de-duplicated restart history and the hit windows,
data_loader.py:61-168) runs as five kernel launches on the device graph instead of ~1,700
Python-level `np.searchsorted` calls.
"""
import pathlib
import random
from typing import List, Optional, Tuple

import numpy as np
import torch
from torch.utils.data import Dataset, Sampler

from www2023tiger_b200 import ops
from .data_classes import ComputationGraph, HitData, SeqRestartData, StaticRestartData
from .graph import Graph


class ChunkSampler(Sampler):
    """DDP partition of the reference (data_loader.py:17-40): rank r reads the contiguous index range
    [shift + r*L, shift + (r+1)*L), L = n // (world*bs) * bs, shift ~ U[0, n mod (world*bs)] drawn
    from Generator(seed + epoch)."""

    def __init__(self, n: int, rank: int, world_size: int, bs: int, seed: int = 0):
        self.n, self.rank, self.world_size, self.bs, self.seed = n, rank, world_size, bs, seed
        self.epoch = 0

    def bounds(self) -> Tuple[int, int]:
        g = torch.Generator()
        g.manual_seed(self.seed + self.epoch)
        residual = self.n % (self.world_size * self.bs)
        shift = int(torch.randint(0, residual + 1, size=(), generator=g))
        lo = shift + len(self) * self.rank
        return lo, lo + len(self)

    def __iter__(self):
        return iter(range(*self.bounds()))

    def __len__(self):
        return self.n // (self.world_size * self.bs) * self.bs

    def set_epoch(self, epoch: int):
        self.epoch = epoch


class GraphCollator:
    def __init__(self, graph: Graph, n_neighbors: int, n_layers: int, *, restarter: str = 'seq',
                 hist_len: Optional[int] = None, n_walks: Optional[int] = None,
                 walk_length: Optional[int] = None, alpha: float = 0.0):
        if n_layers < 1:
            raise ValueError('n_layers must be >= 1')
        if restarter not in ('seq', 'static'):
            raise NotImplementedError(f"restarter '{restarter}'")
        self.graph = graph
        self.n_nodes = graph.num_node
        self.n_neighbors, self.n_layers = n_neighbors, n_layers
        self.restarter, self.hist_len = restarter, hist_len
        self.n_walks, self.walk_length, self.alpha = n_walks, walk_length, alpha
        self._bitmap = None
        self._scratch = None

    # the pieces of the reference collator, each usable on its own -------------------------------
    def collate_memory_nodes(self, nids: torch.Tensor, ts64: torch.Tensor, ts_period: int = 0):
        """data_loader.py:105-131.  Device tensors in; returns (layers, (np involved ids, device involved ids),
        local_index).  layers[n_layers] holds the neighbors of the batch nodes, layers[l - 1] the neighbors of
        layers[l]'s neighbors queried at THEIR (float32) event times (the reference recurses with the float32 table it
        just produced, :124-131); every layer marks the involved-node bitmap."""
        dev = self.graph.device
        if self._bitmap is None or self._bitmap.device != dev:
            self._bitmap = torch.zeros(ops.bitmap_words(self.n_nodes), dtype=torch.int32, device=dev)
        k = self.n_neighbors
        layers = [None] * (self.n_layers + 1)
        layers[0] = (nids, None, None)
        q_n, q_t, period = nids, ts64, ts_period
        n_queries = 0
        for depth in range(self.n_layers, 0, -1):
            nn_, ne_, nt_, _ = ops.find_recent(self.graph.csr, q_n, q_t, k, ts_period=period, want_dirs=False,
                                               bitmap=self._bitmap)
            layers[depth] = (nn_, ne_, nt_)
            n_queries += q_n.numel()
            q_n, q_t, period = nn_.reshape(-1), nt_.reshape(-1).double(), 0
        cap = n_queries * (k + 1)
        involved = torch.empty(cap, dtype=torch.int64, device=dev)
        counts = torch.zeros(4, dtype=torch.int32, device=dev)
        local_index = torch.zeros(self.n_nodes, dtype=torch.int64, device=dev)
        ops.compact_involved(self._bitmap, self.n_nodes, involved, counts, local_index=local_index)
        u = int(counts[0])                                   # host sync: the drivers need the ids on the host
        involved = involved[:u]
        return layers, (involved.cpu().numpy(), involved), local_index

    def collate_restart_data(self, nids: torch.Tensor, ts64: torch.Tensor, ts_period: int = 0):
        """data_loader.py:95-168: de-duplicate positives (latest occurrence, float64 times), then the
        restarter's inputs at those times."""
        dev = self.graph.device
        n = nids.numel()
        if n > 2048 and (self._scratch is None or self._scratch.slot_ts.device != dev):
            self._scratch = ops.SelectScratch(self.n_nodes, dev)
        _, uniq, index, count = ops.select_latest(nids, ts64, self._scratch)
        p = int(count)
        uniq, index = uniq[:p], index[:p]
        period = ts_period or ts64.numel()
        sel_ts64 = ts64[index % period].contiguous()
        ts32 = sel_ts64.float()
        if self.restarter == 'seq':
            hn, he, ht, hd = ops.find_recent(self.graph.csr, uniq, sel_ts64, self.hist_len)
            anon = ops.anonymized_reindex(hn) if p else hn.clone()
            return SeqRestartData(index, uniq, ts32, hn, anon, he, ht, hd)
        _, _, prev_ts, _ = ops.find_recent(self.graph.csr, uniq, sel_ts64, 1, want_dirs=False)
        return StaticRestartData(index, uniq, ts32, prev_ts)

    def check_in_window(self, center_nodes: torch.Tensor, neighbors: torch.Tensor) -> torch.Tensor:
        """data_loader.py:61-67 given the target nodes' neighbor rows."""
        return ops.hit_window(center_nodes, neighbors)

    def __call__(self, batch: List[Tuple[int, int, int, float, int, int]]):
        src, dst, neg, ts, eids, labels = (np.array(x) for x in zip(*batch))
        dev = self.graph.device
        B = len(src)
        to_dev = lambda x, dt: torch.as_tensor(np.ascontiguousarray(x)).to(dt).to(dev)
        d_src, d_dst, d_neg = (to_dev(x, torch.int64) for x in (src, dst, neg))
        ts64 = to_dev(ts, torch.float64)
        batch_nids = torch.cat([d_src, d_dst, d_neg])
        layers, nodes, local_index = self.collate_memory_nodes(batch_nids, ts64, ts_period=B)
        restart_data = self.collate_restart_data(batch_nids[:2 * B], ts64, ts_period=B)
        # hit windows reuse the neighbor rows just computed: rows [0,B) = N(src), [B,2B) = N(dst),
        # [2B,3B) = N(neg), all at the events' times (data_loader.py:69-75)
        neigh = layers[self.n_layers][0]
        n_src, n_dst, n_neg = neigh[:B], neigh[B:2 * B], neigh[2 * B:]
        hit_data = HitData(self.check_in_window(d_src, n_dst), self.check_in_window(d_dst, n_src),
                           self.check_in_window(d_src, n_neg), self.check_in_window(d_neg, n_src))
        cg = ComputationGraph([layers, nodes], restart_data, hit_data, self.n_nodes, local_index=local_index)
        as_long = lambda x: torch.from_numpy(np.ascontiguousarray(x)).long()
        return (as_long(src), as_long(dst), as_long(neg), torch.from_numpy(np.ascontiguousarray(ts)).float(),
                as_long(eids), as_long(labels), cg)


class RandEdgeSampler:
    """Uniform negatives over the distinct sources / destinations (data_loader.py:283-313); the
    source draw is kept because it advances the RNG state the destination draw depends on."""

    def __init__(self, src_list: np.ndarray, dst_list: np.ndarray, seed: Optional[int] = None):
        self.seed = seed
        self.rng = np.random.RandomState(seed)
        self.src_list = np.unique(src_list)
        self.dst_list = np.unique(dst_list)

    def sample(self, size: int) -> Tuple[np.ndarray, np.ndarray]:
        si = self.rng.randint(0, len(self.src_list), size)
        di = self.rng.randint(0, len(self.dst_list), size)
        return self.src_list[si], self.dst_list[di]

    def reset_random_state(self):
        self.rng = np.random.RandomState(self.seed)

    def pre_sample_neg_dsts(self, n_total: int, bs: int = 200) -> np.ndarray:
        self.reset_random_state()
        chunks = [self.sample(min(bs, n_total - lo))[1] for lo in range(0, n_total, bs)]
        out = np.concatenate(chunks) if chunks else np.zeros(0, dtype=np.int64)
        assert len(out) == n_total
        return out


class InteractionData(Dataset):
    """A slice of the interaction stream (data_loader.py:214-280)."""

    def __init__(self, src, dst, ts, eids, labels, seed=0, eval=False, neg_dst=None):
        if len({len(x) for x in (src, dst, ts, eids, labels)}) != 1:
            raise AssertionError('columns of different lengths')
        self.src, self.dst, self.ts, self.eids, self.labels = src, dst, ts, eids, labels
        self.eval, self.seed = eval, seed
        self.neg_dst = None
        self.neg_dst_sampler = RandEdgeSampler(src, dst, seed)
        if eval:
            self.neg_dst = neg_dst if neg_dst is not None else \
                self.neg_dst_sampler.pre_sample_neg_dsts(len(ts), bs=200)

    def get_subset(self, start, end):
        cut = lambda x: x[start:end]
        return InteractionData(cut(self.src), cut(self.dst), cut(self.ts), cut(self.eids), cut(self.labels),
                               self.seed, self.eval, self.neg_dst)

    def get_neg_dst_item(self, i) -> int:
        if self.eval:
            return self.neg_dst[i]
        return self.neg_dst_sampler.sample(1)[1].item()

    def __getitem__(self, i):
        return (self.src[i], self.dst[i], self.get_neg_dst_item(i), self.ts[i], self.eids[i], self.labels[i])

    def __len__(self):
        return len(self.ts)

    def __repr__(self):
        n = len(set(self.src).union(self.dst))
        return f'Data(#edges={len(self)}, #nodes={n}, trange=({self.ts.min():.1f}, {self.ts.max():.1f}))'


def load_jodie_data(name: str, train_seed: int, *, root='.', data_seed=2020, val_p=0.7, test_p=0.85):
    """JODIE / TGN `ml_<name>.csv` (+ `.npy`, `_node.npy`) with the reference's chronological
    70/15/15 split and inductive-node masking (data_loader.py:316-404).  `random.sample` gets a sorted
    list (sets are rejected by Python >= 3.11, where the reference's own loader fails)."""
    import pandas as pd
    root = pathlib.Path(root)
    df = pd.read_csv(root / f'data/ml_{name}.csv')
    load = lambda p: np.load(p) if p.exists() else None
    efeats = load(root / f'data/ml_{name}.npy')
    nfeats = load(root / f'data/ml_{name}_node.npy')
    val_time, test_time = np.quantile(df.ts, [val_p, test_p])
    src, dst, eids, labels, ts = df.u.values, df.i.values, df.idx.values, df.label.values, df.ts.values
    full_data = InteractionData(src, dst, ts, eids, labels)
    random.seed(data_seed)
    nodes = set(src) | set(dst)
    late = set(src[ts > val_time]) | set(dst[ts > val_time])
    hidden = set(random.sample(sorted(late), int(0.1 * len(nodes))))
    visible_edge = ~np.isin(src, list(hidden)) & ~np.isin(dst, list(hidden))
    pick = lambda mask, **kw: InteractionData(*[x[mask] for x in (src, dst, ts, eids, labels)], **kw)
    train_data = pick((ts <= val_time) & visible_edge, seed=train_seed, eval=False)
    unseen = nodes - (set(train_data.src) | set(train_data.dst))
    new_edge = np.isin(src, list(unseen)) | np.isin(dst, list(unseen))
    val_mask, test_mask = (ts <= test_time) & (ts > val_time), ts > test_time
    return (nfeats, efeats, full_data, train_data, pick(val_mask, seed=0, eval=True),
            pick(test_mask, seed=2, eval=True), pick(val_mask & new_edge, seed=1, eval=True),
            pick(test_mask & new_edge, seed=3, eval=True))
