"""Seeded synthetic JODIE-shaped interaction streams.

There is no network for the real JODIE csv files (reference loader:
tiger/data/data_loader.py:316-404), so benchmarks, parity tests and golden
fixtures all run on bipartite streams of the same *shape* as the datasets
BASELINE.json names: users ``1..n_u``, items ``n_u+1..n_u+n_i``, id 0 reserved for
padding, Zipf-distributed endpoints, integer-second float64 timestamps (so ties
occur), edge ids ``1..E`` and an edge-feature table whose row 0 is the padding row.
"""
from dataclasses import dataclass
from typing import Optional

import numpy as np


@dataclass
class StreamShape:
    name: str
    n_users: int
    n_items: int
    n_events: int
    efeat_dim: int          # 0 => no edge-feature table (efeats=None)
    dim: Optional[int]      # --dim flag (only used when a table is absent)
    horizon: float = 2.678e6
    restarter: str = 'seq'
    msg_src: str = 'left'
    upd_src: str = 'right'


# The five BASELINE.json configs.
SHAPES = {
    'wikipedia': StreamShape('wikipedia', 8227, 1000, 157474, 172, None, restarter='seq'),
    'reddit': StreamShape('reddit', 10000, 984, 672447, 172, None, restarter='static'),
    'mooc': StreamShape('mooc', 7047, 97, 411749, 4, 100, restarter='seq',
                        msg_src='right', upd_src='right'),
    'lastfm': StreamShape('lastfm', 980, 1000, 1293103, 0, 100, horizon=1.4e8, restarter='seq'),
    'scaled': StreamShape('scaled', 900000, 100000, 50_000_000, 172, None, restarter='seq'),
}


@dataclass
class Stream:
    shape: StreamShape
    src: np.ndarray      # int64 [E]
    dst: np.ndarray      # int64 [E]
    ts: np.ndarray       # float64 [E], non-decreasing
    eids: np.ndarray     # int64 [E], 1..E
    labels: np.ndarray   # int64 [E], zeros
    efeats: Optional[np.ndarray]  # float32 [E+1, de] (row 0 zeros) or None
    nfeats: Optional[np.ndarray]  # float32 [N, d] or None

    @property
    def n_nodes(self) -> int:
        """Number of rows of every per-node table (= max id + 1, id 0 is padding)."""
        return self.shape.n_users + self.shape.n_items + 1

    @property
    def n_events(self) -> int:
        return len(self.src)

    @property
    def dim(self) -> int:
        """Memory / node-feature width, as init_utils.py:131-136 derives it."""
        if self.nfeats is not None:
            return self.nfeats.shape[1]
        if self.shape.dim is not None:
            return self.shape.dim
        if self.efeats is not None:
            return self.efeats.shape[1]
        if self.shape.efeat_dim > 0:      # edge table generated on the device (with_efeats=False)
            return self.shape.efeat_dim
        raise ValueError('dim undefined')


def _zipf_ids(rng: np.random.RandomState, n: int, size: int, s: float, first_id: int) -> np.ndarray:
    p = 1.0 / np.arange(1, n + 1, dtype=np.float64) ** s
    cdf = np.cumsum(p)
    cdf /= cdf[-1]
    ranks = np.searchsorted(cdf, rng.random_sample(size), side='right')
    ranks = np.minimum(ranks, n - 1)
    perm = rng.permutation(n)  # popularity rank -> id offset (no id/popularity locality)
    return (perm[ranks] + first_id).astype(np.int64)


def make_stream(shape, seed: int = 0, n_events: Optional[int] = None,
                with_efeats: bool = True, nfeat_dim: int = 0) -> Stream:
    """Generate the stream for a named shape (or a StreamShape).

    ``with_efeats=False`` skips materialising the N(0,1) edge table on the host
    (callers that generate it on the device pass this for the large shapes).
    ``nfeat_dim>0`` adds an N(0,1) node-feature table (the synthetic BASELINE shapes
    use nfeats=None; non-zero tables are exercised by the parity tests only).
    """
    if isinstance(shape, str):
        shape = SHAPES[shape]
    rng = np.random.RandomState(seed)
    E = shape.n_events if n_events is None else n_events
    src = _zipf_ids(rng, shape.n_users, E, 0.8, 1)
    dst = _zipf_ids(rng, shape.n_items, E, 1.0, shape.n_users + 1)
    ts = np.floor(np.sort(rng.uniform(0.0, shape.horizon, E))).astype(np.float64)
    eids = np.arange(1, E + 1, dtype=np.int64)
    labels = np.zeros(E, dtype=np.int64)
    efeats = None
    if shape.efeat_dim > 0 and with_efeats:
        efeats = rng.standard_normal((E + 1, shape.efeat_dim)).astype(np.float32)
        efeats[0] = 0.0
    nfeats = None
    if nfeat_dim > 0:
        nfeats = rng.standard_normal((shape.n_users + shape.n_items + 1, nfeat_dim)).astype(np.float32)
        nfeats[0] = 0.0
    return Stream(shape, src, dst, ts, eids, labels, efeats, nfeats)


class NegativeSampler:
    """Uniform negative destinations, drawn exactly like the reference's
    RandEdgeSampler (tiger/data/data_loader.py:283-313): two ``randint`` draws per
    call (the source draw is discarded but consumes RNG state), pre-sampled in
    chunks of 200 for evaluation streams."""

    def __init__(self, src: np.ndarray, dst: np.ndarray, seed: Optional[int] = None):
        self.seed = seed
        self.rng = np.random.RandomState(seed)
        self.src_list = np.unique(src)
        self.dst_list = np.unique(dst)

    def sample(self, size: int):
        src_index = self.rng.randint(0, len(self.src_list), size)
        dst_index = self.rng.randint(0, len(self.dst_list), size)
        return self.src_list[src_index], self.dst_list[dst_index]

    def reset_random_state(self):
        self.rng = np.random.RandomState(self.seed)

    def pre_sample_neg_dsts(self, n_total: int, bs: int = 200) -> np.ndarray:
        self.reset_random_state()
        out = []
        residual = n_total
        while residual > 0:
            take = min(bs, residual)
            out.append(self.sample(take)[1])
            residual -= take
        return np.concatenate(out) if out else np.zeros(0, dtype=np.int64)
