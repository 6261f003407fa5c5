"""Helpers shared by the golden-fixture tests."""
import glob
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
_ALL = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, '*.npz')))
CASES = [c for c in _ALL if not c.startswith(('full_', 'train_', 'var_'))]   # toy-size fixtures (make_golden.py)
VAR_CASES = [c for c in _ALL if c.startswith('var_')]           # non-default operator variants (make_golden.py)
TRAIN_CASES = [c for c in _ALL if c.startswith('train_')]       # training-step fixtures (make_golden_train.py)
FULL_CASES = [c for c in _ALL if c.startswith('full_')]         # BASELINE-dimension fixtures (make_golden_full.py)


class Golden:
    def __init__(self, name):
        self.name = name
        self.z = dict(np.load(os.path.join(GOLDEN_DIR, name + '.npz')))
        m = lambda k: self.z['meta_' + k]
        self.bs = int(m('bs'))
        self.n_batches = int(m('n_batches'))
        self.lazy_restart = bool(int(m('lazy_restart')))
        self.K = int(m('n_neighbors'))
        self.n_heads = int(m('n_heads'))
        self.hist_len = int(m('hist_len'))
        self.N = int(m('n_nodes'))
        self.dim = int(m('dim'))
        self.restarter = str(m('restarter'))
        self.msg_src = str(m('msg_src'))
        self.upd_src = str(m('upd_src'))
        opt = lambda k, default: type(default)(self.z['meta_' + k]) if 'meta_' + k in self.z else default
        self.n_layers, self.hit_type = opt('n_layers', 1), opt('hit_type', 'bin')
        self.tsfm, self.upd = opt('tsfm', 'id'), opt('upd', 'gru')
        self.src, self.dst, self.ts = self.z['stream_src'], self.z['stream_dst'], self.z['stream_ts']
        self.eids, self.neg = self.z['stream_eids'], self.z['stream_neg']
        self.efeats = self.z.get('stream_efeats')
        self.nfeats = self.z.get('stream_nfeats')
        self.W = {k[2:]: torch.from_numpy(v) for k, v in self.z.items() if k.startswith('w_')}

    def batch(self, ib):
        lo, hi = ib * self.bs, (ib + 1) * self.bs
        return self.src[lo:hi], self.dst[lo:hi], self.neg[lo:hi], self.ts[lo:hi], self.eids[lo:hi]

    def b(self, ib, key):
        return self.z[f'b{ib}_{key}']

    def has(self, ib, key):
        return f'b{ib}_{key}' in self.z


def assert_close(a, b, tol=1e-5, what=''):
    """max|a-b| / max(|b|, tiny) <= tol  (the parity norm of SURVEY.md H2)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f'{what}: shape {a.shape} vs {b.shape}'
    if a.size == 0:
        return
    scale = max(np.abs(b).max(), 1e-30)
    err = np.abs(a - b).max() / scale
    assert err <= tol, f'{what}: rel err {err:.3e} > {tol}'


class FullGolden:
    """A BASELINE-dimension fixture of tests/golden/make_golden_full.py.  The stream and the parameters are
    regenerated from their seeds (they are deterministic functions of them), only the reference's outputs are read
    from the file."""

    def __init__(self, name):
        from www2023tiger_b200.init import perturb_biases, random_weights
        from www2023tiger_b200.synthetic import SHAPES, NegativeSampler, make_stream
        self.name = name
        self.z = dict(np.load(os.path.join(GOLDEN_DIR, name + '.npz')))
        m = lambda k: self.z['meta_' + k]
        self.shape = SHAPES[str(m('shape'))]
        self.restarter = str(m('restarter'))
        self.start, self.warm, self.rec = int(m('start')), int(m('warm')), int(m('rec'))
        self.bs, self.K, self.n_heads, self.hist_len = int(m('bs')), int(m('n_neighbors')), int(m('n_heads')), int(m('hist_len'))
        self.n_rows = int(m('n_rows'))
        st = make_stream(self.shape, seed=0)
        self.st = st
        self.neg = NegativeSampler(st.src, st.dst, seed=0).pre_sample_neg_dsts(st.n_events, self.bs)
        self.N, self.dim = st.n_nodes, st.dim
        assert self.dim == int(m('dim'))
        self.E = self.start + (self.warm + self.rec) * self.bs       # events the reference graph was built from
        self.efeats = st.efeats[:self.E + 1] if st.efeats is not None else None
        de = st.efeats.shape[1] if st.efeats is not None else self.dim
        self.W = perturb_biases(random_weights(self.dim, de, n_nodes=self.N, restarter=self.restarter,
                                               hist_len=self.hist_len, seed=int(m('weight_seed')), nonzero_static=True))
        self.msg_src, self.upd_src = self.shape.msg_src, self.shape.upd_src
        self.rows_kept = np.concatenate([np.arange(self.n_rows // 2), self.bs + np.arange(self.n_rows // 2)])

    def stream_prefix(self):
        E, st = self.E, self.st
        return st.src[:E], st.dst[:E], st.ts[:E], st.eids[:E]

    def batch(self, ib):
        """ib counts from the first warm-up batch."""
        lo = self.start + ib * self.bs
        s = slice(lo, lo + self.bs)
        st = self.st
        return st.src[s], st.dst[s], self.neg[s], st.ts[s], st.eids[s]

    def b(self, ir, key):
        """ir counts recorded batches (after the warm-up)."""
        return self.z[f'b{ir}_{key}']

    def has(self, ir, key):
        return f'b{ir}_{key}' in self.z


def check_full_batch(g: 'FullGolden', ir: int, got: dict, tol=1e-5, what=''):
    """Compares one recorded batch.  `got` holds numpy arrays under the fixture's names; wide float tensors are
    given in full (`h_left` [2B,d], tables [N,width]) and reduced here the way the generator reduced them."""
    what = f'{g.name} rec {ir} {what}'
    for k in ('neigh_nids', 'neigh_eids', 'involved', 'restart_nids', 'outdated', 'pending_after'):
        if k in got:
            assert np.array_equal(np.asarray(got[k]).astype(np.int64), g.b(ir, k).astype(np.int64)), what + k
    if 'neigh_ts' in got:
        assert np.array_equal(got['neigh_ts'], g.b(ir, 'neigh_ts')), what + 'neigh_ts'
    if 'winner' in got:
        w = np.zeros(2 * g.bs, dtype=np.uint8)
        w[g.b(ir, 'r_index')] = 1
        assert np.array_equal(np.asarray(got['winner']).astype(np.uint8), w), what + 'winners'
    rowsum = lambda a: np.asarray(a, dtype=np.float64).sum(1)
    if 'h_left' in got:
        assert_close(got['h_left'][g.rows_kept], g.b(ir, 'h_left'), tol, what + 'h_left rows')
        # a row sum cancels: its error is measured against the rows' magnitude
        scale = float(np.abs(g.b(ir, 'h_left')).max()) * np.sqrt(got['h_left'].shape[1])
        assert np.abs(rowsum(got['h_left']) - g.b(ir, 'h_left_rowsum')).max() <= tol * scale, what + 'h_left sums'
    for k in ('pos_scores', 'neg_scores'):
        if k in got:
            assert_close(got[k], g.b(ir, k), tol, what + k)
    if 'loss' in got:
        assert_close(np.reshape(got['loss'], (1,)), g.b(ir, 'loss').reshape(1), tol, what + 'loss')
    if 'mutual_loss' in got:
        assert_close(np.reshape(got['mutual_loss'], (1,)), g.b(ir, 'mutual_loss').reshape(1), 2 * tol, what + 'mutual')
    for k in ('h_prev_left', 'h_prev_right', 'surrogate_left', 'surrogate_right', 'restart_hl', 'restart_hr'):
        if k in got and g.has(ir, k + '_rowsum'):
            a = np.asarray(got[k])
            ref = g.b(ir, k + '_rowsum')
            scale = max(float(np.abs(a).max()), 1e-30) * np.sqrt(a.shape[1])
            assert np.abs(rowsum(a) - ref).max() <= tol * scale, what + k + ' sums'
            if g.has(ir, k):
                full = g.b(ir, k)
                assert_close(a[:len(full)], full, tol, what + k + ' rows')
    if 'restart_pt' in got and g.has(ir, 'restart_pt'):
        assert np.array_equal(got['restart_pt'], g.b(ir, 'restart_pt')), what + 'restart_pt'
    kept, pos = g.b(ir, 'kept_nodes').astype(np.int64), g.b(ir, 'pos_nodes').astype(np.int64)
    for t in ('left_vals', 'right_vals', 'msg_vals'):
        if t in got:
            tab = np.asarray(got[t])
            assert_close(tab[kept], g.b(ir, t + '_rows'), tol, what + t + ' rows')
            s = np.array([tab.astype(np.float64).sum(), np.abs(tab.astype(np.float64)).sum()])
            ref = g.b(ir, t + '_sums')
            assert abs(s[1] - ref[1]) <= tol * ref[1] and abs(s[0] - ref[0]) <= tol * ref[1], what + t + ' sums'
    for t in ('left_ts', 'right_ts', 'msg_ts'):
        if t in got:
            assert np.array_equal(np.asarray(got[t])[pos], g.b(ir, t + '_pos')), what + t
