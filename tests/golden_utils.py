"""Helpers shared by the golden-fixture tests."""
import glob
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CASES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, '*.npz')))


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
