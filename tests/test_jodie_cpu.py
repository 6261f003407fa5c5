"""f4: the on-disk format.  `load_jodie_data` of the drop-in package (rewritten for Python >= 3.11, where the
reference's own `random.sample(set, k)` raises) against the reference's loader run in a subprocess with exactly that
one call made 3.11-safe (random.sample over the SORTED set, which is what the drop-in does)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from jodie_utils import write_toy_dataset
from www2023tiger_b200.tiger.data.data_loader import load_jodie_data

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference' if os.path.isdir('/root/reference/tiger') else os.path.join(ROOT, 'oracle', '_ref')

REF_SCRIPT = r'''
import random, sys, json
import numpy as np
sys.path.insert(0, sys.argv[1])
sys.path.insert(0, sys.argv[2])           # torch_scatter shim
_orig = random.sample
random.sample = lambda pop, k: _orig(sorted(pop), k)      # the only change: sets are rejected by Python >= 3.11
from tiger.data.data_loader import load_jodie_data
out = load_jodie_data('toy', train_seed=5, root=sys.argv[3])
names = ['full', 'train', 'val', 'test', 'ind_val', 'ind_test']
res = {'nfeats_shape': list(out[0].shape) if out[0] is not None else None, 'efeats_sum': float(out[1].sum())}
for n, d in zip(names, out[2:]):
    res[n] = {'eids': d.eids.tolist(), 'neg': (d.neg_dst.tolist() if d.neg_dst is not None else None), 'eval': bool(d.eval),
              'seed': d.seed}
print(json.dumps(res))
'''


def test_load_jodie_data_properties(tmp_path):
    st = write_toy_dataset(str(tmp_path))
    nfeats, efeats, full, train, val, test, ind_val, ind_test = load_jodie_data('toy', train_seed=5, root=str(tmp_path))
    assert nfeats.shape == (st.n_nodes, 6) and np.array_equal(efeats, st.efeats)
    assert len(full) == st.n_events and np.array_equal(full.src, st.src) and np.array_equal(full.ts, st.ts)
    val_time, test_time = np.quantile(st.ts, [0.7, 0.85])
    assert train.ts.max() <= val_time and val.ts.min() > val_time and val.ts.max() <= test_time and test.ts.min() > test_time
    # chronological 70 / 15 / 15 split; the training set additionally loses the edges of the hidden (inductive) nodes
    assert len(val) + len(test) + int((st.ts <= val_time).sum()) == st.n_events and 0 < len(train) < (st.ts <= val_time).sum()
    train_nodes = set(train.src) | set(train.dst)
    for ind in (ind_val, ind_test):
        assert len(ind) > 0
        assert all((s not in train_nodes) or (d not in train_nodes) for s, d in zip(ind.src, ind.dst))
    # evaluation sets carry pre-sampled negatives (seeds 0, 2, 1, 3), the training set samples on the fly
    assert not train.eval and train.neg_dst is None and train.seed == 5
    for d, seed in ((val, 0), (test, 2), (ind_val, 1), (ind_test, 3)):
        assert d.eval and d.seed == seed and len(d.neg_dst) == len(d)
    src, dst, neg, ts, eid, label = train[0]
    assert neg in set(train.dst)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'tiger')), reason='no reference sources (vendor with tools/vendor_ref.py)')
def test_load_jodie_data_equals_reference_loader(tmp_path):
    write_toy_dataset(str(tmp_path))
    shim = os.path.join(ROOT, 'tests', 'golden', '_shim')
    res = subprocess.run([sys.executable, '-c', REF_SCRIPT, REF, shim, str(tmp_path)], capture_output=True, text=True,
                         timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    ref = json.loads(res.stdout.strip().splitlines()[-1])
    out = load_jodie_data('toy', train_seed=5, root=str(tmp_path))
    assert list(out[0].shape) == ref['nfeats_shape'] and abs(float(out[1].sum()) - ref['efeats_sum']) < 1e-3
    for n, d in zip(['full', 'train', 'val', 'test', 'ind_val', 'ind_test'], out[2:]):
        r = ref[n]
        assert d.eids.tolist() == r['eids'], n                       # the same events in every split
        assert bool(d.eval) == r['eval'] and d.seed == r['seed'], n
        assert (d.neg_dst.tolist() if d.neg_dst is not None else None) == r['neg'], n   # the same pre-sampled negatives
