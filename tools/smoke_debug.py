"""Repeats the smoke comparison and reports which quantity deviates (debug aid)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from oracle import tiger_oracle as O
from www2023tiger_b200 import ops
from www2023tiger_b200.engine import TigerEngine
from www2023tiger_b200.init import perturb_biases, random_weights
from www2023tiger_b200.synthetic import NegativeSampler, StreamShape, make_stream
st = make_stream(StreamShape('smoke', 300, 40, 4000, 16, None, horizon=4000.), seed=0)
N, d, B, K = st.n_nodes, st.dim, 100, 10
neg = NegativeSampler(st.src, st.dst, seed=0).pre_sample_neg_dsts(st.n_events)
W = perturb_biases(random_weights(d, d, n_nodes=N, restarter='static', nonzero_static=True, seed=0))
dev = lambda x, dt: torch.as_tensor(x).to(dt).cuda().contiguous()
csr = ops.csr_build(dev(st.src, torch.int64), dev(st.dst, torch.int64), dev(st.ts, torch.float64), dev(st.eids, torch.int64), N)
eng = TigerEngine(W, csr, n_nodes=N, dim=d, efeats=dev(st.efeats, torch.float32), n_neighbors=K, n_head=2,
                  batch_size=B, restarter='static', lazy_restart=True)
graph = O.OracleGraph(st.src, st.dst, st.ts, st.eids, n_nodes=N)
refs = []
model = O.OracleTIGER(W, graph, N, d, st.efeats, None, n_neighbors=K, n_head=2, restarter='static')
uptodate = np.zeros(N, dtype=bool)
for ib in range(10, 16):
    s = slice(ib * B, (ib + 1) * B)
    b = O.collate(graph, st.src[s], st.dst[s], neg[s], st.ts[s], st.eids[s], K)
    rn = O.lazy_restart_nodes(b.involved, uptodate)
    model.restart(rn, np.full(len(rn), b.ts.min(), dtype=np.float32))
    ref = model.contrast_step(b)
    refs.append((s, ref['h_left_with_negs'].numpy().copy(), model.right_vals.numpy().copy(), model.left_vals.numpy().copy(),
                 model.msg_vals.numpy().copy()))
rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 12):
    eng.reset()
    line = []
    for (s, emb, rv, lv, mv) in refs:
        eng.set_batch(st.src[s], st.dst[s], neg[s], st.ts[s], st.eids[s])
        eng.step()
        e = (rel(eng.emb.cpu().numpy(), emb), rel(eng.right_vals.cpu().numpy(), rv), rel(eng.left_vals.cpu().numpy(), lv),
             rel(eng.msg_vals.cpu().numpy(), mv))
        line.append('/'.join(f'{x:.0e}' for x in e[:4]))
    bad = any(float(t) > 1e-5 for l in line for t in l.split('/'))
    print(('BAD ' if bad else 'ok  ') + ' | '.join(line), flush=True)
