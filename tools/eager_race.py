"""Debug aid: does the first batch after a reset reproduce bit for bit with eager launches when the side-stream
branches (TIGER_EAGER_BRANCHES=1) and / or programmatic launches (TIGER_EAGER_PDL=1) are enabled in eager mode?"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch, time
import gpu_utils as gu
from www2023tiger_b200.init import perturb_biases, random_weights
from www2023tiger_b200.synthetic import NegativeSampler, StreamShape, make_stream
st = make_stream(StreamShape('r', 300, 40, 4000, 16, None, horizon=4000.), seed=0)
B, K = 100, 10
neg = NegativeSampler(st.src, st.dst, seed=0).pre_sample_neg_dsts(st.n_events)
N, d = st.n_nodes, st.dim
W = perturb_biases(random_weights(d, d, n_nodes=N, restarter='static', nonzero_static=True, seed=0))
csr = gu.device_csr(st.src, st.dst, st.ts, st.eids, N)
e = gu.engine_from(W, csr, N=N, dim=d, efeats=st.efeats, nfeats=None, K=K, H=2, B=B, msg_src='left',
                   upd_src='right', restarter='static', lazy_restart=True, want_targets=False)
cols = tuple(a[10 * B:11 * B] for a in (st.src, st.dst, neg, st.ts, st.eids))
ref = None
bad = 0
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 400):
    e.reset()
    e.set_batch(*cols)
    torch.cuda.synchronize()
    if it % 3 == 0:
        time.sleep(0.002)        # idle GPU, as between the oracle-interleaved steps of smoke()
    e.step()
    torch.cuda.synchronize()
    if ref is None:
        ref = e.emb.clone()
    elif not torch.equal(e.emb, ref):
        bad += 1
        print('iteration', it, 'max abs diff', float((e.emb - ref).abs().max()), flush=True)
print('mismatches:', bad)
