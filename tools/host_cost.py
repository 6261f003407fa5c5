"""CPU cost of one StreamRunner.submit_host call (the host must stay ahead of a ~75 us GPU step): 48 submissions
into 64 staging slots, timed before anything is waited for."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import gpu_utils as gu
from www2023tiger_b200.engine import StreamRunner
from www2023tiger_b200.init import random_weights
from www2023tiger_b200.synthetic import NegativeSampler, StreamShape, make_stream
st = make_stream(StreamShape('h', 700, 90, 30000, 16, None), seed=5)
B, K = 200, 10
neg = NegativeSampler(st.src, st.dst, seed=0).pre_sample_neg_dsts(st.n_events)
N, d = st.n_nodes, st.dim
W = random_weights(d, st.efeats.shape[1], n_nodes=N, restarter='static', seed=3)
csr = gu.device_csr(st.src, st.dst, st.ts, st.eids, N)
e = gu.engine_from(W, csr, N=N, dim=d, efeats=st.efeats, nfeats=None, K=K, H=2, B=B, msg_src='left', upd_src='right',
                   restarter='static', lazy_restart=True, want_targets=False)
cols = lambda ib: tuple(a[3000 + ib * B:3000 + (ib + 1) * B] for a in (st.src, st.dst, neg, st.ts, st.eids))
runner = StreamRunner(e, n_slots=64)
e.set_batch(*cols(0))
runner.capture(warmup=1)
e.reset()
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    slots = [runner.submit_host(*cols(ib)) for ib in range(48)]
    t1 = time.perf_counter()
    for s in slots:
        runner.wait(s)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f'submit_host: {(t1 - t0) / 48 * 1e6:.1f} us CPU per step; drained after {(t2 - t0) / 48 * 1e6:.1f} us per step')
