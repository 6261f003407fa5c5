"""Pins the CPU oracle (oracle/tiger_oracle.py) against fixtures produced by the
unmodified reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import tiger_oracle as O
from golden_utils import CASES, Golden, assert_close

FP_TOL = 2e-6   # oracle vs reference: same torch-CPU kernels, only op grouping differs


@pytest.fixture(scope='module', params=CASES)
def golden(request):
    return Golden(request.param)


def make_oracle(g: Golden):
    graph = O.OracleGraph(g.src, g.dst, g.ts, g.eids, n_nodes=g.N)
    model = O.OracleTIGER(g.W, graph, g.N, g.dim, g.efeats, g.nfeats, n_neighbors=g.K, n_head=g.n_heads,
                          msg_src=g.msg_src, upd_src=g.upd_src, restarter=g.restarter, hist_len=g.hist_len)
    return graph, model


def test_adjacency_matches_reference(golden):
    graph, _ = make_oracle(golden)
    assert np.array_equal(np.diff(graph.indptr), golden.z['adj_lens'])
    assert np.array_equal(graph.nbr, golden.z['adj_nbr'])
    assert np.array_equal(graph.eid, golden.z['adj_eid'])
    assert np.array_equal(graph.ts, golden.z['adj_ts'])
    assert np.array_equal(graph.flag, golden.z['adj_flag'])


def test_known_answer_index_functions(golden):
    z = golden.z
    for fn in (O.select_latest, O.select_latest_scan):
        u, ix = fn(z['kat_sl_ids'], z['kat_sl_ts'])
        assert np.array_equal(u, z['kat_sl_unique']) and np.array_equal(ix, z['kat_sl_index'])
    assert np.array_equal(O.anonymized_reindex(z['kat_anon_in']), z['kat_anon_out'])
    graph, _ = make_oracle(golden)
    hn, he, ht, hd = graph.get_history(z['kat_hist_q_nids'], z['kat_hist_q_ts'], 7)
    assert np.array_equal(hn, z['kat_hist_nids']) and np.array_equal(he, z['kat_hist_eids'])
    assert np.array_equal(ht, z['kat_hist_ts']) and ht.dtype == np.float32
    assert np.array_equal(hd, z['kat_hist_dirs'])


def test_survey_q14_examples():
    got = O.anonymized_reindex(np.array([[0, 0, 5, 7, 5], [3, 4, 3, 3, 9]]))
    assert got.tolist() == [[0, 0, 1, 2, 1], [2, 3, 2, 2, 1]]


def test_stream_replay_matches_reference(golden):
    g = golden
    graph, model = make_oracle(g)
    uptodate = np.zeros(g.N, dtype=bool)
    for ib in range(g.n_batches):
        src, dst, neg, ts, eids = g.batch(ib)
        b = O.collate(graph, src, dst, neg, ts, eids, g.K, restarter=g.restarter, hist_len=g.hist_len)
        # --- integer / index parity: bit exact ---
        assert np.array_equal(b.neigh_nids, g.b(ib, 'neigh_nids'))
        assert np.array_equal(b.neigh_eids, g.b(ib, 'neigh_eids'))
        assert np.array_equal(b.neigh_ts, g.b(ib, 'neigh_ts'))
        assert np.array_equal(b.involved, g.b(ib, 'involved'))
        assert np.array_equal(b.local_index, g.b(ib, 'local_index'))
        for hn in ('src_hits', 'dst_hits', 'neg_src_hits', 'neg_dst_hits'):
            assert np.array_equal(getattr(b, hn), g.b(ib, hn))
        r = b.restart
        assert np.array_equal(r.index, g.b(ib, 'r_index')) and np.array_equal(r.nids, g.b(ib, 'r_nids'))
        assert np.array_equal(r.ts, g.b(ib, 'r_ts'))
        if g.restarter == 'seq':
            assert np.array_equal(r.hist_nids, g.b(ib, 'r_hist_nids'))
            assert np.array_equal(r.anonymized_ids, g.b(ib, 'r_anon'))
            assert np.array_equal(r.hist_eids, g.b(ib, 'r_hist_eids'))
            assert np.array_equal(r.hist_ts, g.b(ib, 'r_hist_ts'))
            assert np.array_equal(r.hist_dirs, g.b(ib, 'r_hist_dirs'))
        else:
            assert np.array_equal(r.prev_ts, g.b(ib, 'r_prev_ts'))
        # --- lazy restart ---
        if g.lazy_restart:
            rn = O.lazy_restart_nodes(b.involved, uptodate)
            assert np.array_equal(rn, g.b(ib, 'restart_nids'))
            if len(rn):
                t0 = np.full(len(rn), b.ts.min(), dtype=np.float32)
                hl, hr, pt = model.restarter_forward(rn, t0)
                assert_close(hl.numpy(), g.b(ib, 'restart_hl'), FP_TOL, 'restart_hl')
                assert_close(hr.numpy(), g.b(ib, 'restart_hr'), FP_TOL, 'restart_hr')
                assert np.array_equal(pt.numpy(), g.b(ib, 'restart_pt'))
                model.restart(rn, t0)
        assert np.array_equal(np.nonzero(model.has_msg)[0], g.b(ib, 'pending_before'))
        out = model.contrast_step(b)
        what = f'{g.name} batch {ib} '
        assert_close(out['h_left'].numpy(), g.b(ib, 'h_left'), FP_TOL, what + 'h_left')
        assert_close(out['pos_scores'].numpy(), g.b(ib, 'pos_scores'), FP_TOL, what + 'pos')
        assert_close(out['neg_scores'].numpy(), g.b(ib, 'neg_scores'), FP_TOL, what + 'neg')
        assert_close(out['loss'].numpy(), g.b(ib, 'loss'), FP_TOL, what + 'loss')
        assert_close(out['h_prev_left'].numpy(), g.b(ib, 'h_prev_left'), FP_TOL, what + 'hpl')
        assert_close(out['h_prev_right'].numpy(), g.b(ib, 'h_prev_right'), FP_TOL, what + 'hpr')
        sl, sr, _ = model.restarter_on_batch(b)
        assert_close(sl.numpy(), g.b(ib, 'surrogate_left'), FP_TOL, what + 'surrogate_left')
        assert_close(sr.numpy(), g.b(ib, 'surrogate_right'), FP_TOL, what + 'surrogate_right')
        assert_close(model.mutual_loss(b, out).numpy(), g.b(ib, 'mutual_loss'), FP_TOL, what + 'mutual')
        assert_close(model.left_vals.numpy(), g.b(ib, 'left_vals'), FP_TOL, what + 'left_vals')
        assert_close(model.right_vals.numpy(), g.b(ib, 'right_vals'), FP_TOL, what + 'right_vals')
        assert np.array_equal(model.left_ts.numpy(), g.b(ib, 'left_ts'))
        assert np.array_equal(model.right_ts.numpy(), g.b(ib, 'right_ts'))
        assert np.array_equal(model.msg_ts.numpy(), g.b(ib, 'msg_ts'))
        assert_close(model.msg_vals.numpy(), g.b(ib, 'msg_vals'), FP_TOL, what + 'msg_vals')
        assert np.array_equal(np.nonzero(model.has_msg)[0], g.b(ib, 'pending_after'))
    model.flush_msg()
    assert_close(model.right_vals.numpy(), g.z['flush_right_vals'], FP_TOL, 'flush vals')
    assert np.array_equal(model.right_ts.numpy(), g.z['flush_right_ts'])


def test_select_latest_vectorised_equals_scan():
    rng = np.random.RandomState(3)
    for n in (0, 1, 5, 400, 1500):
        ids = rng.randint(0, max(2, n // 3), n)
        ts = np.floor(rng.uniform(0, 30, n)).astype(np.float32)
        a, b = O.select_latest(ids, ts), O.select_latest_scan(ids, ts)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_chunk_range_partitions_like_chunk_sampler():
    n, bs = 157474, 200
    for W in (1, 2, 4, 8):
        ranges = [O.chunk_range(n, r, W, bs, seed=0) for r in range(W)]
        L = n // (W * bs) * bs
        for r, (lo, hi) in enumerate(ranges):
            assert hi - lo == L and lo == ranges[0][0] + r * L
        assert 0 <= ranges[0][0] <= n % (W * bs) and ranges[-1][1] <= n


# ------------------------------------------------------------------------------------------
# BASELINE dimensions: d / d_e / K / hist_len / batch of the BASELINE.json configs, on the BASELINE-shaped
# streams (tests/golden/make_golden_full.py)
# ------------------------------------------------------------------------------------------
from golden_utils import FULL_CASES, FullGolden, check_full_batch   # noqa: E402


def replay_full_on_oracle(g: FullGolden):
    """Yields (recorded-batch index, dict of results in the fixture's names) for the oracle port."""
    src, dst, ts, eids = g.stream_prefix()
    graph = O.OracleGraph(src, dst, ts, eids, n_nodes=g.N)
    model = O.OracleTIGER(g.W, graph, g.N, g.dim, g.efeats, None, n_neighbors=g.K, n_head=g.n_heads,
                          msg_src=g.msg_src, upd_src=g.upd_src, restarter=g.restarter, hist_len=g.hist_len)
    uptodate = np.zeros(g.N, dtype=bool)
    with torch.no_grad():
        for ib in range(g.warm + g.rec):
            b = O.collate(graph, *g.batch(ib), g.K, restarter=g.restarter, hist_len=g.hist_len)
            rn = O.lazy_restart_nodes(b.involved, uptodate)
            r_ts = np.full(len(rn), b.ts.min(), dtype=np.float32)
            got = {}
            if ib >= g.warm and len(rn):
                hl, hr, pt = model.restarter_forward(rn, r_ts)
                got.update(restart_hl=hl.numpy(), restart_hr=hr.numpy(), restart_pt=pt.numpy())
            model.restart(rn, r_ts)
            out = model.contrast_step(b)
            if ib < g.warm:
                continue
            sl, sr, _ = model.restarter_on_batch(b)
            w = np.zeros(2 * g.bs, dtype=np.uint8)
            w[b.restart.index] = 1
            got.update(neigh_nids=b.neigh_nids, neigh_eids=b.neigh_eids, neigh_ts=b.neigh_ts, involved=b.involved,
                       restart_nids=rn, outdated=out['outdated'], winner=w, h_left=out['h_left'].numpy(),
                       pos_scores=out['pos_scores'].numpy(), neg_scores=out['neg_scores'].numpy(),
                       loss=out['loss'].numpy(), mutual_loss=model.mutual_loss(b, out).numpy(),
                       h_prev_left=out['h_prev_left'].numpy(), h_prev_right=out['h_prev_right'].numpy(),
                       surrogate_left=sl.numpy(), surrogate_right=sr.numpy(),
                       left_vals=model.left_vals.numpy(), right_vals=model.right_vals.numpy(),
                       msg_vals=model.msg_vals.numpy(), left_ts=model.left_ts.numpy(),
                       right_ts=model.right_ts.numpy(), msg_ts=model.msg_ts.numpy(),
                       pending_after=np.nonzero(model.has_msg)[0])
            yield ib - g.warm, got


@pytest.mark.parametrize('name', FULL_CASES)
def test_oracle_matches_reference_at_baseline_dimensions(name):
    g = FullGolden(name)
    n = 0
    for ir, got in replay_full_on_oracle(g):
        check_full_batch(g, ir, got, tol=5e-6, what='oracle ')
        n += 1
    assert n == g.rec
