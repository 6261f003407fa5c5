"""CPU oracle for TIGER's per-batch temporal memory path.

TEST INFRASTRUCTURE ONLY.  This module is a from-scratch CPU restatement (numpy for
the integer/index work, torch-CPU fp32 for the floating-point work) of the algorithm in
the reference repository yzhang1918/www2023tiger.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and only as the checker / the reported CPU baseline.
Nothing under ``www2023tiger_b200/`` imports it; the product path has no CPU fallback.

Parity status: PINNED.  ``tests/golden/*.npz`` hold inputs, weights and per-batch
outputs produced by the *unmodified* reference classes (``tests/golden/make_golden.py``,
run in the build container with /root/reference on sys.path and a 12-line
``torch_scatter.scatter_max`` shim that implements torch_scatter's CPU tie rule);
``tests/test_oracle_golden.py`` checks every function here against them.

The one third-party arithmetic dependency of the reference, ``torch_scatter.scatter_max``
(un-vendored, unpinned, used at tiger/model/utils.py:15), is restated from its published
CPU semantics: sequential scan, strict ``>`` update, so the lowest position wins ties.

All ``file:line`` citations are relative to the reference repository root.
"""
import math
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

F = torch.nn.functional


# --------------------------------------------------------------------------------------
# a1  temporal adjacency (tiger/data/graph.py:11-36, 226-241)
# --------------------------------------------------------------------------------------
class OracleGraph:
    """Per-node, time-sorted adjacency.

    data2adjlist (graph.py:226-241) appends ``(dst, eid, t, 0)`` to the list of ``src``
    and ``(src, eid, t, 1)`` to the list of ``dst`` for every event in stream order;
    Graph.__init__ (graph.py:30-36) then sorts each list by time with Python's stable
    sort.  Equivalent: a stable sort of the 2E interleaved entries by (owner, time).
    """

    def __init__(self, src, dst, ts, eids, n_nodes: Optional[int] = None):
        src = np.asarray(src, dtype=np.int64)
        dst = np.asarray(dst, dtype=np.int64)
        ts = np.asarray(ts, dtype=np.float64)
        eids = np.asarray(eids, dtype=np.int64)
        E = len(src)
        if n_nodes is None:
            n_nodes = int(max(src.max(), dst.max())) + 1 if E else 1
        self.num_node = n_nodes
        owner = np.empty(2 * E, dtype=np.int64)
        other = np.empty(2 * E, dtype=np.int64)
        owner[0::2], owner[1::2] = src, dst
        other[0::2], other[1::2] = dst, src
        flag = np.zeros(2 * E, dtype=np.int64)
        flag[1::2] = 1
        t2 = np.repeat(ts, 2)
        e2 = np.repeat(eids, 2)
        order = np.lexsort((np.arange(2 * E), t2, owner))
        self.indptr = np.zeros(n_nodes + 1, dtype=np.int64)
        np.cumsum(np.bincount(owner, minlength=n_nodes), out=self.indptr[1:])
        self.nbr = other[order]
        self.eid = e2[order]
        self.ts = t2[order]
        self.flag = flag[order]

    # a2  graph.py:44-53 (strict '<' via searchsorted side='left') + :117-127 (recent_edges)
    def find_recent(self, nids, ts, k: int):
        nids = np.asarray(nids)
        ts = np.asarray(ts)
        assert len(nids) == len(ts)
        n = len(nids)
        out_n = np.zeros((n, k), dtype=np.int64)
        out_e = np.zeros((n, k), dtype=np.int64)
        out_t = np.zeros((n, k), dtype=np.float32)
        out_d = np.zeros((n, k), dtype=np.int64)
        for i in range(n):
            lo, hi = self.indptr[nids[i]], self.indptr[nids[i] + 1]
            cut = lo + np.searchsorted(self.ts[lo:hi], ts[i], side='left')
            beg = max(lo, cut - k)
            m = cut - beg
            if m == 0:
                continue
            out_n[i, k - m:] = self.nbr[beg:cut]
            out_e[i, k - m:] = self.eid[beg:cut]
            out_t[i, k - m:] = self.ts[beg:cut]   # float64 -> float32 (graph.py:91,126)
            out_d[i, k - m:] = self.flag[beg:cut]
        return out_n, out_e, out_t, out_d

    # graph.py:150-155
    def get_history(self, nids, ts, hist_len: int):
        return self.find_recent(nids, ts, hist_len)


# --------------------------------------------------------------------------------------
# a8  select_latest_nids (tiger/model/utils.py:10-16) with torch_scatter CPU semantics
# --------------------------------------------------------------------------------------
def select_latest_scan(nids: np.ndarray, ts: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Normative definition: unique ids ascending; per id, the position of the maximum
    timestamp found by a left-to-right scan with a strict '>' update (lowest position
    wins ties)."""
    uniq, inv = np.unique(nids, return_inverse=True)
    best = np.full(len(uniq), -np.inf)
    arg = np.full(len(uniq), len(nids), dtype=np.int64)
    for pos in range(len(nids)):
        g = inv[pos]
        if ts[pos] > best[g]:
            best[g] = ts[pos]
            arg[g] = pos
    return uniq.astype(np.int64), arg


def select_latest(nids: np.ndarray, ts: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Vectorised equivalent of select_latest_scan (checked against it in the tests)."""
    nids = np.asarray(nids)
    ts = np.asarray(ts)
    if len(nids) == 0:
        return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
    pos = np.arange(len(nids))
    order = np.lexsort((pos, -ts.astype(np.float64), nids))  # id asc, ts desc, pos asc
    sorted_ids = nids[order]
    first = np.ones(len(nids), dtype=bool)
    first[1:] = sorted_ids[1:] != sorted_ids[:-1]
    return sorted_ids[first].astype(np.int64), order[first].astype(np.int64)


# a22  anonymized_reindex (tiger/model/utils.py:19-27)
def anonymized_reindex(hist_nids: np.ndarray) -> np.ndarray:
    out = np.zeros_like(hist_nids)
    for i, row in enumerate(hist_nids):
        rank: Dict[int, int] = OrderedDict()
        for v in row[::-1]:
            if v not in rank:
                rank[v] = len(rank) + 1
        out[i] = [rank[v] for v in row]
    out[hist_nids == 0] = 0
    return out


# --------------------------------------------------------------------------------------
# a3/a4  batch collation (tiger/data/data_loader.py:61-168, data_classes.py:150-165)
# --------------------------------------------------------------------------------------
class OracleBatch:
    """What GraphCollator.__call__ (data_loader.py:77-93) hands to the model."""
    pass


def collate(graph: OracleGraph, src, dst, neg, ts, eids, n_neighbors: int,
            restarter: Optional[str] = None, hist_len: int = 0) -> OracleBatch:
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    neg = np.asarray(neg, dtype=np.int64)
    ts = np.asarray(ts, dtype=np.float64)
    b = OracleBatch()
    b.src, b.dst, b.neg, b.eids = src, dst, neg, np.asarray(eids, dtype=np.int64)
    b.ts64 = ts
    b.ts = ts.astype(np.float32)                                # data_loader.py:92
    # collate_memory_nodes (data_loader.py:105-131), n_layers = 1
    batch_nids = np.concatenate([src, dst, neg])
    nn_, ne_, nt_, _ = graph.find_recent(batch_nids, np.tile(ts, 3), n_neighbors)
    b.batch_nids = batch_nids
    b.neigh_nids, b.neigh_eids, b.neigh_ts = nn_, ne_, nt_
    b.involved = np.unique(np.concatenate([batch_nids, nn_.ravel()]))   # sorted (:121)
    # ComputationGraph.local_index (data_classes.py:163-165)
    b.local_index = np.zeros(graph.num_node, dtype=np.int64)
    b.local_index[b.involved] = np.arange(len(b.involved))
    # collate_hit_data (data_loader.py:61-75)

    def hits(center, target):
        neigh, *_ = graph.find_recent(target, ts, n_neighbors)
        return (center[:, None] == neigh).astype(np.float32)
    b.src_hits = hits(src, dst)
    b.dst_hits = hits(dst, src)
    b.neg_src_hits = hits(src, neg)
    b.neg_dst_hits = hits(neg, src)
    # collate_restart_data (data_loader.py:95-168); select_latest on float64 times
    b.restart = None
    if restarter is not None:
        pos = np.concatenate([src, dst])
        ts2 = np.tile(ts, 2)
        uniq, index = select_latest(pos, ts2)
        r = OracleBatch()
        r.index, r.nids, r.ts64 = index, uniq, ts2[index]
        r.ts = r.ts64.astype(np.float32)
        if restarter == 'seq':
            r.hist_nids, r.hist_eids, r.hist_ts, r.hist_dirs = graph.get_history(uniq, r.ts64, hist_len)
            r.anonymized_ids = anonymized_reindex(r.hist_nids)
        elif restarter == 'static':
            _, _, prev_ts, _ = graph.get_history(uniq, r.ts64, 1)
            r.prev_ts = prev_ts                                  # [n, 1] float32 (:161-165)
        else:
            raise NotImplementedError(restarter)
        b.restart = r
    return b


# --------------------------------------------------------------------------------------
# floating-point operators
# --------------------------------------------------------------------------------------
def time_encode(ts: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a10  TimeEncode.forward (tiger/model/time_encoding.py:16-27)."""
    return torch.cos(ts.unsqueeze(-1) * w + b)


def gru_cell(x, h, w_ih, w_hh, b_ih, b_hh):
    """a13  nn.GRUCell as called at tiger/model/update_modules.py:36 (gate order r,z,n)."""
    gi = x @ w_ih.t() + b_ih
    gh = h @ w_hh.t() + b_hh
    d = h.shape[1]
    r = torch.sigmoid(gi[:, :d] + gh[:, :d])
    z = torch.sigmoid(gi[:, d:2 * d] + gh[:, d:2 * d])
    n = torch.tanh(gi[:, 2 * d:] + r * gh[:, 2 * d:])
    return (h - n) * z + n


def merge_layer(x1, x2, fc1_w, fc1_b, fc2_w, fc2_b):
    """MergeLayer.forward in eval mode (tiger/model/basic_modules.py:16-19)."""
    h = torch.relu(torch.cat([x1, x2], -1) @ fc1_w.t() + fc1_b)
    return h @ fc2_w.t() + fc2_b


def _mha(query, key, value, wq, wk, wv, in_bias, wo, bo, key_padding_mask, n_head):
    """torch F.multi_head_attention_forward, need_weights=True branch, eval mode, inputs
    [L, n, E] / [S, n, C] (seq-first); q is scaled by sqrt(1/head_dim) before q.k^T and
    masked keys receive -inf before the softmax."""
    L, n, E = query.shape
    S = key.shape[0]
    bq, bk, bv = in_bias.chunk(3)
    q = query @ wq.t() + bq
    k = key @ wk.t() + bk
    v = value @ wv.t() + bv
    hd = E // n_head
    q = q.reshape(L, n * n_head, hd).transpose(0, 1) * math.sqrt(1.0 / hd)
    k = k.reshape(S, n * n_head, hd).transpose(0, 1)
    v = v.reshape(S, n * n_head, hd).transpose(0, 1)
    mask = torch.zeros(n, S).masked_fill(key_padding_mask, float('-inf'))
    mask = mask.view(n, 1, 1, S).expand(-1, n_head, -1, -1).reshape(n * n_head, 1, S)
    p = torch.softmax(torch.baddbmm(mask, q, k.transpose(1, 2)), dim=-1)
    o = torch.bmm(p, v).transpose(0, 1).contiguous().view(L * n, E)
    return (o @ wo.t() + bo).view(L, n, E)


def temporal_attention(W: Dict[str, torch.Tensor], prefix: str, n_head: int,
                       qx, qt, kx, ky, kt, padding_mask):
    """a16  TemporalAttention.forward (tiger/model/temporal_agg_modules.py:210-235)."""
    query = torch.cat([qx, qt], 1).unsqueeze(0)
    kv = torch.cat([kx, ky, kt], 2).transpose(0, 1)
    padding_mask = padding_mask.clone()
    invalid = padding_mask.all(1, keepdim=True)
    padding_mask[invalid.squeeze(1), -1] = False
    m = prefix + 'mha_fn.'
    h = _mha(query, kv, kv, W[m + 'q_proj_weight'], W[m + 'k_proj_weight'], W[m + 'v_proj_weight'],
             W[m + 'in_proj_bias'], W[m + 'out_proj.weight'], W[m + 'out_proj.bias'],
             padding_mask, n_head).squeeze(0)
    h = h.masked_fill(invalid, 0.)
    g = prefix + 'merger.'
    return merge_layer(h, qx, W[g + 'fc1.weight'], W[g + 'fc1.bias'], W[g + 'fc2.weight'], W[g + 'fc2.bias'])


# --------------------------------------------------------------------------------------
# the model state + per-batch path
# --------------------------------------------------------------------------------------
class OracleTIGER:
    """State and per-batch forward path of TIGE/TIGER in eval mode.

    ``W`` uses the reference's state_dict key names.
    """

    def __init__(self, W: Dict[str, torch.Tensor], graph: OracleGraph, n_nodes: int, dim: int,
                 efeats: Optional[np.ndarray], nfeats: Optional[np.ndarray] = None, *,
                 n_neighbors: int = 10, n_head: int = 2, msg_src: str = 'left', upd_src: str = 'right',
                 restarter: str = 'seq', hist_len: int = 40, hit_type: str = 'bin'):
        if msg_src not in ('left', 'right') or upd_src not in ('left', 'right'):
            raise ValueError('msg_src/upd_src')                 # tiger.py:156-160
        self.W = {k: torch.as_tensor(v) for k, v in W.items()}
        self.graph = graph
        self.N = n_nodes
        self.efeats = None if efeats is None else torch.as_tensor(efeats, dtype=torch.float32)
        self.nfeats = None if nfeats is None else torch.as_tensor(nfeats, dtype=torch.float32)
        # feature_getter.py:76-77
        self.d = self.nfeats.shape[1] if self.nfeats is not None else dim
        self.de = self.efeats.shape[1] if self.efeats is not None else dim
        self.M = 3 * self.d + self.de                            # tiger.py:62
        self.K = n_neighbors
        self.n_head = n_head
        self.msg_src, self.upd_src = msg_src, upd_src
        self.restarter, self.hist_len, self.hit_type = restarter, hist_len, hit_type
        self.reset()

    # ---- state ----
    def reset(self):                                             # tiger.py:457-463
        N, d = self.N, self.d
        self.left_vals = torch.zeros(N, d)
        self.left_ts = torch.zeros(N)
        self.right_vals = torch.zeros(N, d)
        self.right_ts = torch.zeros(N)
        self.msg_vals = torch.zeros(N, self.M)
        self.msg_ts = torch.zeros(N)
        self.has_msg = np.zeros(N, dtype=bool)                   # memory.py:68 (python set)

    def _mem(self, which):
        return (self.left_vals, self.left_ts) if which == 'left' else (self.right_vals, self.right_ts)

    # a25  feature_getter.py:80-106
    def nf(self, nids: torch.Tensor) -> torch.Tensor:
        if self.nfeats is None:
            return torch.zeros(*nids.shape, self.d)
        return self.nfeats[nids]

    def ef(self, eids: torch.Tensor) -> torch.Tensor:
        if self.efeats is None:
            return torch.zeros(*eids.shape, self.de)
        return self.efeats[eids]

    # ---- a12/a13: steps 1-2 (tiger.py:206-221, 292-356) ----
    def updated_reprs(self, involved: np.ndarray):
        outdated = involved[self.has_msg[involved]]             # ascending
        reprs = self.right_vals[torch.from_numpy(involved)].clone()
        h_new = None
        if len(outdated):
            o = torch.from_numpy(outdated)
            msgs, msg_ts = self.msg_vals[o], self.msg_ts[o]     # message_modules.py:152-160
            if (self._mem(self.msg_src)[1][o] > msg_ts).any():
                raise ValueError('Messages happened later than memory updating.')
            if self.msg_src == 'left' and not (msg_ts == self.left_ts[o]).all():
                raise ValueError("Messages' ts should be equal to last update ts")  # tiger.py:325
            W = self.W
            h_new = gru_cell(msgs, self._mem(self.upd_src)[0][o],
                             W['right_mem_updater.cell.weight_ih'], W['right_mem_updater.cell.weight_hh'],
                             W['right_mem_updater.cell.bias_ih'], W['right_mem_updater.cell.bias_hh'])
            reprs[torch.from_numpy(np.searchsorted(involved, outdated))] = h_new
        return outdated, h_new, reprs

    # ---- a15/a16: step 3 (temporal_agg_modules.py:29-83) ----
    def embed(self, reprs, b: OracleBatch):
        W = self.W
        w, ph = W['time_encoder.basis_freq'], W['time_encoder.phase']
        li = torch.from_numpy(b.local_index)
        center = torch.from_numpy(b.batch_nids)
        neigh = torch.from_numpy(b.neigh_nids)
        t3 = torch.from_numpy(b.ts).repeat(3)
        c = reprs[li[center]] + self.nf(center)
        x = reprs[li[neigh.flatten()]] + self.nf(neigh.flatten())
        x = x.reshape(neigh.shape[0], neigh.shape[1], self.d)
        e = self.ef(torch.from_numpy(b.neigh_eids))
        dt = t3[:, None] - torch.from_numpy(b.neigh_ts)
        kt = time_encode(dt, w, ph)
        qt = time_encode(torch.zeros_like(dt[:, 0]), w, ph)
        return temporal_attention(W, 'temporal_embedding_fn.fns.0.', self.n_head,
                                  c, qt, x, e, kt, neigh == 0)

    # ---- a9: step 5 (tiger.py:422-442, memory.py:77-106) ----
    def store_events(self, b: OracleBatch):
        W = self.W
        src, dst = torch.from_numpy(b.src), torch.from_numpy(b.dst)
        ts = torch.from_numpy(b.ts)
        vals, upd_ts = self._mem(self.msg_src)
        sp, dp = upd_ts[src], upd_ts[dst]
        if (sp > ts).any() or (dp > ts).any():
            raise ValueError('Events occur before the udpated memory.')
        pos = np.concatenate([b.src, b.dst])
        if self.has_msg[pos].any():
            raise ValueError('Node has unused messages.')        # memory.py:85-87
        sv = vals[src] + self.nf(src)
        dv = vals[dst] + self.nf(dst)
        ev = self.ef(torch.from_numpy(b.eids))
        w, ph = W['time_encoder.basis_freq'], W['time_encoder.phase']
        rows = torch.cat([torch.cat([sv, dv, ev, time_encode(ts - sp, w, ph)], 1),
                          torch.cat([dv, sv, ev, time_encode(ts - dp, w, ph)], 1)], 0)
        uniq, index = select_latest(pos, np.tile(b.ts, 2))
        u = torch.from_numpy(uniq)
        self.msg_vals[u] = rows[torch.from_numpy(index)]
        self.msg_ts[u] = ts.repeat(2)[torch.from_numpy(index)]
        self.has_msg[uniq] = True
        return uniq, index

    # ---- the whole of TIGE.contrast_learning (tiger.py:174-290) ----
    def contrast_step(self, b: OracleBatch) -> Dict[str, object]:
        B = len(b.src)
        W = self.W
        out: Dict[str, object] = {}
        outdated, h_new, reprs = self.updated_reprs(b.involved)           # steps 1-2
        out['outdated'], out['h_new'], out['reprs'] = outdated, h_new, reprs
        z = self.embed(reprs, b)                                           # step 3
        out['h_left_with_negs'] = z
        pos = np.concatenate([b.src, b.dst])
        ts2 = np.tile(b.ts, 2)
        if len(outdated):                                                  # step 4 (:230-241)
            uniq, _ = select_latest(pos, ts2)
            hit = uniq[self.has_msg[uniq]]          # positives are involved => outdated
            if len(hit):
                rows = torch.from_numpy(np.searchsorted(outdated, hit))
                h = torch.from_numpy(hit)
                new_ts = self.msg_ts[h]
                if (self.right_ts[h] > new_ts).any():
                    raise ValueError('You are not allowed to modify past memory.')
                self.has_msg[hit] = False
                self.right_ts[h] = new_ts
                self.right_vals[h] = h_new[rows]
            out['right_written'] = hit
        uniq_s, index_s = self.store_events(b)                             # step 5
        out['msg_nodes'], out['msg_index'] = uniq_s, index_s
        p = torch.from_numpy(pos)
        out['h_prev_left'] = self.left_vals[p].clone()                     # :248-251
        out['h_prev_right'] = self.right_vals[p].clone()
        uniq, index = select_latest(pos, ts2)                              # step 6 (:408-420)
        u, ix = torch.from_numpy(uniq), torch.from_numpy(index)
        new_ts = torch.from_numpy(ts2)[ix]
        if (self.left_ts[u] > new_ts).any():
            raise ValueError('You are not allowed to modify past memory.')
        self.left_ts[u] = new_ts
        self.left_vals[u] = z[:2 * B][ix]
        out['pos_unique'], out['pos_index'] = uniq, index
        x, y, ny = z.reshape(3, B, self.d)                                 # step 7 (:259-288)
        if self.hit_type == 'bin':
            emb = W['hit_embedding.weight']

            def he(hits):
                return emb[torch.from_numpy(hits).max(1).values.long()]
            xp, yp = x + he(b.src_hits), y + he(b.dst_hits)
            xn, yn = x + he(b.neg_src_hits), ny + he(b.neg_dst_hits)
        elif self.hit_type == 'none':
            xp = xn = x
            yp, yn = y, ny
        else:
            raise NotImplementedError(self.hit_type)
        s = 'score_fn.'
        sw = (W[s + 'fc1.weight'], W[s + 'fc1.bias'], W[s + 'fc2.weight'], W[s + 'fc2.bias'])
        ps = merge_layer(xp, yp, *sw).squeeze(1)
        ns = merge_layer(xn, yn, *sw).squeeze(1)
        labels = torch.cat([torch.ones_like(ps), torch.zeros_like(ns)])
        out['loss'] = F.binary_cross_entropy_with_logits(torch.cat([ps, ns]), labels)
        out['pos_scores'], out['neg_scores'], out['h_left'] = ps, ns, z[:2 * B]
        return out

    # ---- a21: SeqRestarter.forward (restarters.py:51-114) ----
    def seq_restarter(self, nids: np.ndarray, hist_nids, hist_eids, hist_ts, hist_dirs, anonymized_ids):
        W = self.W
        p = 'restarter_fn.'
        d, de = self.d, self.de
        nid_t = torch.from_numpy(np.asarray(nids, dtype=np.int64))
        hn, he_ = torch.from_numpy(hist_nids), torch.from_numpy(hist_eids)
        ht, hd = torch.from_numpy(hist_ts), torch.from_numpy(hist_dirs)
        an = torch.from_numpy(anonymized_ids)
        n, L = hn.shape
        mask = hn == 0
        mask[:, -1] = False                                   # before the all() test: SURVEY Q5
        invalid = mask.all(1, keepdim=True)
        r = nid_t.unsqueeze(1).repeat(1, L)
        s_ids = r * hd + hn * (1 - hd)                        # restarters.py:93-94
        d_ids = r * (1 - hd) + hn * hd
        tv = time_encode(ht[:, -1].unsqueeze(1) - ht, W[p + 'time_encoder.basis_freq'],
                         W[p + 'time_encoder.phase'])
        full = torch.cat([self.nf(s_ids), self.nf(d_ids), W[p + 'anony_emb.weight'][an], self.ef(he_), tv], 2)
        dm = full.shape[2]
        full[:, -1, :dm - d] = 0.                             # :104; last_event_feat is this view => zeros (Q13)
        last_event_feat = torch.zeros(n, dm - d)
        qkv = full.transpose(0, 1)
        wi = W[p + 'mha_fn.in_proj_weight']
        o = _mha(qkv, qkv, qkv, wi[:dm], wi[dm:2 * dm], wi[2 * dm:], W[p + 'mha_fn.in_proj_bias'],
                 W[p + 'mha_fn.out_proj.weight'], W[p + 'mha_fn.out_proj.bias'], mask, self.n_head)
        h_left = torch.relu(o.mean(0)) @ W[p + 'out_fn.weight'].t() + W[p + 'out_fn.bias']
        h_right = merge_layer(h_left, last_event_feat, W[p + 'merger.fc1.weight'], W[p + 'merger.fc1.bias'],
                              W[p + 'merger.fc2.weight'], W[p + 'merger.fc2.bias'])
        h_left = h_left.masked_fill(invalid, 0.)
        h_right = h_right.masked_fill(invalid, 0.)
        return h_left, h_right, ht[:, -1]

    # ---- a23: StaticRestarter.forward (restarters.py:262-277) ----
    def static_restarter(self, nids: np.ndarray, prev_ts: np.ndarray):
        n = torch.from_numpy(np.asarray(nids, dtype=np.int64))
        return (self.W['restarter_fn.left_emb.weight'][n], self.W['restarter_fn.right_emb.weight'][n],
                torch.from_numpy(np.asarray(prev_ts, dtype=np.float32)))

    def restarter_forward(self, nids: np.ndarray, ts: np.ndarray):
        """Restarter called with computation_graph=None: history looked up at ``ts``."""
        if self.restarter == 'seq':
            hn, he_, ht, hd = self.graph.get_history(nids, ts, self.hist_len)
            return self.seq_restarter(nids, hn, he_, ht, hd, anonymized_reindex(hn))
        if self.restarter == 'static':
            _, _, pt, _ = self.graph.get_history(nids, ts, 1)
            return self.static_restarter(nids, pt[:, 0])
        raise NotImplementedError(self.restarter)

    def restarter_on_batch(self, b: OracleBatch):
        """Restarter called with the collated restart_data (training targets, tiger.py:576-581)."""
        r = b.restart
        if self.restarter == 'seq':
            return self.seq_restarter(r.nids, r.hist_nids, r.hist_eids, r.hist_ts, r.hist_dirs, r.anonymized_ids)
        # StaticRestartData.prev_ts keeps its [n,1] shape (data_loader.py:161-167)
        n = torch.from_numpy(r.nids)
        return (self.W['restarter_fn.left_emb.weight'][n], self.W['restarter_fn.right_emb.weight'][n],
                torch.from_numpy(r.prev_ts))

    # ---- a24: TIGER.restart (tiger.py:594-609) ----
    def restart(self, nids: np.ndarray, ts: np.ndarray):
        nids = np.asarray(nids, dtype=np.int64)
        if len(nids) == 0:
            return
        self.has_msg[nids] = False                           # memory.py:136 (only the set changes, Q1)
        hl, hr, pt = self.restarter_forward(nids, ts)
        n = torch.from_numpy(nids)
        self.left_ts[n], self.left_vals[n] = pt, hl
        self.right_ts[n], self.right_vals[n] = pt, hr

    # ---- a20: mutual loss (tiger.py:547-592) ----
    def mutual_loss(self, b: OracleBatch, step_out: Dict[str, object]):
        ix = torch.from_numpy(b.restart.index)
        sl, sr, _ = self.restarter_on_batch(b)
        targets = torch.cat([step_out['h_prev_left'][ix], step_out['h_prev_right'][ix]], 0)
        preds = torch.cat([sl, sr], 0)
        valid = torch.where(~(targets == 0).all(1))[0]
        if len(valid):
            return F.mse_loss(preds[valid], targets[valid])
        return torch.tensor(0.)

    # ---- a27: flush_msg (tiger.py:444-455) ----
    def flush_msg(self):
        nodes = np.nonzero(self.has_msg)[0]
        if len(nodes) == 0:
            return
        outdated, h_new, _ = self.updated_reprs(nodes)
        o = torch.from_numpy(outdated)
        new_ts = self.msg_ts[o]
        if (self.right_ts[o] > new_ts).any():
            raise ValueError('You are not allowed to modify past memory.')
        self.right_ts[o] = new_ts
        self.right_vals[o] = h_new
        self.has_msg[outdated] = False


# a26  ChunkSampler (tiger/data/data_loader.py:17-40)
def chunk_range(n: int, rank: int, world_size: int, bs: int, seed: int = 0, epoch: int = 0) -> Tuple[int, int]:
    g = torch.Generator()
    g.manual_seed(seed + epoch)
    residual = n % (world_size * bs)
    shift = int(torch.randint(0, residual + 1, size=(), generator=g))
    length = n // (world_size * bs) * bs
    lo = shift + length * rank
    return lo, lo + length


def lazy_restart_nodes(involved: np.ndarray, uptodate: np.ndarray) -> np.ndarray:
    """restart_nodes = set(involved) - uptodate (train_self_supervised.py:158-163,
    eval_utils.py:37-42); ``uptodate`` is a bool[N] flag table updated in place."""
    r = involved[~uptodate[involved]]
    uptodate[r] = True
    return r
