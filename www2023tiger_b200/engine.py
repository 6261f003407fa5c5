"""The per-batch temporal memory path as one fixed, sync-free launch sequence.

`TigerEngine` owns the device state of TIGE/TIGER (both memories, the last-message store, the
pending / up-to-date flags) and runs TIGE.contrast_learning (reference tiger/model/tiger.py:174-290)
as a fixed sequence of kernel launches (13 per batch) whose data-dependent sizes (involved U, outdated O, restarted R) stay on
the device.  `StreamRunner` adds the device neighbor finder in front, captures the sequence in a
CUDA graph and feeds it from pinned host buffers.

Launch sequence of one batch (B events, K neighbors):
  1 tiger_find_recent        3B queries -> neighbor tables, involved bitmap, float32 batch times
  2 tiger_compact_involved   involved / outdated / restart lists + gru_row map
  3 tiger_static_restart     (lazy-restart mode only) re-initialise not-yet-seen nodes
  4 tiger_gru_update         h(t'+) for the O outdated nodes      (steps 1-2 of the reference)
  5 tiger_temporal_attention h(t-) for [src;dst;neg]              (step 3)
  6 tiger_select_latest      argmax-by-timestamp winners of [src;dst]
  7 tiger_right_writeback    persist h(t'+) of outdated positives (step 4) [+ restarter targets]
  8 tiger_store_messages     build + store the new raw messages   (step 5)
  9 tiger_left_writeback     persist h(t-) of positives           (step 6)
 10 tiger_link_score         hit flags, scorer, BCE loss          (step 7)
"""
from typing import Dict, Optional

import os

import numpy as np
import torch
from torch import Tensor

from . import _lib, ops
from ._lib import raise_on_err_flags
from .ops import f32, f64, i32, i64, u8

KERNELS_PER_STEP = {  # launches of our kernels per batch (for the bench `gpu_launches` field)
    'find_recent': 1, 'compact_involved': 1, 'gru_update': 1, 'temporal_attention': 4, 'select_latest': 1,
    'right_writeback': 1, 'store_messages': 2, 'left_writeback': 0, 'link_score': 1,
}


class TigerEngine:
    def __init__(self, weights: Dict[str, Tensor], csr: ops.DeviceCSR, *, n_nodes: int, dim: int,
                 efeats: Optional[Tensor], nfeats: Optional[Tensor] = None, n_neighbors: int = 10,
                 n_head: int = 2, batch_size: int = 200, msg_src: str = 'left', upd_src: str = 'right',
                 restarter: Optional[str] = None, hist_len: int = 40, hit_type: str = 'bin',
                 lazy_restart: bool = False, want_restarter_targets: bool = False, device='cuda'):
        if msg_src not in ('left', 'right') or upd_src not in ('left', 'right'):
            raise ValueError(f'Invalid msg_src={msg_src} / upd_src={upd_src}')   # tiger.py:156-160
        if hit_type not in ('bin', 'none'):
            raise NotImplementedError(f'hit_type={hit_type}')
        if lazy_restart and restarter not in ('static', 'seq'):
            raise NotImplementedError(f'lazy restart needs a seq or static restarter, got {restarter!r}')
        dev = torch.device(device)
        self.device, self.csr = dev, csr
        self.N, self.K, self.H, self.B = n_nodes, n_neighbors, n_head, batch_size
        self.efeats = efeats
        self.nfeats = nfeats
        self.d = nfeats.shape[1] if nfeats is not None else dim              # feature_getter.py:76-77
        self.de = efeats.shape[1] if efeats is not None else dim
        self.M = 3 * self.d + self.de                                          # tiger.py:62
        self.msg_src, self.upd_src = msg_src, upd_src
        self.restarter, self.hit_type, self.lazy_restart = restarter, hit_type, lazy_restart
        self.want_targets = want_restarter_targets
        N, d, B, K = self.N, self.d, self.B, self.K
        z = lambda *s, dt=f32: torch.zeros(*s, dtype=dt, device=dev)
        # ---- persistent state (reference: Memory x2, MessageStoreNoGradLastOnly) ----
        self.left_vals, self.left_ts, self.left_active = z(N, d), z(N), z(N, dt=u8)
        self.right_vals, self.right_ts, self.right_active = z(N, d), z(N), z(N, dt=u8)
        self.msg_vals, self.msg_ts, self.has_msg = z(N, self.M), z(N), z(N, dt=u8)
        self.uptodate = z(N, dt=u8)
        # ---- per-batch scratch (fixed capacity) ----
        self.cap = 3 * B * (K + 1)
        self.bitmap = z(ops.bitmap_words(N), dt=i32)
        self.involved, self.outdated, self.restart_nodes = z(self.cap, dt=i64), z(self.cap, dt=i64), z(self.cap, dt=i64)
        self.gru_row = torch.full((N,), -1, dtype=i32, device=dev)
        self.counts = z(4, dt=i32)
        self.err_flags = z(1, dt=i32)
        self.neigh_nids, self.neigh_eids = z(3 * B, K, dt=i64), z(3 * B, K, dt=i64)
        self.neigh_ts = z(3 * B, K)
        self.ts32 = z(B)
        self.h_new = z(self.cap, d)
        self.emb = z(3 * B, d)
        self.winner = z(2 * B, dt=u8)
        self.sel_count = z(1, dt=i32)
        self.out_buf = z(2 * B + 1)                    # [pos scores | neg scores | loss]
        self.hprev_left = z(2 * B, d) if want_restarter_targets else None
        self.hprev_right = z(2 * B, d) if want_restarter_targets else None
        # ---- batch inputs: [src | dst | neg | eids] int64 and ts float64, one contiguous buffer ----
        self.inp = z(5 * B, dt=i64)
        self.bind_io(self.inp, self.out_buf)
        # ---- parameters ----
        self._side = torch.cuda.Stream(device=dev)
        self._serial = int(os.environ.get('TIGER_SERIAL', '0'))   # debug aid, bit mask: branches kept on the main stream
        self._ev_fork, self._ev_side = torch.cuda.Event(), torch.cuda.Event()
        self._ev_fork0, self._ev_rst = torch.cuda.Event(), torch.cuda.Event()
        self.gru_pack = None
        self.attn_pack = ops.AttnPack(self.d, self.de, dev, n_head)
        self.score_pack = ops.ScorePack(self.d, dev)
        self.score_fold = ops.ScoreFold(self.d, dev)
        self.pq = torch.zeros(3 * batch_size, 2 * self.d, dtype=f32, device=dev)   # [W1a z | W1b z] per query row
        self.hist_len = hist_len
        self.seq = ops.SeqRestarterOp(self.d, self.de, hist_len, n_head, self.cap, dev) if restarter == 'seq' else None
        self.load_weights(weights)

    # ------------------------------------------------------------------ parameters
    def load_weights(self, W: Dict[str, Tensor]):
        """(Re)pack parameters given under the reference's state_dict names."""
        g = lambda k: W[k].detach().to(self.device, f32).contiguous()
        c = 'right_mem_updater.cell.'
        args = (g(c + 'weight_ih'), g(c + 'weight_hh'), g(c + 'bias_ih'), g(c + 'bias_hh'))
        if self.gru_pack is None:
            self.gru_pack = ops.GruPack(*args)
        else:
            self.gru_pack.refresh(*args)
        a = 'temporal_embedding_fn.fns.0.'
        self.time_w, self.time_b = g('time_encoder.basis_freq'), g('time_encoder.phase')
        self.attn_pack.refresh(g(a + 'mha_fn.q_proj_weight'), g(a + 'mha_fn.k_proj_weight'),
                               g(a + 'mha_fn.v_proj_weight'), g(a + 'mha_fn.in_proj_bias'),
                               g(a + 'mha_fn.out_proj.weight'), g(a + 'mha_fn.out_proj.bias'),
                               g(a + 'merger.fc1.weight'), g(a + 'merger.fc1.bias'),
                               g(a + 'merger.fc2.weight'), g(a + 'merger.fc2.bias'), self.time_w, self.time_b)
        s = 'score_fn.'
        self.score_pack.refresh(g(s + 'fc1.weight'), g(s + 'fc1.bias'), g(s + 'fc2.weight'), g(s + 'fc2.bias'),
                                g('hit_embedding.weight') if self.hit_type == 'bin' else None)
        self.score_fold.refresh(g(s + 'fc1.weight'), g(s + 'fc1.bias'), g(s + 'fc2.weight'), g(s + 'fc2.bias'),
                                g(a + 'merger.fc2.weight'), g(a + 'merger.fc2.bias'),
                                g('hit_embedding.weight') if self.hit_type == 'bin' else None)
        self.attn_pack.attach_score_fold(self.score_fold, self.pq)
        if self.restarter == 'static':
            self.left_emb = g('restarter_fn.left_emb.weight')
            self.right_emb = g('restarter_fn.right_emb.weight')
        elif self.restarter == 'seq':
            self.seq.set_weights(W)

    # ------------------------------------------------------------------ state
    def reset(self):
        """TIGE.reset (tiger.py:457-463) + a fresh up-to-date set."""
        for t in (self.left_vals, self.left_ts, self.left_active, self.right_vals, self.right_ts,
                  self.right_active, self.has_msg, self.uptodate, self.err_flags):
            t.zero_()

    def clear_messages(self):
        """msg_store.clear() + uptodate_nodes = set() of the restart trigger
        (train_self_supervised.py:153-156)."""
        self.has_msg.zero_()
        self.uptodate.zero_()

    def _mem(self, which):
        if which == 'left':
            return self.left_vals, self.left_ts
        return self.right_vals, self.right_ts

    def check_errors(self):
        """Host sync: raise the reference's ValueError if an invariant kernel flag is set."""
        raise_on_err_flags(int(self.err_flags.item()) & 0xffffffff)

    def set_batch(self, src, dst, neg, ts, eids):
        """Copy one batch (numpy or tensors; ts float64) into the device input buffer."""
        B = self.B
        host = torch.empty(5 * B, dtype=i64)
        h = host.numpy()
        h[:B], h[B:2 * B], h[2 * B:3 * B], h[3 * B:4 * B] = src, dst, neg, eids
        h[4 * B:].view(np.float64)[:] = ts
        self.inp.copy_(host)

    # ------------------------------------------------------------------ the launch sequence
    def launch_finder(self):
        ops.find_recent(self.csr, self.batch_nids, self.ts64, self.K, ts_period=self.B, want_dirs=False,
                        ts32_out=self.ts32, bitmap=self.bitmap,
                        out=(self.neigh_nids, self.neigh_eids, self.neigh_ts, None))

    def launch_model(self, with_scorer: bool = True):
        d, B = self.d, self.B
        # Branches (restarter beside the GRU, write-back / message store beside the attention chain) exist only under
        # graph capture, where they become parallel paths of the graph with explicit edges; eager launches (tests,
        # smoke, debugging) stay on one stream.  TIGER_SERIAL is a debug bit mask (1: restarter, 2: write-back /
        # message branch) that keeps a branch on the main stream under capture as well; TIGER_EAGER_BRANCHES=1
        # re-enables the branches for eager launches (tools/eager_race.py).
        cur = torch.cuda.current_stream()
        serial = self._serial if (torch.cuda.is_current_stream_capturing() or os.environ.get('TIGER_EAGER_BRANCHES') == '1') else 7
        side0 = cur if serial & 1 else self._side
        side1 = cur if serial & 2 else self._side
        msg_vals_mem, msg_ts_mem = self._mem(self.msg_src)
        upd_vals, _ = self._mem(self.upd_src)
        ops.compact_involved(self.bitmap, self.N, self.involved, self.counts, has_msg=self.has_msg,
                             uptodate=self.uptodate if self.lazy_restart else None, outdated=self.outdated,
                             gru_row=self.gru_row, restart_nodes=self.restart_nodes if self.lazy_restart else None,
                             err_flags=self.err_flags)
        # Fork 0: the restart only writes rows of nodes compact_involved put on the restart list, which it also
        # removed from the outdated list (csrc/graph.cu: pend = member && has_msg && !rst), i.e. rows the GRU
        # neither reads nor writes - the restarter runs next to the GRU on the side stream; the attention and
        # the message builder (same side stream, later) read restarted rows and wait for it.
        main = torch.cuda.current_stream()
        if self.lazy_restart:
            self._ev_fork0.record(main)
            with torch.cuda.stream(side0):
                side0.wait_event(self._ev_fork0)
                if self.restarter == 'static':
                    ops.static_restart(self.restart_nodes, self.cap, self.csr, self.left_emb, self.right_emb, d,
                                       count=self.counts[2:], batch_ts=self.ts32, left_vals=self.left_vals,
                                       left_ts=self.left_ts, left_active=self.left_active, right_vals=self.right_vals,
                                       right_ts=self.right_ts, right_active=self.right_active, has_msg=self.has_msg)
                else:
                    # TIGER.restart with the seq restarter (tiger.py:594-609, restarters.py:51-114): history at
                    # ts.min() of the batch, surrogate h(t'-) / h(t'+), both memories overwritten without checks
                    # (compact_involved has already dropped the pending-message flags of these nodes)
                    R = self.counts[2:]
                    ops.min_time(self.ts32, self.seq.tmin)
                    self.seq.history(self.csr, self.restart_nodes, self.seq.tmin, self.cap, ts_period=1, count=R)
                    hl, hr, pt = self.seq.forward(self.restart_nodes, self.cap, self.nfeats, self.efeats, count=R)
                    ops.scatter_rows(self.left_vals, self.restart_nodes, hl, ts_table=self.left_ts, ts=pt,
                                     active=self.left_active, count=R)
                    ops.scatter_rows(self.right_vals, self.restart_nodes, hr, ts_table=self.right_ts, ts=pt,
                                     active=self.right_active, count=R)
                self._ev_rst.record(side0)
        ops.gru_update(self.gru_pack, node_ids=self.outdated, x_table=self.msg_vals, h_table=upd_vals,
                       n_rows=self.cap, out=self.h_new, count=self.counts[1:], msg_ts=self.msg_ts,
                       check_mem_ts=msg_ts_mem, check_equal=(self.msg_src == 'left'), err_flags=self.err_flags)
        # Fork: the write-back / message branch (argmax-by-timestamp selection -> right write-back -> message
        # build + store) only needs the GRU output and the batch, and touches rows the attention never reads
        # (right-memory rows of nodes WITH a pending message are read from h_new, csrc/attention.cu
        # resolve_row), so it runs on a side stream next to the attention chain.  Captured in a CUDA graph
        # the two branches become parallel paths of the graph.
        if self.lazy_restart and (serial & 2):
            main.wait_event(self._ev_rst)
        self._ev_fork.record(main)
        with torch.cuda.stream(side1):
            side1.wait_event(self._ev_fork)
            ops.select_latest(self.pos, self.ts32, want_unique=False, winner=self.winner, want_count=False)
            ops.right_writeback(self.pos, self.winner, self.gru_row, self.h_new, d, self.right_vals, self.right_ts,
                                self.right_active, self.msg_ts, self.has_msg, self.left_vals, self.hprev_left,
                                self.hprev_right, self.err_flags)
            ops.store_messages(self.src, self.dst, self.eids, self.ts32, self.winner, msg_vals_mem, msg_ts_mem,
                               self.nfeats, self.efeats, d, self.de, self.time_w, self.time_b, self.msg_vals,
                               self.msg_ts, self.has_msg, self.err_flags)
            self._ev_side.record(side1)
        if self.lazy_restart:
            main.wait_event(self._ev_rst)
        # update_left_memory is fused into the last attention product, which first waits for the side branch (the
        # message builder reads the OLD left-memory rows, and `winner` comes from the selection kernel)
        self.attn_pack.attach_left_writeback(self.pos, self.winner, self.ts32, self.left_vals, self.left_ts,
                                             self.left_active, self.err_flags, ready_event=self._ev_side)
        ops.temporal_attention(self.attn_pack, self.H, self.batch_nids, self.ts32, self.neigh_nids, self.neigh_eids,
                               self.neigh_ts, rows_a=self.right_vals, rows_b=self.h_new, sel=self.gru_row,
                               nfeats=self.nfeats, efeats=self.efeats, out=self.emb)
        # the link scorer reads the projections `pq` of the embeddings and changes no state: in the pipelined replay
        # it is a separate graph on the copy-out stream (launch_scorer), beside the next batch's model kernels
        if with_scorer:
            self.launch_scorer()

    def launch_scorer(self):
        ops.link_score_folded(self.score_fold, self.pq, self.src, self.dst, self.neg,
                              self.neigh_nids if self.hit_type == 'bin' else None, self.scores, self.loss)

    def bind_pq(self, pq: Tensor):
        """The [3B, 2d] first-layer projections of the link scorer, written by the last attention product and read
        by the scorer: one buffer per pipeline slot, because the scorer of batch i runs beside batch i+1."""
        assert pq.shape == self.pq.shape and pq.dtype == f32
        self.pq = pq
        self.attn_pack.attach_score_fold(self.score_fold, pq)

    def bind_io(self, inp: Tensor, out_buf: Tensor):
        """Points the batch-input views ([src | dst | neg | eids | ts as f64 bits], int64 [5B]) and the result
        views ([pos scores | neg scores | loss], f32 [2B + 1]) at the given device buffers.  Launches issued (or
        captured into a CUDA graph) afterwards read / write them: the host path keeps one buffer pair per
        staging slot so that uploads and downloads of neighbouring batches overlap the kernels."""
        B = self.B
        assert inp.numel() == 5 * B and inp.dtype == i64 and out_buf.numel() == 2 * B + 1 and out_buf.dtype == f32
        self.inp, self.out_buf = inp, out_buf
        self.batch_nids = inp[:3 * B]
        self.src, self.dst, self.neg = inp[:B], inp[B:2 * B], inp[2 * B:3 * B]
        self.pos = inp[:2 * B]
        self.eids = inp[3 * B:4 * B]
        self.ts64 = inp[4 * B:].view(f64)
        self.scores, self.loss = out_buf[:2 * B], out_buf[2 * B:]

    def finder_buffers(self, fresh: bool = False):
        """The finder's outputs (neighbor tables, float32 batch times, involved-node bitmap).  They depend on the
        static graph and the batch only, not on the memory state, so the finder of batch i+1 may run while batch
        i is still in its model kernels - provided it writes its own set (`fresh=True` allocates one)."""
        if not fresh:
            return self.neigh_nids, self.neigh_eids, self.neigh_ts, self.ts32, self.bitmap
        return tuple(torch.zeros_like(t) for t in
                     (self.neigh_nids, self.neigh_eids, self.neigh_ts, self.ts32, self.bitmap))

    def bind_finder(self, bufs):
        self.neigh_nids, self.neigh_eids, self.neigh_ts, self.ts32, self.bitmap = bufs

    def launches_per_step(self) -> int:
        n = sum(KERNELS_PER_STEP.values())
        if self.lazy_restart:
            n += 1 if self.restarter == 'static' else ops.SeqRestarterOp.LAUNCHES + 2
        if self.want_targets:
            n += 1
        return n

    def step(self):
        """One batch, eager launches on the current stream (inputs already in self.inp)."""
        self.launch_finder()
        self.launch_model()


class StreamRunner:
    """Replays an event stream through a TigerEngine.

    `capture()` records the whole batch as one CUDA graph (`run_device`, inputs in engine.inp) and, per staging
    slot, a finder graph and a model graph bound to the slot's own device input / result buffers and finder
    outputs.  `submit_host` / `submit_device` hand a batch to the native pipeline (csrc/pipe.cu): upload + finder
    on a copy-in stream beside the previous batch's model kernels, model graphs in batch order on the current
    stream, download on a copy-out stream - seven CUDA calls per batch, issued from C."""

    def __init__(self, engine: TigerEngine, n_slots: int = 4):
        self.e = engine
        self.graph = None
        self.pipe = None
        B = engine.B
        self.n_slots = n_slots
        self.h_in = [torch.empty(5 * B, dtype=i64).pin_memory() for _ in range(n_slots)]
        self.h_out = [torch.empty(2 * B + 1, dtype=f32).pin_memory() for _ in range(n_slots)]
        self._h_in_np = [t.numpy() for t in self.h_in]
        dev = engine.inp.device
        self.d_in = [torch.zeros(5 * B, dtype=i64, device=dev) for _ in range(n_slots)]
        self.d_out = [torch.zeros(2 * B + 1, dtype=f32, device=dev) for _ in range(n_slots)]
        self.finder_bufs = [engine.finder_buffers(fresh=True) for _ in range(n_slots)]
        self.pq_bufs = [torch.zeros_like(engine.pq) for _ in range(n_slots)]
        self.slot_busy = [False] * n_slots
        self.check_on_wait = False     # True: wait() also polls the device error word (one more host sync per batch)
        self.slot = 0
        self._lib = _lib.load()

    def __del__(self):
        if getattr(self, 'pipe', None):
            self._lib.tiger_pipe_destroy(self.pipe)
            self.pipe = None

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise _lib.TigerLibraryError(f'{what} failed with code {rc}')

    def capture(self, warmup: int = 2):
        """Warm up (eagerly, state restored afterwards is the caller's business) and capture."""
        e = self.e
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                e.step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            e.step()
        torch.cuda.synchronize()
        # Per-slot graphs for the pipeline.  The finder (a function of the static graph and the batch only) is
        # captured separately and replayed on the copy-in stream right behind the slot's upload, i.e. beside the
        # model kernels of the previous batch; it writes the slot's own neighbor tables / bitmap, which that
        # slot's model graph reads.
        if self.pipe:
            self._lib.tiger_pipe_destroy(self.pipe)
        self.pipe = self._lib.tiger_pipe_create(self.n_slots)
        if not self.pipe:
            raise _lib.TigerLibraryError('tiger_pipe_create failed')
        inp0, out0, find0, pq0 = e.inp, e.out_buf, e.finder_buffers(), e.pq
        cap = torch.cuda.Stream()
        with torch.cuda.stream(cap):
            for slot in range(self.n_slots):
                e.bind_io(self.d_in[slot], self.d_out[slot])
                e.bind_finder(self.finder_bufs[slot])
                e.bind_pq(self.pq_bufs[slot])
                for kind, launch in ((0, e.launch_finder), (1, lambda: e.launch_model(with_scorer=False)),
                                     (2, e.launch_scorer)):
                    self._check(self._lib.tiger_pipe_capture_begin(cap.cuda_stream), 'capture_begin')
                    try:
                        launch()
                    finally:
                        rc = self._lib.tiger_pipe_capture_end(self.pipe, cap.cuda_stream, slot, kind)
                    self._check(rc, 'capture_end')
        e.bind_io(inp0, out0)
        e.bind_finder(find0)
        e.bind_pq(pq0)
        torch.cuda.synchronize()

    def join(self):
        """Makes the current stream wait for the copy-out stream's work (link scorer, download) of the latest batch."""
        if self.pipe:
            self._check(self._lib.tiger_pipe_join(self.pipe, _lib.stream_ptr()), 'pipe_join')

    def run_device(self):
        """Inputs already resident in engine.inp."""
        if self.graph is not None:
            self.graph.replay()
        else:
            self.e.step()

    def _next_slot(self) -> int:
        slot = self.slot
        self.slot = (slot + 1) % self.n_slots
        return slot

    def submit_device(self, batch: Tensor) -> int:
        """One step on a batch that is already resident in HBM (int64 [5B], layout of engine.inp): staged into the
        next slot's input buffer on the copy-in stream (beside the previous batch's kernels), then that slot's
        graphs.  Results stay in d_out[slot]; returns the slot."""
        if not self.pipe:
            self.e.inp.copy_(batch, non_blocking=True)
            self.run_device()
            return 0
        slot = self._next_slot()
        self._check(self._lib.tiger_pipe_submit(self.pipe, slot, batch.data_ptr(), self.d_in[slot].data_ptr(),
                                                batch.numel() * 8, None, None, 0, _lib.stream_ptr()), 'pipe_submit')
        return slot

    def fill_host(self, slot: int, src, dst, neg, ts, eids):
        B = self.e.B
        h = self._h_in_np[slot]
        h[:B], h[B:2 * B], h[2 * B:3 * B], h[3 * B:4 * B] = src, dst, neg, eids
        h[4 * B:].view(np.float64)[:] = ts

    def submit_host(self, src, dst, neg, ts, eids) -> int:
        """Host-buffer step: H2D of the batch, the graphs, D2H of scores + loss.  Returns the slot whose
        results become readable after `wait(slot)`."""
        slot = self._next_slot()
        if self.slot_busy[slot]:
            raise RuntimeError(f'staging slot {slot} still holds unread results: call wait(slot) on every slot returned '
                               f'by submit_host before more than {self.n_slots} batches are in flight')
        self.fill_host(slot, src, dst, neg, ts, eids)
        if not self.pipe:
            e = self.e
            e.inp.copy_(self.h_in[slot], non_blocking=True)
            self.run_device()
            self.h_out[slot].copy_(e.out_buf, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        else:
            self._check(self._lib.tiger_pipe_submit(self.pipe, slot, self.h_in[slot].data_ptr(),
                                                    self.d_in[slot].data_ptr(), self.h_in[slot].numel() * 8,
                                                    self.d_out[slot].data_ptr(), self.h_out[slot].data_ptr(),
                                                    self.h_out[slot].numel() * 4, _lib.stream_ptr()), 'pipe_submit')
        self.slot_busy[slot] = True
        return slot

    def wait(self, slot: int):
        if self.pipe and self.slot_busy[slot]:
            self._check(self._lib.tiger_pipe_wait(self.pipe, slot, 1), 'pipe_wait')
        self.slot_busy[slot] = False
        if self.check_on_wait:
            self.e.check_errors()          # attribute an invariant violation to (at most n_slots batches around) this one
        out = self.h_out[slot]
        B = self.e.B
        return out[:B], out[B:2 * B], out[2 * B]

    @property
    def h2d_bytes_per_step(self) -> int:
        return self.h_in[0].numel() * 8

    @property
    def d2h_bytes_per_step(self) -> int:
        return self.h_out[0].numel() * 4
