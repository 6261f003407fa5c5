"""N > 1 host logic on CPU (gloo, world_size 2): the reference's DDP partition (ChunkSampler,
tiger/data/data_loader.py:17-40) as mirrored by the drop-in package and by bench.py, and the
max-over-ranks timing reduction of the benchmark.  No GPU, no data-path collective: ranks only
exchange their chunk bounds and a timing scalar, exactly like the real multi-GPU run."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_events, bs, seed, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'www2023tiger_b200'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import bench
        from tiger.data.data_loader import ChunkSampler
        sampler = ChunkSampler(n_events, rank, world, bs, seed)
        lo, hi = sampler.bounds()
        assert (lo, hi) == bench.chunk_bounds(n_events, rank, world, bs, seed)
        assert list(iter(sampler))[:3] == [lo, lo + 1, lo + 2] and len(sampler) == hi - lo
        mine = torch.tensor([lo, hi], dtype=torch.int64)
        gathered = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(gathered, mine)
        bounds = [tuple(int(x) for x in g) for g in gathered]
        length = n_events // (world * bs) * bs
        for r, (a, b) in enumerate(bounds):
            assert b - a == length and (b - a) % bs == 0            # whole batches, equal work per rank
            if r > 0:
                assert a == bounds[r - 1][1]                        # contiguous time chunks, no overlap, no gap
        assert 0 <= bounds[0][0] <= n_events % (world * bs) and bounds[-1][1] <= n_events
        # the benchmark's clock: every rank reports the slowest rank's time
        slowest = bench.max_over_ranks(10.0 + 5.0 * rank, world)
        assert slowest == 10.0 + 5.0 * (world - 1)
        # per-rank batch windows of bench.py follow the same partition
        class A:
            seed_ = seed
        args = type('Args', (), {'seed': seed, 'skip_batches': 1000})()
        w_lo, w_n = bench.batch_window(args, n_events, rank, world)
        assert (w_lo, w_n) == (lo, length // bs)
        open(os.path.join(out_dir, f'ok{rank}'), 'w').write(f'{lo} {hi}')
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n_events,bs', [(157474, 200), (4001, 200)])
def test_chunk_partition_and_timing_reduce_world2(tmp_path, n_events, bs):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, n_events, bs, 0, str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == ['ok0', 'ok1']


# ------------------------------------------------------------------------------------------
# DDP gradient exchange of the native training step (www2023tiger_b200/train.py): the flat gradient buffer is
# all-reduced in three slices, each as soon as the backward pass has completed it; the two gates that decide which
# tensors Adam steps ride at the tail of the same buffer.  Host logic only (gloo, CPU tensors, no kernels).
# ------------------------------------------------------------------------------------------
def _grad_worker(rank, world, port, restarter, out_dir):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from www2023tiger_b200.init import build_model
        from www2023tiger_b200.train import FlatParams, _param_group
        torch.manual_seed(0)
        model = build_model(None, torch.zeros(50, 6), None, 40, 49, torch.device('cpu'), dim=None, n_neighbors=5,
                            restarter_type=restarter, hist_len=4)
        before = {k: v.clone() for k, v in model.state_dict().items()}
        fp = FlatParams(model, _param_group)
        # re-pointing the parameters into the flat buffer changes no value and no state_dict key
        after = model.state_dict()
        assert list(before) == list(after) and all(torch.equal(before[k], after[k]) for k in before)
        assert all(p.data_ptr() == fp.p[n].data_ptr() for n, p in zip(fp.names, fp.params))
        ranges = fp.comm_ranges()
        assert ranges is not None
        spans = sorted(ranges.values())
        assert spans[0][0] == 0 and spans[-1][1] == fp.grad_all.numel() == fp.numel + FlatParams.N_GATES
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))          # a partition of [gradients | gates]
        for name, (lo, hi) in ranges.items():
            inside = [n for n, o in zip(fp.names, fp.offsets) if lo <= o < hi]
            want = {'restarter': ('restarter_fn.',), 'attention': ('temporal_embedding_fn.', 'hit_embedding.', 'score_fn.'),
                    'gru': ('time_encoder.', 'right_mem_updater.')}[name]
            assert inside and all(n.startswith(want) for n in inside), (name, inside)
        assert [fp.groups[fp.names.index(n)] for n in ('time_encoder.phase', 'right_mem_updater.cell.bias_hh')] == [0, 1]
        assert all(g == 2 for n, g in zip(fp.names, fp.groups) if n.startswith('restarter_fn.'))
        # rank-dependent gradients and gates; sliced async all-reduces in backward order == one all-reduce of everything
        g = torch.Generator().manual_seed(rank)
        fp.grad_all.copy_(torch.randn(fp.grad_all.numel(), generator=g))
        fp.gates[0] = float(rank == 1)              # only rank 1 used the GRU cell in this step
        whole = fp.grad_all.clone()
        dist.all_reduce(whole)
        works = [dist.all_reduce(fp.grad_all[ranges[n][0]:ranges[n][1]], async_op=True)
                 for n in ('restarter', 'attention', 'gru')]
        for w in works:
            w.wait()
        assert torch.equal(fp.grad_all, whole)
        assert float(fp.gates[0]) == 1.0            # used on any rank -> stepped on every rank
        open(os.path.join(out_dir, f'ok{rank}'), 'w').write('ok')
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('restarter', ['seq', 'static'])
def test_flat_gradient_slices_and_gates_all_reduce_world2(tmp_path, restarter):
    world, port = 2, _free_port()
    mp.spawn(_grad_worker, args=(world, port, restarter, str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == ['ok0', 'ok1']
