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
