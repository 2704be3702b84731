"""Data-parallel host logic (margipose_b200/parallel.py) with world_size 2 over gloo on CPU:
batch sharding, start-up broadcast and the single flat-gradient all-reduce."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from margipose_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        assert parallel.world() == (rank, world)
        # start-up broadcast of flat buffers
        flat = torch.full((1000,), float(rank + 1))
        cnt = torch.full((3,), rank + 5, dtype=torch.int64)
        parallel.broadcast_flat([flat, cnt], src=0)
        assert torch.all(flat == 1.0) and torch.all(cnt == 5)
        # one all-reduce over the flat gradient buffer == mean of the per-rank gradients
        g = torch.arange(1000, dtype=torch.float32) * (rank + 1)
        extra = torch.tensor([float(rank + 1), 1.0])
        parallel.allreduce_mean_(g, extra)
        want = torch.arange(1000, dtype=torch.float32) * (sum(range(1, world + 1)) / world)
        torch.testing.assert_close(g, want)
        torch.testing.assert_close(extra, torch.tensor([float(sum(range(1, world + 1))), float(world)]))
        # bucketed path of TrainStep: async sums over the gradient ranges a bucket plan hands out, in plan order;
        # the 1 / world factor is applied later by the SGD kernel (grad_scale)
        n_param = 1000
        plan = parallel.bucket_plan(bwd_marks=[3, 7], stage_ranges=[(100, 400), (400, 900)], n_param=n_param)
        g2 = torch.arange(n_param, dtype=torch.float32) * (rank + 1)
        works = [parallel.allreduce_sum_(g2[a:b], async_op=True) for _lo, _hi, ranges in plan for a, b in ranges]
        for w in works:
            if w is not None:
                w.wait()
        torch.testing.assert_close(g2, torch.arange(n_param, dtype=torch.float32) * sum(range(1, world + 1)))
        lo, hi = parallel.shard_batch(7)
        out.put((rank, lo, hi))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_allreduce_and_shards():
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    shards = sorted(out.get(timeout=10) for _ in range(2))
    assert shards == [(0, 0, 4), (1, 4, 7)]      # contiguous, covering, near-equal


@pytest.mark.parametrize('batch,world', [(32, 1), (256, 8), (10, 4), (3, 8)])
def test_shard_batch_partitions(batch, world):
    spans = [parallel.shard_batch(batch, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == batch
    for (a, b), (c, d) in zip(spans, spans[1:]):
        assert b == c and b >= a
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1


def test_single_process_is_identity():
    g = torch.ones(8)
    assert parallel.allreduce_mean_(g) is g and torch.all(g == 1)
    assert parallel.world() == (0, 1)


def test_bucket_plan_covers_every_gradient_once():
    """Stages finish in reverse order; the last piece (stem backward) carries everything outside the stages."""
    plan = parallel.bucket_plan(bwd_marks=[2, 5, 9], stage_ranges=[(10, 40), (40, 70), (70, 100)], n_param=130)
    assert [(lo, hi) for lo, hi, _r in plan] == [(0, 2), (2, 5), (5, 9), (9, None)]
    assert [r for _lo, _hi, r in plan] == [[(70, 100)], [(40, 70)], [(10, 40)], [(0, 10), (100, 130)]]
    covered = sorted(i for _lo, _hi, ranges in plan for a, b in ranges for i in range(a, b))
    assert covered == list(range(130))
