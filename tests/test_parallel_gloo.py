"""CPU tests (gloo, world_size 2) of the N>1 host logic: batch sharding, the single flat-gradient all-reduce
(sum of per-shard gradients pre-scaled by 1/world == full-batch mean gradient), max-over-ranks timing."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from unipose_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert parallel.world() == (rank, world)
        torch.manual_seed(0)
        x = torch.randn(10, 6)
        y = torch.randn(10, 3)
        w = torch.randn(6, 3, requires_grad=True)
        start, count = parallel.shard_range(10, rank, world)
        # local step on the shard: MSE 'mean' over the local shard, loss gradient pre-scaled by 1/world
        loss = ((x[start:start + count] @ w - y[start:start + count]) ** 2).mean() / world
        loss.backward()
        flat = w.grad.detach().reshape(-1).clone()
        parallel.allreduce_sum_(flat)
        # full-batch reference (equal shard sizes -> mean of shard means == global mean)
        w2 = w.detach().clone().requires_grad_(True)
        ((x @ w2 - y) ** 2).mean().backward()
        ok_grad = torch.allclose(flat, w2.grad.reshape(-1), atol=1e-6)
        # the bucketed form (what TrainStep overlaps with the backward) gives the same sum
        buckets = parallel.make_buckets([(0, 5, 3), (5, 7, 9), (12, 6, 4)], bucket_bytes=40, n_ops=11)
        flat_b = w.grad.detach().reshape(-1).clone()
        parallel.allreduce_buckets_(flat_b, buckets)
        ok_grad = ok_grad and torch.equal(flat_b, flat) and buckets[-1][1] == 18
        tmax = parallel.max_over_ranks([1.0 + rank, 5.0 - rank], "cpu")
        out.put((rank, ok_grad, tmax, (start, count)))
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions():
    for total in (1, 7, 32, 33, 256):
        for world in (1, 2, 3, 8):
            cover = []
            for r in range(world):
                s, c = parallel.shard_range(total, r, world)
                cover.extend(range(s, s + c))
            assert cover == list(range(total))


def test_two_rank_gradient_allreduce_and_timing():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_grad, tmax, shard in res:
        assert ok_grad, rank
        assert tmax == [2.0, 5.0]
    assert sorted(r[3] for r in res) == [(0, 5), (5, 5)]


def test_make_buckets_covers_buffer_in_backward_order():
    items = [(0, 1000, 5), (1000, 3000, 9), (4000, 10, 7), (4010, 5000, 20), (9010, 20, 18)]
    b = parallel.make_buckets(items, bucket_bytes=12000, n_ops=25)
    assert b[0][0] == 0 and b[-1][1] == 9030
    assert all(x[1] == y[0] for x, y in zip(b, b[1:]))                # contiguous, no gaps
    assert all(x[2] <= y[2] for x, y in zip(b, b[1:])) and b[-1][2] == 25
    # a bucket is never released before the last gradient inside it is final
    for lo, hi, ready in b[:-1]:
        assert ready >= max(r for off, n, r in items if lo <= off < hi)
    assert parallel.make_buckets([(0, 10, 1)], bucket_bytes=1 << 30, n_ops=4) == [(0, 10, 4)]
