"""world_size-2 gloo test of the N>1 host logic (scene sharding, max-over-ranks timing, throughput aggregation)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from coalign_b200 import dist_utils as D
    r, w = D.init("gloo")
    assert (r, w) == (rank, world)
    mine = D.shard_scenes(list(range(10)), r, w)
    D.barrier()
    ms = 10.0 + 5.0 * rank                       # rank 1 is slower
    worst = D.max_over_ranks(ms)
    total = D.sum_over_ranks(len(mine))
    out.put((rank, mine, worst, total))
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, s0, w0, t0), (r1, s1, w1, t1) = res
    assert s0 == [0, 2, 4, 6, 8] and s1 == [1, 3, 5, 7, 9]          # disjoint, complete
    assert w0 == w1 == 15.0                                         # max over ranks
    assert t0 == t1 == 10.0


def _bucket_worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from coalign_b200 import dist_utils as D
    from coalign_b200.trainer import allreduce_buckets
    D.init("gloo")
    g = torch.Generator().manual_seed(rank)
    flat = torch.randn(1000, generator=g)
    mine = flat.clone()
    buckets = [(0, 300), (300, 304), (304, 900), (900, 1000)]           # completion-ordered gradient slices
    works = []
    for i in range(len(buckets)):                                       # one call per finished backward segment
        works += allreduce_buckets(flat, buckets, upto=i + 1, start=i)
    for w in works:
        w.wait()
    out.put((rank, mine, flat))
    dist.destroy_process_group()


def test_two_rank_bucketed_gradient_allreduce():
    """N>1 path of the trainer (gloo stands in for NCCL): all-reducing the completion-ordered buckets one by one gives the
    sum of the ranks' flat gradient buffers, i.e. what DistributedDataParallel leaves in .grad before its division."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted((q.get(timeout=120) for _ in ps), key=lambda t: t[0])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    total = res[0][1] + res[1][1]
    assert torch.allclose(res[0][2], total) and torch.allclose(res[1][2], total)
