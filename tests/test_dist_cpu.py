"""CPU: the N > 1 path (batch sharding, no data-path collective) on a world_size-2 gloo group."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spike2former_b200 import dist as s2f_dist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert s2f_dist.env_world() == (rank, world, rank)
        b, e = s2f_dist.shard_range(total, world, rank)
        # each rank "segments" its own images: the label map of image i is filled with i
        labels = torch.stack([torch.full((4, 6), i, dtype=torch.uint8) for i in range(b, e)]) if e > b else \
            torch.zeros((0, 4, 6), dtype=torch.uint8)
        full = s2f_dist.gather_label_maps(labels, total)
        slowest = s2f_dist.max_over_ranks(10.0 + rank)
        dist.barrier()
        q.put((rank, (b, e), full[:, 0, 0].tolist(), slowest))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_timing_protocol():
    world, total = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ranges = [o[1] for o in out]
    assert ranges == [(0, 4), (4, 7)]                               # disjoint, complete, balanced
    for o in out:
        assert o[2] == list(range(total))                           # every rank sees every image exactly once
        assert o[3] == 11.0                                         # step time = slowest rank


def test_shard_range_properties():
    for total in (0, 1, 16, 17, 129):
        for world in (1, 2, 4, 8):
            spans = [s2f_dist.shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _bucket_worker(rank, world, port, q):
    import torch.nn as nn

    from spike2former_b200 import train

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        model = nn.Sequential(nn.Linear(64, 300), nn.ReLU(), nn.Linear(300, 300), nn.ReLU(), nn.Linear(300, 5))
        unused = nn.Parameter(torch.zeros(7))                       # a parameter that never receives a gradient
        params = list(model.parameters()) + [unused]
        buckets = train.GradBuckets(params, bucket_mb=0.2)           # several buckets
        g = torch.Generator().manual_seed(1)
        x, y = torch.randn(8, 64, generator=g), torch.randn(8, 5, generator=g)
        b, e = s2f_dist.shard_range(8, world, rank)
        for _ in range(2):                                           # hooks re-arm every step
            model.zero_grad()
            buckets.start()
            ((model(x[b:e]) - y[b:e]) ** 2).mean().backward()
            buckets.finish()
        got = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
        model.zero_grad()
        ((model(x) - y) ** 2).mean().backward()                      # the full batch on one process
        want = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
        q.put((rank, len(buckets.buckets), float((got - want).abs().max()), buckets.bytes))
    finally:
        dist.destroy_process_group()


def test_gradient_buckets_average_like_one_big_batch():
    """train.GradBuckets (the training step's NCCL all-reduce, here over gloo): two ranks with half the batch each end up
    with the gradient of the whole batch; buckets are launched from autograd hooks."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bucket_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=60) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, nb, err, nbytes in out:
        assert nb >= 2 and err < 1e-6 and nbytes == 4 * (64 * 300 + 300 + 300 * 300 + 300 + 300 * 5 + 5 + 7)
