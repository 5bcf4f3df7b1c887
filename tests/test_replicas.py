"""CPU: the N>1 host-side logic (replicas, no collective) over gloo with world_size 2."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tensorrt_laboratory_b200 import replicas

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_round_robin_sharding_covers_every_request_once():
    for world in (1, 2, 3, 8):
        parts = [replicas.shard_requests(37, world, r) for r in range(world)]
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(37))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        replicas.shard_requests(4, 2, 2)


def test_least_outstanding():
    assert replicas.least_outstanding([3, 1, 1, 5]) == 1
    assert replicas.least_outstanding([0]) == 0
    assert replicas.max_over_ranks(3.5) == 3.5 and replicas.gather_over_ranks(2.0) == [2.0]   # no process group: identity


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # rank r pretends its K steps took (1 + r) seconds: the job time is the max, the work is the sum
    v = replicas.aggregate_throughput(1.0 + rank, steps_per_rank=100, batch=8, dist=dist)
    assert replicas.max_over_ranks(10.0 * (rank + 1), dist) == 10.0 * world
    assert replicas.gather_over_ranks(float(rank), dist) == [float(r) for r in range(world)]
    mine = replicas.shard_requests(10, world, rank)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    q.put((rank, v, gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_weak_scaling_aggregate_over_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, v, gathered in res:
        assert abs(v - 2 * 100 * 8 / 2.0) < 1e-9  # 2 ranks x 100 steps x 8 / max(1 s, 2 s)
        assert sorted(i for part in gathered for i in part) == list(range(10))


def test_single_process_aggregate():
    assert replicas.aggregate_throughput(0.5, 10, 8) == 160.0
