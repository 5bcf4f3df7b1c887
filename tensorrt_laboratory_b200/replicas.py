"""Multi-GPU = independent model replicas ("replicas only", no collective).

The reference spreads requests over replica processes with an Envoy ``round_robin`` cluster
(reference ``examples/99_LoadBalancer/lb-envoy.j2``) or in-process with ``batch_count % models.size()``
(``trtlab/tensorrt/src/infer_bench.cc:66``).  These helpers are the host-side logic of that topology for
one process per GPU launched by torchrun; the only cross-rank traffic is the benchmark's barrier and a
MAX-reduce of the elapsed time (plumbing, not data path).
"""
from __future__ import annotations

from typing import List, Optional


def shard_requests(n_requests: int, world: int, rank: int) -> List[int]:
    """Round-robin assignment of request ids to replica ``rank`` (Envoy round_robin semantics)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    return list(range(rank, n_requests, world))


def least_outstanding(outstanding: List[int]) -> int:
    """Pick the replica with the fewest in-flight requests (ties -> lowest index)."""
    if not outstanding:
        raise ValueError("no replicas")
    return min(range(len(outstanding)), key=lambda i: (outstanding[i], i))


def _tensor(x: float, dist):
    import torch
    t = torch.tensor([float(x)], dtype=torch.float64)
    return t.cuda() if dist.get_backend() == "nccl" else t


def max_over_ranks(x: float, dist=None) -> float:
    """MAX of one scalar over the replicas (the job's elapsed time is its slowest rank's); identity without a process group."""
    if dist is None or not dist.is_initialized():
        return float(x)
    t = _tensor(x, dist)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_over_ranks(x: float, dist=None) -> List[float]:
    """Every rank's value of one scalar, in rank order (per-replica rates in the N-GPU bench line)."""
    if dist is None or not dist.is_initialized():
        return [float(x)]
    t = _tensor(x, dist)
    out = [t.clone() for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [float(o.item()) for o in out]


def aggregate_throughput(local_elapsed_s: float, steps_per_rank: int, batch: int, dist=None) -> float:
    """Whole-job inferences/s for weak scaling: every rank runs ``steps_per_rank`` steps; the job time is
    the MAX over ranks.  ``dist``: an initialised ``torch.distributed`` module or None (single process)."""
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    return world * steps_per_rank * batch / max_over_ranks(local_elapsed_s, dist)
