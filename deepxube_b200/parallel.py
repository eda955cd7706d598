"""Multi-GPU plumbing for the instance-sharded path (one process per GPU, `torch.distributed`).

Problem instances never interact (each InstanceGraph owns private OPEN / CLOSED,
deepxube/pathfinding/graph_search.py:23-25), so the path shards by instance with no data-path collective:
rank r solves instances r, r + W, r + 2W, ... with its own engine and replicated weights; the only
communication is the final gather of per-instance results and the (sum of work, max of time) reduction
used for throughput reporting.  Backend: NCCL on GPUs, gloo in the CPU tests.
"""
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple


def shard_round_robin(n_items: int, rank: int, world: int) -> List[int]:
    """Indices owned by `rank`: i with i mod world == rank."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    return list(range(rank, n_items, world))


def merge_sharded(per_rank: Sequence[Sequence[Any]], n_items: int) -> List[Any]:
    """Inverse of shard_round_robin: per_rank[r][k] is the result of item r + k*world."""
    world = len(per_rank)
    out: List[Any] = [None] * n_items
    for r, items in enumerate(per_rank):
        idxs = shard_round_robin(n_items, r, world)
        if len(items) != len(idxs):
            raise ValueError(f"rank {r} returned {len(items)} results for {len(idxs)} items")
        for i, v in zip(idxs, items):
            out[i] = v
    return out


def solve_sharded(n_items: int, solve_fn: Callable[[List[int]], List[Any]], dist: Optional[Any] = None) -> Optional[List[Any]]:
    """Every rank calls solve_fn on its shard of item indices; rank 0 gets all results in item order
    (other ranks get None).  Without a process group this is a plain call."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return solve_fn(list(range(n_items)))
    rank, world = dist.get_rank(), dist.get_world_size()
    mine = solve_fn(shard_round_robin(n_items, rank, world))
    gathered: Optional[List[Any]] = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0)
    return merge_sharded(gathered, n_items) if rank == 0 else None


def reduce_throughput(work: float, seconds: float, dist: Optional[Any] = None, device: Optional[Any] = None) -> Tuple[float, float]:
    """(sum of work over ranks, max of seconds over ranks): whole-job throughput = sum / max."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return work, seconds
    import torch
    t = torch.tensor([work, seconds], dtype=torch.float64, device=device)
    s = t.clone()
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(s[0]), float(t[1])
