"""world_size-2 gloo tests (CPU) of the instance-sharded multi-GPU plumbing: sharding, result gathering in
item order, and the (sum of work, max of time) reduction the bench reports."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank: int, world: int, port: int, q) -> None:
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from deepxube_b200 import parallel
    n = 7

    def fake_solve(idxs):            # stands in for an engine solving its shard: returns (index, cost) records
        return [{"idx": i, "cost": 10 * i + rank} for i in idxs]

    res = parallel.solve_sharded(n, fake_solve, dist)
    tot, tmax = parallel.reduce_throughput(100.0 * (rank + 1), 1.0 + rank, dist)
    q.put((rank, res, tot, tmax))
    dist.barrier()
    dist.destroy_process_group()


def test_sharding_helpers():
    sys.path.insert(0, ROOT)
    from deepxube_b200 import parallel
    assert parallel.shard_round_robin(7, 0, 2) == [0, 2, 4, 6] and parallel.shard_round_robin(7, 1, 2) == [1, 3, 5]
    assert parallel.shard_round_robin(2, 3, 8) == []
    assert parallel.merge_sharded([["a", "c"], ["b"]], 3) == ["a", "b", "c"]
    with pytest.raises(ValueError):
        parallel.merge_sharded([["a"], ["b"]], 3)
    with pytest.raises(ValueError):
        parallel.shard_round_robin(4, 2, 2)
    assert parallel.solve_sharded(3, lambda idxs: [i * i for i in idxs]) == [0, 1, 4]
    assert parallel.reduce_throughput(5.0, 2.0) == (5.0, 2.0)


def test_two_rank_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = {}
    for _ in range(2):
        rank, res, tot, tmax = q.get(timeout=120)
        out[rank] = (res, tot, tmax)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res0, tot, tmax = out[0]
    assert out[1][0] is None
    assert [r["idx"] for r in res0] == list(range(7))
    assert [r["cost"] for r in res0] == [10 * i + (i % 2) for i in range(7)]       # item i was solved by rank i mod 2
    assert tot == 300.0 and tmax == 2.0 and out[1][1:] == (300.0, 2.0)
