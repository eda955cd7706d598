"""Small helpers on top of the PathFind API (used by bench.py and the multi-GPU driver)."""
from typing import List, Tuple

import numpy as np


def solve_batch_e2e(eng, states: np.ndarray, goals: np.ndarray, batch_size: int, weight: float, max_itrs: int) -> Tuple[int, int, int]:
    """One end-to-end batch through the C ABI with HOST buffers: reset, upload the start states, run up to
    max_itrs iterations, read every instance's status back.  -> (nodes generated, H2D bytes, D2H bytes)"""
    eng.reset()
    eng.add_instances(states, goals, batch_size, weight)
    eng.run(max_itrs)
    sts = eng.status()
    nodes = int(sum(s.num_nodes_generated for s in sts))
    n = states.shape[0]
    h2d = int(states.nbytes + goals.nbytes + n * (4 + 8))            # rows + per-instance batch size / weight
    d2h = int(len(sts) * 56)                                         # dx_inst_status per instance
    return nodes, h2d, d2h


def shard_round_robin(n_items: int, rank: int, world: int) -> List[int]:
    """Instance i -> rank i mod world (instances never interact, no data-path collective)."""
    return list(range(rank, n_items, world))
