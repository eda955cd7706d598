"""Replay buffers of the RL updaters, array-backed.

Reference: deepxube/updaters/utils/replay_buffer_utils.py:28-84 -- a `deque(maxlen=max_size)` of per-element tuples,
`add` = `deque.extend` (the oldest elements fall out), `sample(num)` = `np.random.randint(0, len(deque), size=num)` indices
into the deque (index 0 = oldest element).  Here the elements live in preallocated numpy arrays used as a ring (states and
goals as packed uint8 rows, the layout the engine and `dx_vi_targets` take), so adding a batch is one slice copy and a sample
is one fancy-index gather -- no per-element Python objects.  `sample` draws from `np.random` with the reference's exact call, so
with the same seed and the same add sequence it returns the same elements as the reference's buffer
(tests/test_host_api.py::test_replay_buffers_match_the_reference_deque).
"""
from typing import Dict, Tuple

import numpy as np


class _Ring:
    """`maxlen` deque of rows over named numpy columns."""

    def __init__(self, max_size: int, columns: Dict[str, Tuple[Tuple[int, ...], np.dtype]]):
        self._cap = int(max_size)
        self._cols = {k: np.zeros((max(self._cap, 1),) + tuple(shape), dtype) for k, (shape, dtype) in columns.items()}
        self._head = 0          # ring position of deque index 0 (the oldest element)
        self._size = 0

    def size(self) -> int:
        return self._size

    def max_size(self) -> int:
        return self._cap

    def _extend(self, **cols: np.ndarray) -> None:
        n = len(next(iter(cols.values())))
        for k, v in cols.items():
            assert len(v) == n, f"column {k} has {len(v)} rows, expected {n}"
        if self._cap == 0 or n == 0:
            return
        if n >= self._cap:                              # only the newest `cap` elements survive (deque.extend semantics)
            for k, v in cols.items():
                self._cols[k][:] = np.asarray(v)[n - self._cap:]
            self._head, self._size = 0, self._cap
            return
        tail = (self._head + self._size) % self._cap
        first = min(n, self._cap - tail)
        for k, v in cols.items():
            v = np.asarray(v)
            self._cols[k][tail:tail + first] = v[:first]
            if first < n:
                self._cols[k][:n - first] = v[first:]
        over = max(0, self._size + n - self._cap)
        self._head = (self._head + over) % self._cap
        self._size = min(self._cap, self._size + n)

    def _take(self, idxs: np.ndarray) -> Dict[str, np.ndarray]:
        pos = (self._head + np.asarray(idxs, np.int64)) % max(self._cap, 1)
        return {k: c[pos] for k, c in self._cols.items()}

    def _sample_idxs(self, num: int) -> np.ndarray:
        assert self._size > 0, f"Replay buffer size should not be {self._size}"
        return np.random.randint(0, self._size, size=num)          # replay_buffer_utils.py:39


class ReplayBufferV(_Ring):
    """ReplayBufferV (replay_buffer_utils.py:56-67): (state, goal, is_solved) per popped node (contexts are None on this path)."""

    def __init__(self, max_size: int, state_bytes: int, goal_bytes: int):
        super().__init__(max_size, {"state": ((state_bytes,), np.uint8), "goal": ((goal_bytes,), np.uint8), "solved": ((), np.bool_)})

    def add(self, state_rows: np.ndarray, goal_rows: np.ndarray, is_solved: np.ndarray) -> None:
        self._extend(state=state_rows, goal=goal_rows, solved=np.asarray(is_solved, bool))

    def sample(self, num: int) -> Tuple[Tuple[np.ndarray, np.ndarray], np.ndarray]:
        """-> ((state rows, goal rows), is_solved)"""
        d = self._take(self._sample_idxs(num))
        return (d["state"], d["goal"]), d["solved"]


class ReplayBufferQ(_Ring):
    """ReplayBufferQ (replay_buffer_utils.py:70-84): (state, goal, action, is_solved, transition cost, next state) per popped edge."""

    def __init__(self, max_size: int, state_bytes: int, goal_bytes: int):
        super().__init__(max_size, {"state": ((state_bytes,), np.uint8), "goal": ((goal_bytes,), np.uint8), "action": ((), np.int32),
                                    "solved": ((), np.bool_), "tc": ((), np.float64), "next": ((state_bytes,), np.uint8)})

    def add(self, state_rows: np.ndarray, goal_rows: np.ndarray, actions: np.ndarray, is_solved: np.ndarray, tcs: np.ndarray,
            next_rows: np.ndarray) -> None:
        self._extend(state=state_rows, goal=goal_rows, action=np.asarray(actions, np.int32), solved=np.asarray(is_solved, bool),
                     tc=np.asarray(tcs, np.float64), next=next_rows)

    def sample(self, num: int) -> Tuple[Tuple[np.ndarray, np.ndarray, np.ndarray], Tuple[np.ndarray, np.ndarray, np.ndarray]]:
        """-> ((state rows, goal rows, action ids), (is_solved, tcs, next-state rows))"""
        d = self._take(self._sample_idxs(num))
        return (d["state"], d["goal"], d["action"]), (d["solved"], d["tc"], d["next"])
