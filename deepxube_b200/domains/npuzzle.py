"""npuzzle domain (host view): NPState(tiles) / NPGoal(tiles) / NPAction(action), registered as "npuzzle" with
parser `npuzzle.<dim>` (deepxube/domains/npuzzle.py:18-73, 324-333).  Four fixed actions U D L R moving the
blank to row+1 / row-1 / col+1 / col-1; a move off the board is a no-op that still costs 1.0 (:264-321).
Only dim <= 15 (uint8 tiles) is supported by the kernels."""
import random
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from deepxube_b200.base.domain import Action, Goal, GoalStartRevWalkable, PackedDomain, State
from deepxube_b200.base.factory import Parser
from deepxube_b200.domains.tables import npuzzle_swap_table
from deepxube_b200.factories.domain_factory import domain_factory


class NPState(State):
    __slots__ = ["tiles", "hash"]

    def __init__(self, tiles: np.ndarray):
        self.tiles = tiles
        self.hash: Optional[int] = None

    def __hash__(self) -> int:
        if self.hash is None:
            self.hash = hash(self.tiles.tobytes())
        return self.hash

    def __eq__(self, other: object) -> bool:
        if isinstance(other, NPState):
            return np.array_equal(self.tiles, other.tiles)
        return NotImplemented


class NPGoal(Goal):
    def __init__(self, tiles: np.ndarray):
        self.tiles = tiles


class NPAction(Action):
    def __init__(self, action: int):
        self.action = action

    def __hash__(self) -> int:
        return self.action

    def __eq__(self, other: object) -> bool:
        if isinstance(other, NPAction):
            return self.action == other.action
        return NotImplemented

    def __repr__(self) -> str:
        return "UDLR"[self.action]


@domain_factory.register_class("npuzzle")
class NPuzzle(PackedDomain[NPState, NPAction, NPGoal], GoalStartRevWalkable[NPState, NPAction, NPGoal]):
    moves: List[str] = ["U", "D", "L", "R"]
    dx_name = "npuzzle"

    def __init__(self, dim: int = 4):
        super().__init__()
        if not 2 <= dim <= 15:
            raise ValueError("npuzzle: the B200 engine packs tiles in uint8 rows, dim must be in 2..15")
        self.dim, self.dx_dim = dim, dim
        self.dtype = np.uint8
        self.num_tiles = dim * dim
        self.goal_tiles = np.concatenate((np.arange(1, dim * dim), [0])).astype(np.uint8)
        self.swap_zero_idxs = npuzzle_swap_table(dim)
        self.num_actions = 4
        self.actions: List[NPAction] = [NPAction(a) for a in range(4)]

    def states_to_rows(self, states: List[NPState]) -> np.ndarray:
        return np.stack([s.tiles for s in states]).astype(np.uint8) if states else np.zeros((0, self.num_tiles), np.uint8)

    def rows_to_states(self, rows: np.ndarray) -> List[NPState]:
        return [NPState(r) for r in np.asarray(rows, np.uint8)]

    def goals_to_rows(self, goals: List[NPGoal]) -> np.ndarray:
        return np.stack([g.tiles for g in goals]).astype(np.uint8) if goals else np.zeros((0, self.num_tiles), np.uint8)

    def action_index(self, action: NPAction) -> int:
        return action.action

    def get_actions_fixed(self) -> List[NPAction]:
        return self.actions.copy()

    def sample_goalstate_goal_pairs(self, num: int) -> Tuple[List[NPState], List[NPGoal]]:
        return [NPState(self.goal_tiles.copy()) for _ in range(num)], [NPGoal(self.goal_tiles.copy()) for _ in range(num)]

    def _walk_rows(self, rows: np.ndarray, num_steps: np.ndarray) -> Tuple[np.ndarray, List[List[int]]]:
        rows = rows.copy()
        n = rows.shape[0]
        z = np.argmax(rows == 0, axis=1)
        taken: List[List[int]] = [[] for _ in range(n)]
        for t in range(int(num_steps.max()) if n else 0):
            live = np.nonzero(num_steps > t)[0]
            acts = np.array([random.randrange(4) for _ in live], dtype=np.int64)
            sw = self.swap_zero_idxs[z[live], acts]
            rows[live, z[live]] = rows[live, sw]
            rows[live, sw] = 0
            z[live] = sw
            for i, a in zip(live.tolist(), acts.tolist()):
                taken[i].append(a)
        return rows, taken

    def random_walk(self, states: List[NPState], num_steps_l: List[int]):
        rows, taken = self._walk_rows(self.states_to_rows(states), np.asarray(num_steps_l, dtype=int))
        return self.rows_to_states(rows), [[NPAction(a) for a in t] for t in taken], [float(len(t)) for t in taken]

    def random_walk_rev_no_path_cost(self, states: List[NPState], num_steps_l: List[int]) -> List[NPState]:
        return self.random_walk(states, num_steps_l)[0]

    # NN input (deepxube/domains/npuzzle.py:160-164): the tiles, one-hot depth d*d; the goal is not encoded
    def get_input_info_flat_sg(self) -> Tuple[List[int], List[int]]:
        return [self.num_tiles], [self.num_tiles]

    def to_np_flat_sg(self, states: List[NPState], goals: List[NPGoal]) -> List[np.ndarray]:
        return [self.states_to_rows(states)]

    def string_to_action(self, act_str: str) -> Optional[NPAction]:
        return {"w": NPAction(0), "s": NPAction(1), "a": NPAction(2), "d": NPAction(3)}.get(act_str)

    def __repr__(self) -> str:
        return f"NPuzzle(dim={self.dim})"


@domain_factory.register_parser("npuzzle")
class NPuzzleParser(Parser):
    def parse(self, args_str: str) -> Dict[str, Any]:
        return {"dim": int(args_str)}

    def help(self) -> str:
        return "An integer for the dimension. E.g. 'npuzzle.4'"
