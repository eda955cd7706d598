"""lightsout domain (host view): LOState(tiles) / LOGoal(tiles) / LOAction(action, dim), registered as "lightsout" with
parser `lightsout.<dim>` (deepxube/domains/lightsout.py:16-58, 60-96, 192-198).  dim*dim fixed actions; action a flips
cell a and its on-board neighbours (move matrix [a, a+dim, a-dim, a+1, a-1] with off-board entries replaced by a; the
reference flips through one fancy-index assignment, so a cell listed twice is flipped once, :178-188).  Goal: all off.
NN input: the tiles with one-hot depth 1, i.e. the 0 / 1 values themselves (:107-108, :113-114)."""
import random
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from deepxube_b200.base.domain import Action, Goal, GoalStartRevWalkable, PackedDomain, State
from deepxube_b200.base.factory import Parser
from deepxube_b200.domains.tables import lightsout_move_table
from deepxube_b200.factories.domain_factory import domain_factory


class LOState(State):
    __slots__ = ["tiles", "hash"]

    def __init__(self, tiles: np.ndarray):
        self.tiles = tiles
        self.hash: Optional[int] = None

    def __hash__(self) -> int:
        if self.hash is None:
            self.hash = hash(self.tiles.tobytes())
        return self.hash

    def __eq__(self, other: object) -> bool:
        if isinstance(other, LOState):
            return np.array_equal(self.tiles, other.tiles)
        return NotImplemented


class LOGoal(Goal):
    def __init__(self, tiles: np.ndarray):
        self.tiles = tiles


class LOAction(Action):
    def __init__(self, action: int, dim: int):
        self.action = action
        self.dim = dim

    def __hash__(self) -> int:
        return hash((self.action, self.dim))

    def __eq__(self, other: object) -> bool:
        if isinstance(other, LOAction):
            return self.action == other.action
        return NotImplemented

    def __repr__(self) -> str:
        return f"{self.action % self.dim},{self.action // self.dim}"


@domain_factory.register_class("lightsout")
class LightsOut(PackedDomain[LOState, LOAction, LOGoal], GoalStartRevWalkable[LOState, LOAction, LOGoal]):
    dx_name = "lightsout"

    def __init__(self, dim: int = 7):
        super().__init__()
        if not 2 <= dim <= 15:
            raise ValueError("lightsout: the B200 engine indexes cells with uint8, dim must be in 2..15")
        self.dim, self.dx_dim = dim, dim
        self.dtype = np.uint8
        self.num_tiles = dim * dim
        self.move_matrix = lightsout_move_table(dim)
        self.goal_np = np.zeros(self.num_tiles, dtype=np.uint8)
        self.num_actions = self.num_tiles
        self.actions_fixed: List[LOAction] = [LOAction(a, dim) for a in range(self.num_tiles)]

    def states_to_rows(self, states: List[LOState]) -> np.ndarray:
        return np.stack([s.tiles for s in states]).astype(np.uint8) if states else np.zeros((0, self.num_tiles), np.uint8)

    def rows_to_states(self, rows: np.ndarray) -> List[LOState]:
        return [LOState(r) for r in np.asarray(rows, np.uint8)]

    def goals_to_rows(self, goals: List[LOGoal]) -> np.ndarray:
        return np.stack([g.tiles for g in goals]).astype(np.uint8) if goals else np.zeros((0, self.num_tiles), np.uint8)

    def action_index(self, action: LOAction) -> int:
        return action.action

    def get_actions_fixed(self) -> List[LOAction]:
        return self.actions_fixed.copy()

    def actions_to_indices(self, actions: List[LOAction]) -> List[int]:
        return [a.action for a in actions]

    def sample_goalstate_goal_pairs(self, num: int) -> Tuple[List[LOState], List[LOGoal]]:
        return [LOState(self.goal_np.copy()) for _ in range(num)], [LOGoal(self.goal_np.copy()) for _ in range(num)]

    def _walk_rows(self, rows: np.ndarray, num_steps: np.ndarray) -> Tuple[np.ndarray, List[List[int]]]:
        """problem-instance generation (host side of the path): random presses from the given rows"""
        rows = rows.copy()
        n = rows.shape[0]
        taken: List[List[int]] = [[] for _ in range(n)]
        for t in range(int(num_steps.max()) if n else 0):
            for i in np.nonzero(num_steps > t)[0].tolist():
                a = random.randrange(self.num_actions)
                rows[i, np.unique(self.move_matrix[a])] ^= 1
                taken[i].append(a)
        return rows, taken

    def random_walk(self, states: List[LOState], num_steps_l: List[int]):
        rows, taken = self._walk_rows(self.states_to_rows(states), np.asarray(num_steps_l, dtype=int))
        return self.rows_to_states(rows), [[LOAction(a, self.dim) for a in t] for t in taken], [float(len(t)) for t in taken]

    def random_walk_rev_no_path_cost(self, states: List[LOState], num_steps_l: List[int]) -> List[LOState]:
        return self.random_walk(states, num_steps_l)[0]

    def get_input_info_flat_sg(self) -> Tuple[List[int], List[int]]:
        return [self.num_tiles], [1]

    def to_np_flat_sg(self, states: List[LOState], goals: List[LOGoal]) -> List[np.ndarray]:
        return [self.states_to_rows(states)]

    def string_to_action(self, act_str: str) -> Optional[LOAction]:
        x_str, y_str = act_str.split(",")
        return LOAction(int(y_str) * self.dim + int(x_str), self.dim)

    def __repr__(self) -> str:
        return f"LightsOut(dim={self.dim})"


@domain_factory.register_parser("lightsout")
class LightsOutParser(Parser):
    def parse(self, args_str: str) -> Dict[str, Any]:
        return {"dim": int(args_str)}

    def help(self) -> str:
        return "An integer for the dimension. E.g. '7'"
