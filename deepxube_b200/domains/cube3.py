"""cube3 domain (host view).  Same public surface as deepxube/domains/cube3.py:14-59, 486-596 minus the
matplotlib visualisation: Cube3State(colors) / Cube3Goal(colors) / Cube3Action(action), registered as
"cube3", 12 fixed actions, reversible (inverse of action a is a^1), goal-start reverse-walkable.
next_state / expand / is_solved run on the GPU (PackedDomain)."""
import random
from typing import List, Optional, Tuple

import numpy as np

from deepxube_b200.base.domain import Action, ActsRev, Goal, GoalStartRevWalkable, PackedDomain, State
from deepxube_b200.domains.tables import CUBE3_ACTION_NAMES, cube3_perm_table
from deepxube_b200.factories.domain_factory import domain_factory

CUBE3_PERM = cube3_perm_table()


class Cube3State(State):
    __slots__ = ["colors", "hash"]

    def __init__(self, colors: np.ndarray) -> None:
        self.colors = colors
        self.hash: Optional[int] = None

    def __hash__(self) -> int:
        if self.hash is None:
            self.hash = hash(self.colors.tobytes())
        return self.hash

    def __eq__(self, other: object) -> bool:
        if isinstance(other, Cube3State):
            return np.array_equal(self.colors, other.colors)
        return NotImplemented


class Cube3Goal(Goal):
    def __init__(self, colors: np.ndarray):
        self.colors = colors


class Cube3Action(Action):
    def __init__(self, action: int) -> None:
        self.action = action

    def __hash__(self) -> int:
        return self.action

    def __eq__(self, other: object) -> bool:
        if isinstance(other, Cube3Action):
            return self.action == other.action
        return NotImplemented

    def __repr__(self) -> str:
        return CUBE3_ACTION_NAMES[self.action]


@domain_factory.register_class("cube3")
class Cube3(PackedDomain[Cube3State, Cube3Action, Cube3Goal], GoalStartRevWalkable[Cube3State, Cube3Action, Cube3Goal],
            ActsRev[Cube3State, Cube3Action, Cube3Goal]):
    atomic_actions: List[str] = CUBE3_ACTION_NAMES
    dx_name, dx_dim = "cube3", 0

    def __init__(self) -> None:
        super().__init__()
        self.cube_len, self.num_colors, self.num_stickers = 3, 6, 54
        self.num_actions = 12
        self.goal_colors = (np.arange(54) // 9).astype(np.uint8)
        self.actions: List[Cube3Action] = [Cube3Action(a) for a in range(12)]

    # row conversion
    def states_to_rows(self, states: List[Cube3State]) -> np.ndarray:
        return np.stack([s.colors for s in states]).astype(np.uint8) if states else np.zeros((0, 54), np.uint8)

    def rows_to_states(self, rows: np.ndarray) -> List[Cube3State]:
        return [Cube3State(r) for r in np.asarray(rows, np.uint8)]

    def goals_to_rows(self, goals: List[Cube3Goal]) -> np.ndarray:
        return np.stack([g.colors for g in goals]).astype(np.uint8) if goals else np.zeros((0, 54), np.uint8)

    def action_index(self, action: Cube3Action) -> int:
        return action.action

    def get_actions_fixed(self) -> List[Cube3Action]:
        return self.actions.copy()

    # goals / instance generation (host side)
    def get_goal_states(self, num_states: int) -> List[Cube3State]:
        return [Cube3State(self.goal_colors.copy()) for _ in range(num_states)]

    def sample_goalstate_goal_pairs(self, num: int) -> Tuple[List[Cube3State], List[Cube3Goal]]:
        return [Cube3State(self.goal_colors.copy()) for _ in range(num)], [Cube3Goal(self.goal_colors.copy()) for _ in range(num)]

    def _walk_rows(self, rows: np.ndarray, num_steps: np.ndarray) -> Tuple[np.ndarray, List[List[int]]]:
        rows = rows.copy()
        taken: List[List[int]] = [[] for _ in range(rows.shape[0])]
        for t in range(int(num_steps.max()) if len(num_steps) else 0):
            live = np.nonzero(num_steps > t)[0]
            acts = np.array([random.randrange(12) for _ in live], dtype=np.int64)
            rows[live] = np.take_along_axis(rows[live], CUBE3_PERM[acts], axis=1)
            for i, a in zip(live.tolist(), acts.tolist()):
                taken[i].append(a)
        return rows, taken

    def random_walk(self, states: List[Cube3State], num_steps_l: List[int]):
        rows, taken = self._walk_rows(self.states_to_rows(states), np.asarray(num_steps_l, dtype=int))
        return self.rows_to_states(rows), [[Cube3Action(a) for a in t] for t in taken], [float(len(t)) for t in taken]

    def random_walk_rev_no_path_cost(self, states: List[Cube3State], num_steps_l: List[int]) -> List[Cube3State]:
        return self.random_walk(states, num_steps_l)[0]

    def sample_rev_state(self, states: List[Cube3State]):
        actions = self.sample_state_action(states)
        prev = self.next_state(states, actions)[0]
        return prev, [Cube3Action(a.action ^ 1) for a in actions], [1.0] * len(states)

    # NN input (deepxube/domains/cube3.py:528-535): the 54 colours, one-hot depth 6; the goal is not encoded
    def get_input_info_flat_sg(self) -> Tuple[List[int], List[int]]:
        return [54], [6]

    def to_np_flat_sg(self, states: List[Cube3State], goals: List[Cube3Goal]) -> List[np.ndarray]:
        return [self.states_to_rows(states)]

    def string_to_action(self, act_str: str) -> Optional[Cube3Action]:
        return Cube3Action(self.atomic_actions.index(act_str)) if act_str in self.atomic_actions else None

    def __repr__(self) -> str:
        return "Cube3()"
