"""grid domain (host view): GridState(robot_x, robot_y) / GridGoal / GridAction, registered as "grid" with
parser `grid.<dim>d` (deepxube/domains/grid.py:24-116).  Actions: 0 x-1, 1 x+1, 2 y-1, 3 y+1, clamped to the
board; start-goal walkable (start = random cell, goal = end of a random walk)."""
import random
from typing import List, Optional, Tuple

import numpy as np

from deepxube_b200.base.domain import Action, Goal, PackedDomain, StartGoalWalkable, State
from deepxube_b200.base.factory import DelimParser
from deepxube_b200.factories.domain_factory import domain_factory


class GridState(State):
    def __init__(self, robot_x: int, robot_y: int):
        self.robot_x, self.robot_y = int(robot_x), int(robot_y)

    def __hash__(self) -> int:
        return hash(self.robot_x + self.robot_y)      # collisions by design, as in the reference (grid.py:29-30)

    def __eq__(self, other: object) -> bool:
        if isinstance(other, GridState):
            return self.robot_x == other.robot_x and self.robot_y == other.robot_y
        return NotImplemented

    def __repr__(self) -> str:
        return f"GridState({self.robot_x}, {self.robot_y})"


class GridGoal(Goal):
    def __init__(self, robot_x: int, robot_y: int):
        self.robot_x, self.robot_y = int(robot_x), int(robot_y)


class GridAction(Action):
    def __init__(self, action: int):
        self.action = action

    def __hash__(self) -> int:
        return self.action

    def __eq__(self, other: object) -> bool:
        if isinstance(other, GridAction):
            return self.action == other.action
        return NotImplemented

    def __repr__(self) -> str:
        return ["UP", "DOWN", "LEFT", "RIGHT"][self.action]


@domain_factory.register_class("grid")
class Grid(PackedDomain[GridState, GridAction, GridGoal], StartGoalWalkable[GridState, GridAction, GridGoal]):
    dx_name = "grid"

    def __init__(self, dim: int = 7):
        super().__init__()
        if not 1 <= dim <= 255:
            raise ValueError("grid: dim must be in 1..255")
        self.dim, self.dx_dim = dim, dim
        self.actions_fixed: List[GridAction] = [GridAction(a) for a in range(4)]

    def states_to_rows(self, states: List[GridState]) -> np.ndarray:
        return np.array([[s.robot_x, s.robot_y] for s in states], dtype=np.uint8).reshape(len(states), 2)

    def rows_to_states(self, rows: np.ndarray) -> List[GridState]:
        return [GridState(int(r[0]), int(r[1])) for r in rows]

    def goals_to_rows(self, goals: List[GridGoal]) -> np.ndarray:
        return np.array([[g.robot_x, g.robot_y] for g in goals], dtype=np.uint8).reshape(len(goals), 2)

    def action_index(self, action: GridAction) -> int:
        return action.action

    def get_actions_fixed(self) -> List[GridAction]:
        return self.actions_fixed.copy()

    def sample_start_states(self, num_states: int) -> List[GridState]:
        return [GridState(random.randint(0, self.dim - 1), random.randint(0, self.dim - 1)) for _ in range(num_states)]

    def sample_goal_from_state(self, states_start: Optional[List[GridState]], states_goal: List[GridState]) -> List[GridGoal]:
        return [GridGoal(s.robot_x, s.robot_y) for s in states_goal]

    def random_walk(self, states: List[GridState], num_steps_l: List[int]):
        """Host-side walk for instance generation (same arithmetic as the kernel, grid.py:74-86)."""
        out, acts_l = [], []
        d = self.dim
        for s, k in zip(states, num_steps_l):
            x, y, acts = s.robot_x, s.robot_y, []
            for _ in range(int(k)):
                a = random.randrange(4)
                acts.append(GridAction(a))
                if a == 0:
                    x = max(x - 1, 0)
                elif a == 1:
                    x = min(x + 1, d - 1)
                elif a == 2:
                    y = max(y - 1, 0)
                else:
                    y = min(y + 1, d - 1)
            out.append(GridState(x, y))
            acts_l.append(acts)
        return out, acts_l, [float(len(a)) for a in acts_l]

    # NN input (deepxube/domains/grid.py:126-133): [sx, sy, gx, gy], one-hot depth dim
    def get_input_info_flat_sg(self) -> Tuple[List[int], List[int]]:
        return [4], [self.dim]

    def to_np_flat_sg(self, states: List[GridState], goals: List[GridGoal]) -> List[np.ndarray]:
        return [np.concatenate([self.states_to_rows(states), self.goals_to_rows(goals)], axis=1)]

    def string_to_action(self, act_str: str) -> Optional[GridAction]:
        return GridAction("wsad".index(act_str)) if act_str in ("w", "s", "a", "d") else None

    def __repr__(self) -> str:
        return f"Grid(dim={self.dim})"


@domain_factory.register_parser("grid")
class GridParser(DelimParser):
    def __init__(self) -> None:
        super().__init__()
        self.add_argument("d", "dim", int, "dimensionality of grid")

    @property
    def delim(self) -> str:
        return "_"
