"""Domain API of the hot path: State / Action / Goal and the mixins the search engine and the CLI use.

Same names, argument meaning and return shapes as deepxube/base/domain.py:
  State / Action / Goal (hash + eq contracts)                       :22-69
  Domain.next_state / is_solved / sample_problem_instances          :78-207
  ActsEnum.get_state_actions / expand (state-major, action-minor)   :278-312
  ActsEnumFixed.get_actions_fixed / get_state_actions / get_num_acts :315-336
  StartGoalWalkable.sample_start_states / sample_goal_from_state / sample_problem_instances   :488-520
  GoalStartRevWalkable.sample_goalstate_goal_pairs / sample_problem_instances                 :559-585
  ActsRev.sample_rev_state (used by the reference's test_domains.py::test_actsrev)

What differs: states live as packed uint8 rows, and next_state / expand / is_solved run on the GPU through
libdxb200 (`PackedDomain`).  There is no CPU implementation of these three methods: without the CUDA
library they raise.  Only problem-instance GENERATION (random walks that produce inputs for the search)
runs on the host, from the same move tables the kernels use (deepxube_b200/domains/tables.py).
"""
import random
from abc import ABC, abstractmethod
from typing import Any, Generic, List, Optional, Tuple, TypeVar

import numpy as np

from deepxube_b200.utils import misc_utils
from deepxube_b200.utils.timing_utils import Times


class State(ABC):
    @abstractmethod
    def __hash__(self) -> int:
        ...

    @abstractmethod
    def __eq__(self, other: object) -> bool:
        ...


class Action(ABC):
    @abstractmethod
    def __hash__(self) -> int:
        ...

    @abstractmethod
    def __eq__(self, other: object) -> bool:
        ...


class Goal(ABC):
    pass


S = TypeVar("S", bound=State)
A = TypeVar("A", bound=Action)
G = TypeVar("G", bound=Goal)


class Domain(ABC, Generic[S, A, G]):
    def __init__(self, *args: Any, **kwargs: Any) -> None:
        self.nnet_par_dict: dict = {}

    @abstractmethod
    def sample_problem_instances(self, num_steps_l: List[int], times: Optional[Times] = None) -> Tuple[List[S], List[G]]:
        ...

    @abstractmethod
    def sample_state_action(self, states: List[S]) -> List[A]:
        ...

    @abstractmethod
    def next_state(self, states: List[S], actions: List[A]) -> Tuple[List[S], List[float]]:
        ...

    @abstractmethod
    def is_solved(self, states: List[S], goals: List[G]) -> List[bool]:
        ...

    def sample_next_state(self, states: List[S]) -> Tuple[List[S], List[A], List[float]]:
        actions = self.sample_state_action(states)
        nxt, tcs = self.next_state(states, actions)
        return nxt, actions, tcs

    def random_walk(self, states: List[S], num_steps_l: List[int]) -> Tuple[List[S], List[List[A]], List[float]]:
        """Random walk of num_steps_l[i] steps from states[i] -> (end states, actions taken, path costs)."""
        cur = list(states)
        actions_l: List[List[A]] = [[] for _ in states]
        costs = [0.0 for _ in states]
        left = np.array(num_steps_l, dtype=int)
        while np.any(left > 0):
            idxs = np.nonzero(left > 0)[0].tolist()
            nxt, acts, tcs = self.sample_next_state([cur[i] for i in idxs])
            for i, s, a, tc in zip(idxs, nxt, acts, tcs):
                cur[i] = s
                actions_l[i].append(a)
                costs[i] += tc
            left[idxs] -= 1
        return cur, actions_l, costs


class ActsFixed(Domain[S, A, G]):
    @abstractmethod
    def sample_action(self, num: int) -> List[A]:
        ...

    def sample_state_action(self, states: List[S]) -> List[A]:
        return self.sample_action(len(states))


class ActsEnum(Domain[S, A, G]):
    @abstractmethod
    def get_state_actions(self, states: List[S]) -> List[List[A]]:
        ...

    def sample_state_action(self, states: List[S]) -> List[A]:
        return [random.choice(acts) for acts in self.get_state_actions(states)]

    def expand(self, states: List[S]) -> Tuple[List[List[S]], List[List[A]], List[List[float]]]:
        """All children of every state: (children[i][a], actions[i][a], transition costs[i][a])."""
        actions_l = self.get_state_actions(states)
        flat_actions, split = misc_utils.flatten(actions_l)
        flat_states: List[S] = []
        for s, acts in zip(states, actions_l):
            flat_states.extend([s] * len(acts))
        nxt, tcs = self.next_state(flat_states, flat_actions)
        return misc_utils.unflatten(nxt, split), actions_l, misc_utils.unflatten(tcs, split)


class ActsEnumFixed(ActsEnum[S, A, G], ActsFixed[S, A, G]):
    def sample_action(self, num: int) -> List[A]:
        fixed = self.get_actions_fixed()
        return [random.choice(fixed) for _ in range(num)]

    def sample_state_action(self, states: List[S]) -> List[A]:
        return self.sample_action(len(states))

    def get_state_actions(self, states: List[S]) -> List[List[A]]:
        return [self.get_actions_fixed().copy() for _ in range(len(states))]

    @abstractmethod
    def get_actions_fixed(self) -> List[A]:
        ...

    def get_num_acts(self) -> int:
        return len(self.get_actions_fixed())


class ActsRev(Domain[S, A, G]):
    @abstractmethod
    def sample_rev_state(self, states: List[S]) -> Tuple[List[S], List[A], List[float]]:
        """-> (previous states, the action that leads from each previous state back to the given one, costs)"""
        ...


class GoalSampleableFromState(Domain[S, A, G]):
    @abstractmethod
    def sample_goal_from_state(self, states_start: Optional[List[S]], states_goal: List[S]) -> List[G]:
        ...


class StartGoalWalkable(GoalSampleableFromState[S, A, G]):
    @abstractmethod
    def sample_start_states(self, num_states: int) -> List[S]:
        ...

    def sample_problem_instances(self, num_steps_l: List[int], times: Optional[Times] = None) -> Tuple[List[S], List[G]]:
        import time
        times = times if times is not None else Times()
        t0 = time.time()
        starts = self.sample_start_states(len(num_steps_l))
        times.record_time("sample_start_states", time.time() - t0)
        t0 = time.time()
        ends = self.random_walk(starts, num_steps_l)[0]
        times.record_time("random_walk", time.time() - t0)
        t0 = time.time()
        goals = self.sample_goal_from_state(starts, ends)
        times.record_time("sample_goal", time.time() - t0)
        return starts, goals


class GoalStartRevWalkable(Domain[S, A, G]):
    @abstractmethod
    def sample_goalstate_goal_pairs(self, num: int) -> Tuple[List[S], List[G]]:
        ...

    @abstractmethod
    def random_walk_rev_no_path_cost(self, states: List[S], num_steps_l: List[int]) -> List[S]:
        ...

    def sample_problem_instances(self, num_steps_l: List[int], times: Optional[Times] = None) -> Tuple[List[S], List[G]]:
        import time
        times = times if times is not None else Times()
        t0 = time.time()
        goal_states, goals = self.sample_goalstate_goal_pairs(len(num_steps_l))
        times.record_time("sample_goalstate_goal_pairs", time.time() - t0)
        t0 = time.time()
        starts = self.random_walk_rev_no_path_cost(goal_states, num_steps_l)
        times.record_time("random_walk", time.time() - t0)
        return starts, goals


# ---------------------------------------------------------------------------------------------------
class PackedDomain(ActsEnumFixed[S, A, G]):
    """Domain whose states are packed uint8 rows handled by libdxb200's kernels.

    Subclasses define `dx_name` / `dx_dim`, the row <-> object conversions, and `_walk_rows` (host-side
    random walk used ONLY to generate problem instances)."""
    dx_name: str = ""
    dx_dim: int = 0

    # -- object <-> row conversion --------------------------------------------------------------------
    @abstractmethod
    def states_to_rows(self, states: List[S]) -> np.ndarray:
        ...

    @abstractmethod
    def rows_to_states(self, rows: np.ndarray) -> List[S]:
        ...

    @abstractmethod
    def goals_to_rows(self, goals: List[G]) -> np.ndarray:
        ...

    @abstractmethod
    def action_index(self, action: A) -> int:
        ...

    # -- device-backed domain API ---------------------------------------------------------------------
    def next_state(self, states: List[S], actions: List[A]) -> Tuple[List[S], List[float]]:
        from deepxube_b200 import _lib
        if len(states) == 0:
            return [], []
        rows = _lib.next_state(self.dx_name, self.dx_dim, self.states_to_rows(states), [self.action_index(a) for a in actions])
        return self.rows_to_states(rows), [1.0] * len(states)

    def expand(self, states: List[S]) -> Tuple[List[List[S]], List[List[A]], List[List[float]]]:
        from deepxube_b200 import _lib
        n, a = len(states), self.get_num_acts()
        if n == 0:
            return [], [], []
        children = self.rows_to_states(_lib.expand(self.dx_name, self.dx_dim, self.states_to_rows(states)))
        fixed = self.get_actions_fixed()
        return ([children[i * a:(i + 1) * a] for i in range(n)], [fixed.copy() for _ in range(n)],
                [[1.0] * a for _ in range(n)])

    def is_solved(self, states: List[S], goals: List[G]) -> List[bool]:
        from deepxube_b200 import _lib
        if len(states) == 0:
            return []
        return _lib.is_solved(self.dx_name, self.dx_dim, self.states_to_rows(states), self.goals_to_rows(goals)).tolist()

    # -- NN input description (deepxube/base/nnet_input.py FlatIn.get_input_info) -------------------------
    def get_input_info_flat(self) -> Tuple[List[int], List[int]]:
        from deepxube_b200 import _lib
        _, _, in_count, depth = _lib.domain_info(self.dx_name, self.dx_dim)
        return [in_count], [depth]
