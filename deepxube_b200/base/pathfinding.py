"""PathFind protocol objects: host-side VIEWS over the device-resident search state.

Mirror of the public surface of deepxube/base/pathfinding.py: `Node` (:17-35, the fields callers read),
`get_path(node)` (:93-120), `Instance` (:132-184) and `PathFind` (:194-332).  The search data (node store,
OPEN, CLOSED) lives on the GPU inside a libdxb200 engine; these classes hold handles and materialise nodes
lazily (goal node, current nodes) when a caller asks for them.
"""
from abc import ABC, abstractmethod
from typing import Any, Callable, List, Optional, Tuple, Type

import numpy as np

from deepxube_b200.utils.timing_utils import Times


class Node:
    """Host view of one search node.  `parent` chains are only materialised along solution paths."""
    __slots__ = ["state", "goal", "path_cost", "heuristic", "q_values", "is_solved", "parent_action", "parent_t_cost", "parent",
                 "instance", "node_id"]

    def __init__(self, state: Any, goal: Any, path_cost: float, heuristic: float, is_solved: Optional[bool] = None,
                 parent_action: Any = None, parent_t_cost: Optional[float] = None, parent: Optional["Node"] = None,
                 instance: Any = None, node_id: int = -1):
        self.state, self.goal = state, goal
        self.path_cost, self.heuristic = path_cost, heuristic
        self.q_values = None
        self.is_solved = is_solved
        self.parent_action, self.parent_t_cost, self.parent = parent_action, parent_t_cost, parent
        self.instance, self.node_id = instance, node_id


def get_path(node: Node) -> Tuple[List[Any], List[Any], List[float], float]:
    """(states along the path, actions, transition costs, path cost) from the root to `node`.
    For a goal node handed out by an engine-backed Instance the parent walk runs on the device."""
    if node.parent is None and node.instance is not None and node.path_cost > 0:
        node.instance._materialise_path(node)
    path, actions, tcs = [], [], []
    cur = node
    while cur.parent is not None:
        path.append(cur.state)
        actions.append(cur.parent_action)
        tcs.append(cur.parent_t_cost)
        cur = cur.parent
    path.append(cur.state)
    path.reverse(); actions.reverse(); tcs.reverse()
    assert sum(tcs) == node.path_cost, "sum of transition costs should equal path cost"
    return path, actions, tcs, node.path_cost


class Instance(ABC):
    """View of one problem instance inside a PathFind (deepxube/base/pathfinding.py:132-184)."""

    def __init__(self, root_node: Node, inst_info: Any):
        self.root_node = root_node
        self.inst_info = inst_info
        self.itr = 0
        self.num_nodes_generated = 0
        self.goal_node: Optional[Node] = None

    def has_soln(self) -> bool:
        return self.goal_node is not None

    def path_cost(self) -> float:
        return np.inf if self.goal_node is None else self.goal_node.path_cost

    @abstractmethod
    def frontier_size(self) -> int:
        ...

    @abstractmethod
    def finished(self) -> bool:
        ...


class PathFind(ABC):
    @staticmethod
    @abstractmethod
    def domain_type() -> Type:
        ...

    @staticmethod
    @abstractmethod
    def pathfind_functions_type() -> Type:
        ...

    @staticmethod
    @abstractmethod
    def description() -> str:
        ...

    @classmethod
    def get_incompat_reason(cls, domain: Any, pathfind_fns_t: Type) -> Optional[str]:
        if not isinstance(domain, cls.domain_type()):
            return f"Domain {domain} is not an instance of {cls.domain_type()}"
        if not issubclass(pathfind_fns_t, cls.pathfind_functions_type()):
            return f"PathFind functions type {pathfind_fns_t} is not a subclass of {cls.pathfind_functions_type()}"
        return None

    def __init__(self, domain: Any, pathfind_fns: Any):
        reason = self.get_incompat_reason(domain, type(pathfind_fns))
        if reason is not None:
            raise TypeError(reason)
        self.domain = domain
        self.pathfind_fns = pathfind_fns
        self.instances: List[Any] = []
        self.times = Times()
        self.itr = 0

    @abstractmethod
    def make_instances(self, states: List[Any], goals: List[Any], inst_infos: Optional[List[Any]] = None,
                       compute_root_vals: bool = True) -> List[Any]:
        ...

    @abstractmethod
    def add_instances(self, instances: List[Any]) -> None:
        ...

    @abstractmethod
    def step(self, verbose: bool = False) -> Tuple[List[Node], List[Any]]:
        ...

    def remove_instances(self, test_rem: Callable[[Any], bool]) -> List[Any]:
        removed = [i for i in self.instances if test_rem(i)]
        self.instances = [i for i in self.instances if not test_rem(i)]
        return removed

    def remove_finished_instances(self, itr_max: int) -> List[Any]:
        return self.remove_instances(lambda inst: inst.finished() or inst.itr >= itr_max)

    def expand_states(self, states: List[Any], goals: List[Any], contexts: List[Any]):
        return self.domain.expand(states)

    def get_state_actions(self, states: List[Any], goals: List[Any], contexts: List[Any]):
        return self.domain.get_state_actions(states)
