"""Solution check (deepxube/pathfinding/utils/performance.py:107-112): replay the actions through next_state."""
from typing import Any, List


def is_valid_soln(state: Any, goal: Any, soln: List[Any], domain: Any) -> bool:
    cur = state
    for action in soln:
        cur = domain.next_state([cur], [action])[0][0]
    return bool(domain.is_solved([cur], [goal])[0])
