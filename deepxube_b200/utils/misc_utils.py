"""flatten / unflatten of lists of lists (deepxube/utils/misc_utils.py:8-28)."""
from typing import Any, List, Tuple


def flatten(data: List[List[Any]]) -> Tuple[List[Any], List[int]]:
    split_idxs: List[int] = []
    flat: List[Any] = []
    for sub in data:
        flat.extend(sub)
        split_idxs.append(len(flat))
    return flat, split_idxs[:-1]


def unflatten(data, split_idxs: List[int]) -> List[List[Any]]:
    out, start = [], 0
    for end in list(split_idxs) + [len(data)]:
        out.append(list(data[start:end]))
        start = end
    return out
