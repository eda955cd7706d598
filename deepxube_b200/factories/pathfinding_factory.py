"""pathfinding registry and `--pathfind` resolution (mirror of deepxube/factories/pathfinding_factory.py:9-66):
every registered name that starts with the given prefix and is compatible with (domain, type(pathfind_fns))
is a candidate; exactly one must remain."""
from typing import Any, Dict, List, Optional, Tuple

from deepxube_b200.base.factory import Factory
from deepxube_b200.utils.command_line_utils import get_name_args

pathfinding_factory: Factory = Factory("PathFind")


def get_pathfind_name_kwargs(pathfind: str) -> Tuple[str, Dict[str, Any]]:
    name, args_str = get_name_args(pathfind)
    return name, pathfinding_factory.get_kwargs(name, args_str)


def get_pathfind_from_arg(domain: Any, pathfind_fns: Any, pathfind_name_args: str) -> Tuple[Any, str, str]:
    prefix, args_str = get_name_args(pathfind_name_args)
    compat: List[str] = []
    reasons: List[str] = []
    for name in pathfinding_factory.get_all_class_names():
        if not name.startswith(prefix):
            continue
        reason: Optional[str] = pathfinding_factory.get_type(name).get_incompat_reason(domain, type(pathfind_fns))
        if reason is None:
            compat.append(name)
        else:
            reasons.append(f"{reason} (PathFind name: {name})")
    if len(compat) == 0:
        raise ValueError(f"Could not find any compatible PathFind for Domain {domain} and Functions {pathfind_fns} for PathFind "
                         f"name: {prefix}.\nIncompatibility reasons:\n" + "\n".join(reasons))
    if len(compat) > 1:
        raise ValueError(f"More than 1 compatable PathFind class for Domain {domain} and Functions {pathfind_fns} for PathFind "
                         f"name: {prefix}: {compat}.")
    name = compat[0]
    kwargs = pathfinding_factory.get_kwargs(name, args_str)
    kwargs.update(domain=domain, pathfind_fns=pathfind_fns)
    full = name if args_str is None else f"{name}.{args_str}"
    return pathfinding_factory.build_class(name, kwargs), name, full
