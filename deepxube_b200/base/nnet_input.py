"""NN input descriptions for the flat (one-hot) network inputs of the in-scope domains.

Mirror of the part of deepxube/base/nnet_input.py the hot path touches: `FlatIn.get_input_info() -> (dims,
one_hot_depths)` (:18-56), the state-goal input `flat_sg` (:186-233) and the fixed-action Q input
`flat_sg_actfix` whose `to_np` appends the [N, n_act] action-index matrix (:262-280); for grid the names are
`grid_flat_in` / `grid_flat_in_qfix` (deepxube/domains/grid.py:126-144).
"""
from typing import Any, List, Optional, Tuple

import numpy as np

from deepxube_b200.utils.command_line_utils import get_name_args


class NNetInput:
    def __init__(self, domain: Any):
        self.domain = domain


class FlatIn(NNetInput):
    def get_input_info(self) -> Tuple[List[int], List[int]]:
        return self.domain.get_input_info_flat_sg()


class StateGoalIn(FlatIn):
    def to_np(self, states: List[Any], goals: List[Any]) -> List[np.ndarray]:
        return self.domain.to_np_flat_sg(states, goals)


class StateGoalActIn(FlatIn):
    """action-as-input Q networks (`heurq_in`): the flat state-goal input plus ONE action index per row whose one-hot
    depth is the number of actions (deepxube/base/nnet_input.py:283-296, `flat_sga`; grid: `grid_flat_in_actin`,
    deepxube/domains/grid.py:147-155)."""

    def get_input_info(self) -> Tuple[List[int], List[int]]:
        dims, depths = self.domain.get_input_info_flat_sg()
        return list(dims) + [1], list(depths) + [self.domain.get_num_acts()]

    def to_np(self, states: List[Any], goals: List[Any], actions: List[Any]) -> List[np.ndarray]:
        acts = np.expand_dims(np.array([self.domain.action_index(a) for a in actions]), 1)
        return self.domain.to_np_flat_sg(states, goals) + [acts]


class StateGoalActFixIn(FlatIn):
    def to_np(self, states: List[Any], goals: List[Any], actions_l: List[List[Any]]) -> List[np.ndarray]:
        acts = np.array([[self.domain.action_index(a) for a in al] for al in actions_l])
        return self.domain.to_np_flat_sg(states, goals) + [acts]


# domain name -> {nnet input name: class}; order matters: the first compatible one is picked
# (deepxube/factories/pathfind_fns_factory.py:23-39)
_NNET_INPUTS = {
    "cube3": {"flat_sg": StateGoalIn, "flat_sg_actfix": StateGoalActFixIn, "flat_sga": StateGoalActIn},
    "npuzzle": {"flat_sg": StateGoalIn},                      # no fixed-action input: heurq_fixout unavailable (npuzzle.py:55)
    "grid": {"grid_flat_in": StateGoalIn, "grid_flat_in_qfix": StateGoalActFixIn, "grid_flat_in_actin": StateGoalActIn},
    "lightsout": {"flat_sg": StateGoalIn, "flat_sg_actfix": StateGoalActFixIn, "flat_sga": StateGoalActIn},      # lightsout.py:60-64
}


def get_domain_nnet_input_keys(domain_name: str) -> List[str]:
    return list(_NNET_INPUTS.get(domain_name, {}))


def get_nnet_input_t(domain_name: str, nnet_input_name: str):
    try:
        return _NNET_INPUTS[domain_name][nnet_input_name]
    except KeyError:
        raise ValueError(f"Unknown nnet input {nnet_input_name!r} for domain {domain_name!r}. "
                         f"Available: {get_domain_nnet_input_keys(domain_name)}")


def get_nnet_input_from_arg(domain: Any, domain_name: str, nnet_input_name_args: str) -> NNetInput:
    name, _ = get_name_args(nnet_input_name_args)
    return get_nnet_input_t(domain_name, name)(domain)
