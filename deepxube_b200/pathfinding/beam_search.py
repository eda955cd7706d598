"""`beam_v` / `beam_q`: beam search with a heuristic that prioritises nodes / edges
(deepxube/pathfinding/beam_search.py:17-125, 175-229).

Every iteration expands the current beam, evaluates ALL children (no CLOSED list), and keeps the `beam_size` children with
the smallest `parent_t_cost + heuristic` (the reference takes the beam_size largest logits = -(cost) with
np.argpartition); an instance is finished as soon as a popped node is a goal.  On the device this is the graph_v pipeline
with the CLOSED filter replaced by keep-all, an OPEN list that is emptied every iteration and the selection key 1 + h
(`DX_ALGO_BEAM_V`); `beam_q` is the graph_q pipeline with the same three changes and the key q(node, action)
(`DX_ALGO_BEAM_Q`): the beam_size edges with the smallest q-values are followed.  The selected nodes come out ordered by cost; the reference keeps np.argpartition's
(implementation-defined) order, so the two agree as SETS; with exact cost ties between different states at the selection
boundary the reference's pick is arbitrary and may differ.  temp > 0 (Boltzmann sampling) and eps > 0 (random
replacements) are rejected, `rollout` is not offered.
"""
import re
from typing import Any, Dict, List, Optional, Type

from deepxube_b200.base.domain import ActsEnum
from deepxube_b200.base.factory import Parser
from deepxube_b200.base.pathfind_fns import PFNsHeurQ, PFNsHeurV
from deepxube_b200.factories.pathfinding_factory import pathfinding_factory
from deepxube_b200.pathfinding.graph_search import GraphSearch


@pathfinding_factory.register_class("beam_v")
class BeamSearchHeurNodeActsEnum(GraphSearch):
    is_q = False
    dx_algo = "beam_v"

    def __init__(self, domain: Any, pathfind_fns: Any, beam_size: int = 1, temp: float = 0.0, eps: float = 0.0,
                 rollout: bool = False, **kwargs: Any):
        self._check_sampling(temp, eps, rollout)
        super().__init__(domain, pathfind_fns, batch_size=beam_size, weight=1.0, eps=0.0, **kwargs)
        self.beam_size_default, self.temp_default = int(beam_size), float(temp)

    @staticmethod
    def _check_sampling(temp: Optional[float], eps: Optional[float], rollout: bool = False) -> None:
        if temp or eps or rollout:
            raise ValueError("beam_v on the B200 engine is deterministic top-k selection: temp > 0 (Boltzmann sampling, "
                             "beam_search.py:47-49), eps > 0 (random replacements, :52-59) and rollout are not supported")

    def make_instances(self, states: List[Any], goals: List[Any], inst_infos: Optional[List[Any]] = None,
                       compute_root_vals: bool = True, beam_size: Optional[int] = None, temp: Optional[float] = None,
                       eps: Optional[float] = None):
        self._check_sampling(temp, eps)
        return super().make_instances(states, goals, inst_infos, compute_root_vals, batch_size=beam_size)

    @staticmethod
    def domain_type() -> Type:
        return ActsEnum

    @staticmethod
    def pathfind_functions_type() -> Type:
        return PFNsHeurV

    @staticmethod
    def description() -> str:
        return "Beam search with heuristic that prioritizes nodes"

    def __repr__(self) -> str:
        return f"{type(self).__name__}(beam_size={self.beam_size_default}, temp=0.0, eps=0.0, rollout=False)"


@pathfinding_factory.register_class("beam_q")
class BeamSearchHeurEdgeActsEnum(BeamSearchHeurNodeActsEnum):
    is_q = True
    dx_algo = "beam_q"

    @staticmethod
    def pathfind_functions_type() -> Type:
        return PFNsHeurQ

    @staticmethod
    def description() -> str:
        return "Beam search with heuristic that prioritizes edges"


@pathfinding_factory.register_parser("beam_v")
class BeamSearchNodeParser(Parser):
    def parse(self, args_str: str) -> Dict[str, Any]:
        kwargs: Dict[str, Any] = {}
        for tok in args_str.split("_"):
            mb, mt, me = re.search(r"^(\S+)B$", tok), re.search(r"^(\S+)T", tok), re.search(r"^(\S+)E", tok)
            if mb is not None:
                kwargs["beam_size"] = int(mb.group(1))
            elif mt is not None:
                kwargs["temp"] = float(mt.group(1))
            elif me is not None:
                kwargs["eps"] = float(me.group(1))
            else:
                raise ValueError(f"Unexpected argument {tok!r}")
        kwargs["rollout"] = False
        return kwargs

    def help(self) -> str:
        return ("<int>B (beam size), <float>T (temperature; must be 0 here), <float>E (epsilon; must be 0 here).\n"
                "E.g. beam_v.10B")


@pathfinding_factory.register_parser("beam_q")
class BeamSearchEdgeParser(BeamSearchNodeParser):
    def help(self) -> str:
        return super().help().replace("beam_v", "beam_q")
