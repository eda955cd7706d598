"""Concrete pathfind-function containers and the `heurv` / `heurq_fixout` function kinds
(mirror of deepxube/pathfind_fns/fns_core.py:19-90 and deepxube/base/pathfind_fns.py:188-243)."""
from dataclasses import dataclass
from typing import Any, Optional

from deepxube_b200.base.domain import ActsEnumFixed, Domain
from deepxube_b200.base.nnet_input import StateGoalActFixIn, StateGoalActIn, StateGoalIn, get_nnet_input_from_arg, get_nnet_input_t
from deepxube_b200.base.pathfind_fns import (DeviceHeurQ, DeviceHeurV, HeurZerosQFn, HeurZerosVFn, PFNsHeurQ, PFNsHeurV)
from deepxube_b200.factories.nnet_factory import deepxube_nnet_factory
from deepxube_b200.factories.pathfind_fns_factory import deepxube_nnet_par_factory, pathfind_fns_factory
from deepxube_b200.utils.command_line_utils import get_name_args


@pathfind_fns_factory.register
@dataclass(frozen=True)
class PFNsHeurVC(PFNsHeurV):
    pass


@pathfind_fns_factory.register
@dataclass(frozen=True)
class PFNsHeurQC(PFNsHeurQ):
    pass


class _HeurNNetPar:
    """Describes one heuristic function: which domain / nnet input / network it uses and how to build it."""
    field_name = ""
    q_fix = False

    @staticmethod
    def domain_type():
        return Domain

    @staticmethod
    def nnet_input_type():
        return StateGoalIn

    @classmethod
    def get_incompat_reason(cls, domain: Any, nnet_input_t: Optional[type], nnet_t: Optional[type]) -> Optional[str]:
        if not isinstance(domain, cls.domain_type()):
            return f"Domain {domain} is not an instance of {cls.domain_type()}"
        if nnet_input_t is not None and not issubclass(nnet_input_t, cls.nnet_input_type()):
            return f"NNetInput type {nnet_input_t} is not a subclass of {cls.nnet_input_type()}"
        return None

    def __init__(self, domain: Any, domain_name: str, nnet_input_name_args: Optional[str], nnet_name_args: Optional[str]):
        nnet_input_t = get_nnet_input_t(domain_name, get_name_args(nnet_input_name_args)[0]) if nnet_input_name_args else None
        nnet_t = deepxube_nnet_factory.get_type(get_name_args(nnet_name_args)[0]) if nnet_name_args else None
        reason = self.get_incompat_reason(domain, nnet_input_t, nnet_t)
        if reason is not None:
            raise TypeError(reason)
        self.domain, self.domain_name = domain, domain_name
        self.nnet_input_name_args, self.nnet_name_args = nnet_input_name_args, nnet_name_args
        self.nnet_file: Optional[str] = None

    def get_field_name(self) -> str:
        return self.field_name

    def set_nnet_file(self, nnet_file: str) -> None:
        self.nnet_file = nnet_file

    def get_nnet_input(self):
        assert self.nnet_input_name_args is not None
        return get_nnet_input_from_arg(self.domain, self.domain_name, self.nnet_input_name_args)

    def _out_dim(self) -> int:
        return 1

    def get_nnet(self):
        assert self.nnet_name_args is not None
        name, args = get_name_args(self.nnet_name_args)
        kwargs = deepxube_nnet_factory.get_kwargs(name, args)
        kwargs.update(nnet_input=self.get_nnet_input(), q_fix=self.q_fix, out_dim=self._out_dim())
        return deepxube_nnet_factory.build_class(name, kwargs)

    def __repr__(self) -> str:
        return f"{type(self).__name__}()"


@deepxube_nnet_par_factory.register_class("heurv")
class HeurVNNetParC(_HeurNNetPar):
    field_name = "heurv"

    def get_default_fn(self):
        return HeurZerosVFn()

    def get_nnet_fn(self, nnet: Any, batch_size: Optional[int], precision: Optional[str] = None):
        return DeviceHeurV(self.domain, nnet, precision, batch_size)


@deepxube_nnet_par_factory.register_class("heurq_fixout")
class HeurQNNetParFixOut(_HeurNNetPar):
    """One evaluation per state, one output per (fixed) action (fns_core.py:60-90)."""
    field_name = "heurq"
    q_fix = True

    @staticmethod
    def domain_type():
        return ActsEnumFixed

    @staticmethod
    def nnet_input_type():
        return StateGoalActFixIn

    def _out_dim(self) -> int:
        return self.domain.get_num_acts()

    def get_default_fn(self):
        return HeurZerosQFn()

    def get_nnet_fn(self, nnet: Any, batch_size: Optional[int], precision: Optional[str] = None):
        return DeviceHeurQ(self.domain, nnet, precision, batch_size)


@deepxube_nnet_par_factory.register_class("heurq_in")
class HeurQNNetParIn(_HeurNNetPar):
    """Action-as-input Q function (fns_core.py:99-133): the network takes (state, goal, action) and returns ONE value; it is
    evaluated once per (state, action) -- every state repeated for each of its actions -- and the values are regrouped per
    state.  On the device the (node, action) rows are never materialised: the input kernels read node r / A and set
    the action feature r % A.  The engine needs a fixed action set (the reference accepts any Domain)."""
    field_name = "heurq"
    q_fix = False

    @staticmethod
    def domain_type():
        return ActsEnumFixed

    @staticmethod
    def nnet_input_type():
        return StateGoalActIn

    def get_default_fn(self):
        return HeurZerosQFn()

    def get_nnet_fn(self, nnet: Any, batch_size: Optional[int], precision: Optional[str] = None):
        return DeviceHeurQ(self.domain, nnet, precision, batch_size)
