"""Heuristic-function protocol and containers (mirror of deepxube/base/pathfind_fns.py:20-61, 196-243).

  HeurVFn(states, goals, contexts) -> List[float]
  HeurQFn(states, goals, actions_l, contexts) -> List[List[float]]
  PFNsHeurV(heurv=...) / PFNsHeurQ(heurq=...)  frozen dataclasses handed to a PathFind

A PathFind accepts ANY such callable.  Three kinds get special treatment by the B200 engine:
  * HeurZerosVFn / HeurZerosQFn (the reference's default functions)  -> evaluated on the device as zeros
  * DeviceHeurV / DeviceHeurQ (a resnet_fc network)                   -> weights are loaded into the engine and
    the network runs inside the search iteration, no host round trip
  * anything else                                                     -> called on the host once per iteration
"""
import os
from dataclasses import dataclass
from typing import Any, List, Optional, Protocol, runtime_checkable

import numpy as np


@runtime_checkable
class HeurVFn(Protocol):
    def __call__(self, states: List[Any], goals: List[Any], contexts: List[Any]) -> List[float]:
        ...


@runtime_checkable
class HeurQFn(Protocol):
    def __call__(self, states: List[Any], goals: List[Any], actions_l: List[List[Any]], contexts: List[Any]) -> List[List[float]]:
        ...


@dataclass(frozen=True)
class PFNs:
    pass


@dataclass(frozen=True)
class PFNsHeurV(PFNs):
    heurv: HeurVFn


@dataclass(frozen=True)
class PFNsHeurQ(PFNs):
    heurq: HeurQFn


class HeurZerosVFn:
    def __call__(self, states: List[Any], goals: List[Any], contexts: List[Any]) -> List[float]:
        return [0.0] * len(states)


class HeurZerosQFn:
    def __call__(self, states: List[Any], goals: List[Any], actions_l: List[List[Any]], contexts: List[Any]) -> List[List[float]]:
        return [[0.0] * len(actions) for actions in actions_l]


def default_precision() -> str:
    """fp32 (parity mode) unless DXB200_PRECISION=bf16."""
    p = os.environ.get("DXB200_PRECISION", "fp32").lower()
    if p not in ("fp32", "bf16"):
        raise ValueError("DXB200_PRECISION must be fp32 or bf16")
    return p


class _DeviceHeur:
    """A resnet_fc network evaluated by libdxb200.  Calling it evaluates on the GPU through dx_heur_eval
    (the nnet_batched path, deepxube/pytorch/nnet_utils.py:73-111, incl. nnet_batch_size chunking)."""

    def __init__(self, domain: Any, nnet: Any, precision: Optional[str] = None, nnet_batch_size: Optional[int] = None):
        self.domain, self.nnet = domain, nnet
        self.precision = precision or default_precision()
        self.nnet_batch_size = nnet_batch_size
        self._eng = None
        self._eng_version = -1

    def _engine(self):
        if self._eng is None:
            from deepxube_b200 import _lib
            mode = _lib.DX_HEUR_NET_BF16 if self.precision == "bf16" else _lib.DX_HEUR_NET_FP32
            a = self.domain.get_num_acts()
            is_q = self.nnet.q_fix or getattr(self.nnet, "act_depth", 0) > 0
            eng = _lib.Engine(self.domain.dx_name, self.domain.dx_dim, "graph_q" if is_q else "graph_v", mode,
                              max_instances=1, max_pop_total=max(1, 16384 // (1 if self.nnet.q_fix else a)), max_nodes=1024)
            load_nnet_into_engine(eng, self.nnet)
            self._eng, self._eng_version = eng, getattr(self.nnet, "version", 0)
        elif getattr(self.nnet, "version", 0) != self._eng_version:      # nnet.load_state_dict since the upload
            load_nnet_into_engine(self._eng, self.nnet)
            self._eng_version = getattr(self.nnet, "version", 0)
        return self._eng

    def _eval(self, states: List[Any], goals: List[Any]) -> np.ndarray:
        n = len(states)
        if n == 0:
            return np.zeros((0, self.nnet.out_dim * max(getattr(self.nnet, "act_depth", 0), 1)), np.float32)
        rows, grows = self.domain.states_to_rows(states), self.domain.goals_to_rows(goals)
        step = self.nnet_batch_size or n
        od = self.nnet.out_dim * max(getattr(self.nnet, "act_depth", 0), 1)          # action-as-input: one value per action
        outs = [self._engine().heur_eval(rows[s:s + step], grows[s:s + step], od) for s in range(0, n, step)]
        return np.concatenate(outs, axis=0)


class DeviceHeurV(_DeviceHeur):
    def __call__(self, states: List[Any], goals: List[Any], contexts: List[Any]) -> List[float]:
        return self._eval(states, goals)[:, 0].astype(np.float64).tolist()     # already clipped at 0 on the device


class DeviceHeurQ(_DeviceHeur):
    def __call__(self, states: List[Any], goals: List[Any], actions_l: List[List[Any]], contexts: List[Any]) -> List[List[float]]:
        assert len({len(a) for a in actions_l}) <= 1, "num actions should be the same for all instances"
        return self._eval(states, goals).astype(np.float64).tolist()


def load_nnet_into_engine(eng: Any, nnet: Any) -> None:
    from deepxube_b200 import _lib
    blob, info = _lib.pack_state_dict(nnet.state_dict())
    eng.load_weights(nnet.in_count, nnet.one_hot_depth, info["res_dim"], info["num_blocks"], info["out_dim"], info["batch_norm"], blob,
                     act_depth=getattr(nnet, "act_depth", 0))
