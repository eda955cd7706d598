"""Oracle restatement of beam search over nodes (``beam_v``).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

  InstanceBeam / InstanceNodeBeam      deepxube/pathfinding/beam_search.py:17-66, 106-115
      no CLOSED list (filter_expanded_nodes returns every child), nothing kept between iterations but the selected
      nodes; record_goal keeps the solved popped node with the smallest path cost (first one on ties);
      finished() == has_soln()
  BeamSearchHeurNode._compute_costs    beam_search.py:175-196   logits = -(parent_t_cost + heuristic)
  select_idxs_from_logits              beam_search.py:39-60     all children if there are <= beam_size, else
      np.argpartition(logits, -beam_size)[-beam_size:] -- the SAME numpy call, so this oracle reproduces the reference's
      selection including its (implementation-defined) order and its choice among exact ties
  driver                               deepxube/base/pathfinding.py:338-400 (PathFindNode.step)

temp > 0 (Boltzmann sampling) and eps > 0 (random replacements) are not restated.
"""
from typing import Callable, List, Optional

import numpy as np

from .domains_np import OracleDomain
from .search import _Store

HeurFn = Callable[[np.ndarray, np.ndarray], np.ndarray]


class _BeamInst:
    def __init__(self, domain: OracleDomain, state: np.ndarray, goal: np.ndarray, beam_size: int):
        self.store = _Store(domain.state_len, domain.num_actions, False)
        self.goal = goal
        self.beam_size = beam_size
        self.store.append(state[None, :], np.zeros(1), np.full(1, -1), np.full(1, -1))
        self.nodes_curr: List[int] = [0]
        self.goal_node: Optional[int] = None
        self.num_nodes_generated = 0
        self.itr = 0

    def finished(self) -> bool:
        return self.goal_node is not None

    def path_cost(self) -> float:
        return float(self.store.g[self.goal_node]) if self.goal_node is not None else float("inf")

    def get_path(self):
        acts, n = [], self.goal_node
        while self.store.parent[n] >= 0:
            acts.append(int(self.store.action[n]))
            n = int(self.store.parent[n])
        return None, acts[::-1]


class OracleBeamV:
    def __init__(self, domain: OracleDomain, heur_fn: Optional[HeurFn], beam_size: int = 1):
        self.domain, self.heur_fn, self.beam_size = domain, heur_fn, beam_size
        self.instances: List[_BeamInst] = []

    def add_instances(self, states: np.ndarray, goals: np.ndarray) -> List[_BeamInst]:
        new = [_BeamInst(self.domain, s, g, self.beam_size) for s, g in zip(states, goals)]
        self.instances.extend(new)
        return new

    def all_finished(self) -> bool:
        return all(i.finished() for i in self.instances)

    def step(self) -> None:
        dom, a = self.domain, self.domain.num_actions
        insts = [i for i in self.instances if not i.finished()]
        if not insts:
            return
        ids_by_inst = []
        for inst in insts:
            popped = np.array(inst.nodes_curr, np.int64)
            st, g = inst.store.state[popped], inst.store.g[popped]
            solved = dom.is_solved(st, np.tile(inst.goal, (len(popped), 1)))
            for n, ok in zip(popped.tolist(), solved.tolist()):      # record_goal: strictly smaller path cost replaces
                if ok and (inst.goal_node is None or inst.store.g[inst.goal_node] > inst.store.g[n]):
                    inst.goal_node = n
            children = dom.expand(st)
            ids = inst.store.append(children, np.repeat(g, a) + 1.0, np.repeat(popped, a), np.tile(np.arange(a), len(popped)))
            inst.num_nodes_generated += len(ids)
            ids_by_inst.append(ids)
        # one heuristic call over the children of all instances (PathFindSetHeurV._set_node_vals, pathfinding.py:702-718)
        all_states = np.concatenate([inst.store.state[ids] for inst, ids in zip(insts, ids_by_inst)])
        all_goals = np.concatenate([np.tile(inst.goal, (len(ids), 1)) for inst, ids in zip(insts, ids_by_inst)])
        h = np.zeros(len(all_states)) if self.heur_fn is None else np.asarray(self.heur_fn(all_states, all_goals), np.float64)
        off = 0
        for inst, ids in zip(insts, ids_by_inst):
            hi = h[off: off + len(ids)]
            off += len(ids)
            inst.store.h[ids] = hi
            logits = [-(1.0 + float(x)) for x in hi]
            if len(logits) <= inst.beam_size:
                nxt = list(range(len(logits)))
            else:
                nxt = np.argpartition(logits, -inst.beam_size)[-inst.beam_size:].tolist()
            inst.nodes_curr = [int(ids[k]) for k in nxt]
            inst.itr += 1
