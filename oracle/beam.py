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

  beam_q: InstanceEdgeBeam / BeamSearchHeurEdge     beam_search.py:117-125, 199-214; driver pathfinding.py:471-525
      every (popped node, action) is an edge with cost q(node, action); the beam_size edges with the smallest q are followed
      (logits = -q_val), their child nodes are generated, evaluated, and become the next popped nodes

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


class OracleBeamQ:
    """heur_fn(states, goals) -> float64 [N, A] (q-values of every action)."""

    def __init__(self, domain: OracleDomain, heur_fn: Optional[HeurFn], beam_size: int = 1):
        self.domain, self.heur_fn, self.beam_size = domain, heur_fn, beam_size
        self.instances: List[_BeamInst] = []

    def _q(self, states: np.ndarray, goals: np.ndarray) -> np.ndarray:
        if self.heur_fn is None:
            return np.zeros((len(states), self.domain.num_actions))
        return np.asarray(self.heur_fn(states, goals), np.float64)

    def add_instances(self, states: np.ndarray, goals: np.ndarray) -> List[_BeamInst]:
        new = [_BeamInst(self.domain, s, g, self.beam_size) for s, g in zip(states, goals)]
        q = self._q(np.asarray(states), np.asarray(goals))          # root values (make_instances, beam_search.py:203-204)
        for inst, qi in zip(new, q):
            inst.q = {0: qi}
        self.instances.extend(new)
        return new

    def all_finished(self) -> bool:
        return all(i.finished() for i in self.instances)

    def step(self) -> None:
        dom, a = self.domain, self.domain.num_actions
        insts = [i for i in self.instances if not i.finished()]
        if not insts:
            return
        new_by_inst = []
        for inst in insts:
            popped = np.array(inst.nodes_curr, np.int64)
            st, g = inst.store.state[popped], inst.store.g[popped]
            solved = dom.is_solved(st, np.tile(inst.goal, (len(popped), 1)))
            for n, ok in zip(popped.tolist(), solved.tolist()):
                if ok and (inst.goal_node is None or inst.store.g[inst.goal_node] > inst.store.g[n]):
                    inst.goal_node = n
            # edges in (node, action) order with their q-values; select by the largest logits = -q
            nodes_e = np.repeat(popped, a)
            acts_e = np.tile(np.arange(a), len(popped))
            logits = [-float(inst.q[int(n)][int(k)]) for n, k in zip(nodes_e, acts_e)]
            if len(logits) <= inst.beam_size:
                nxt = list(range(len(logits)))
            else:
                nxt = np.argpartition(logits, -inst.beam_size)[-inst.beam_size:].tolist()
            pn, pa = nodes_e[nxt], acts_e[nxt]
            children = dom.next_state(inst.store.state[pn], pa)
            ids = inst.store.append(children, inst.store.g[pn] + 1.0, pn, pa)
            inst.num_nodes_generated += len(ids)
            new_by_inst.append(ids)
        all_states = np.concatenate([inst.store.state[ids] for inst, ids in zip(insts, new_by_inst)])
        all_goals = np.concatenate([np.tile(inst.goal, (len(ids), 1)) for inst, ids in zip(insts, new_by_inst)])
        q = self._q(all_states, all_goals)
        off = 0
        for inst, ids in zip(insts, new_by_inst):
            inst.q = {int(n): q[off + k] for k, n in enumerate(ids)}
            inst.store.h[ids] = q[off: off + len(ids)].min(axis=1)
            off += len(ids)
            inst.nodes_curr = [int(n) for n in ids]
            inst.itr += 1
