"""Batch weighted A* (`graph_v`) and batch weighted Q* (`graph_q`) on the B200 engine.

Registered under the reference's names with the reference's constructor kwargs and argument parser
(deepxube/pathfinding/graph_search.py:83-92, 182-209, 250-274): `graph.<int>B_<float>W_<float>E`,
defaults batch_size=1, weight=1.0, eps=0.0.  Semantics reproduced by the device kernels: SURVEY.md App. B.
eps > 0 (random pops) is rejected -- it is not part of the accelerated path.

Host protocol kept: make_instances / add_instances / step / instances / times / itr / Instance.finished(),
goal_node, num_nodes_generated, itr, path_cost(), has_soln(), frontier_size(), lb, ub; get_path(goal_node).
The heuristic callable decides the engine mode (deepxube_b200/base/pathfind_fns.py): zeros and resnet_fc
networks run inside the device iteration; any other callable is invoked on the host once per step.
"""
import os
import re
import time
import weakref
from typing import Any, Dict, List, Optional, Tuple, Type

import numpy as np

from deepxube_b200 import _lib
from deepxube_b200.base.domain import ActsEnum, PackedDomain
from deepxube_b200.base.factory import Parser
from deepxube_b200.base.pathfind_fns import (DeviceHeurQ, DeviceHeurV, HeurZerosQFn, HeurZerosVFn, PFNsHeurQ, PFNsHeurV,
                                             load_nnet_into_engine)
from deepxube_b200.base.pathfinding import Instance, Node, PathFind
from deepxube_b200.factories.pathfinding_factory import pathfinding_factory


class InstanceGraph(Instance):
    """InstanceGraph view (graph_search.py:20-80): lb / ub / batch_size / weight / eps + engine-backed status."""

    def __init__(self, root_node: Node, inst_info: Any, batch_size: int = 1, weight: float = 1.0, eps: float = 0.0):
        super().__init__(root_node, inst_info)
        self.batch_size, self.weight, self.eps = int(batch_size), float(weight), float(eps)
        self.ub: float = np.inf
        self.lb: float = 0.0
        self._finished = False
        self._open_size = 0
        self._num_curr = 1
        self._owner: Optional["GraphSearch"] = None
        self._slot = -1
        self._goal_id = -1

    def frontier_size(self) -> int:
        return self._open_size

    def finished(self) -> bool:
        return self._finished

    def get_nodes(self) -> List[Node]:
        """The nodes the next step will expand, in pop order (Instance.get_nodes, base/pathfinding.py:147-148)."""
        if self._owner is None or self._slot < 0:
            return [self.root_node]
        st, g, h = self._owner._engine.curr_nodes(self._slot, self._num_curr + 1)
        states = self._owner.domain.rows_to_states(st)
        return [Node(s, self.root_node.goal, float(gi), float(hi), instance=None) for s, gi, hi in zip(states, g, h)]

    def _update(self, s: Any) -> None:
        self.lb, self.ub = s.lb, s.ub
        self.itr, self.num_nodes_generated = s.itr, s.num_nodes_generated
        self._finished, self._open_size, self._num_curr = bool(s.finished), s.open_size, s.num_curr
        if s.goal_node >= 0 and s.goal_node != self._goal_id:
            self._goal_id = s.goal_node
            # cost of the goal node = ub (record_goal, graph_search.py:35-41); the path is materialised on demand
            self.goal_node = Node(None, self.root_node.goal, float(s.ub), 0.0, is_solved=True, instance=self, node_id=s.goal_node)

    def _materialise_path(self, node: Node) -> None:
        acts, rows = self._owner._engine.get_path(self._slot, cap=max(int(node.path_cost) + 8, 64), with_states=True)
        dom = self._owner.domain
        states = dom.rows_to_states(rows)
        fixed = dom.get_actions_fixed()
        prev = Node(states[0], node.goal, 0.0, self.root_node.heuristic)
        for k, a in enumerate(acts):
            last = k == len(acts) - 1
            cur = node if last else Node(states[k + 1], node.goal, float(k + 1), 0.0)
            cur.state, cur.parent, cur.parent_action, cur.parent_t_cost = states[k + 1], prev, fixed[a], 1.0
            prev = cur
        if not acts:
            node.state = states[0]


class GraphSearch(PathFind):
    is_q = False
    dx_algo: Optional[str] = None        # engine algorithm name; None: graph_q / graph_v from is_q

    def __init__(self, domain: Any, pathfind_fns: Any, batch_size: int = 1, weight: float = 1.0, eps: float = 0.0,
                 max_nodes: Optional[int] = None, max_instances: Optional[int] = None, precision: Optional[str] = None):
        super().__init__(domain, pathfind_fns)
        if not isinstance(domain, PackedDomain):
            raise TypeError(f"{type(self).__name__}: domain {domain} has no B200 expansion kernel (supported: cube3, npuzzle, grid); "
                            "deepxube_b200 has no CPU fallback")
        self.batch_size_default, self.weight_default, self.eps_default = int(batch_size), float(weight), float(eps)
        self._check_eps(self.eps_default)
        self._max_nodes = max_nodes or int(os.environ.get("DXB200_MAX_NODES", "0")) or None
        self._max_instances = max_instances
        self._precision = precision
        self._engine: Optional[_lib.Engine] = None
        self._fn = pathfind_fns.heurq if self.is_q else pathfind_fns.heurv
        self._pending: List[InstanceGraph] = []
        self._max_backups = 0
        # stage_times: split every step's device time over the reference's Times keys (expand / filt / heur / cost / pushpop,
        # deepxube/base/pathfinding.py:426-744, graph_search.py:156) from CUDA events recorded inside the captured
        # iteration (4-5 % overhead); off: the whole step is booked as "search"
        self.stage_times_on = os.environ.get("DXB200_STAGE_TIMES", "0") == "1"
        self._stage_last: Dict[str, float] = {}
        self._weights_version = -1                     # version of the nnet whose parameters the engine holds
        self._detached: "weakref.WeakSet[InstanceGraph]" = weakref.WeakSet()   # removed instances whose nodes still live in the engine

    @staticmethod
    def _check_eps(eps: float) -> None:
        if eps != 0.0:
            raise ValueError("eps > 0 (random pops, graph_search.py:58-61) is not supported by the B200 engine")

    # ---- engine management ----------------------------------------------------------------------------
    def _heur_mode(self) -> int:
        fn = self._fn
        if isinstance(fn, (HeurZerosVFn, HeurZerosQFn)):
            return _lib.DX_HEUR_ZERO
        if isinstance(fn, (DeviceHeurV, DeviceHeurQ)):
            prec = self._precision or fn.precision
            return _lib.DX_HEUR_NET_BF16 if prec == "bf16" else _lib.DX_HEUR_NET_FP32
        return _lib.DX_HEUR_HOST

    def _make_engine(self, new_insts: List[InstanceGraph]) -> None:
        a = self.domain.get_num_acts()
        n_inst = max(self._max_instances or 0, len(new_insts))
        pop_total = sum(i.batch_size for i in new_insts)
        if self._max_instances and self._max_instances > len(new_insts):
            pop_total += (self._max_instances - len(new_insts)) * self.batch_size_default
        max_nodes = self._max_nodes or max(1 << 18, min(pop_total * a * 256, 48_000_000))
        mode = self._heur_mode()
        self._engine = _lib.Engine(self.domain.dx_name, self.domain.dx_dim, self.dx_algo or ("graph_q" if self.is_q else "graph_v"), mode,
                                   max_instances=n_inst, max_pop_total=max(pop_total, n_inst), max_nodes=max_nodes,
                                   max_backups=self._max_backups)
        if mode in (_lib.DX_HEUR_NET_FP32, _lib.DX_HEUR_NET_BF16):
            load_nnet_into_engine(self._engine, self._fn.nnet)
            self._weights_version = getattr(self._fn.nnet, "version", 0)

    def _sync_weights(self) -> None:
        """Re-upload the parameters when the network changed since they were loaded (nnet.load_state_dict bumps
        `version`): the reference builds every PathFind around the CURRENT function (updaters/updater_rl.py)."""
        eng = self._engine
        if eng is None or eng.heur_mode not in (_lib.DX_HEUR_NET_FP32, _lib.DX_HEUR_NET_BF16):
            return
        v = getattr(self._fn.nnet, "version", 0)
        if v != self._weights_version:
            load_nnet_into_engine(eng, self._fn.nnet)
            self._weights_version = v

    def enable_backups(self, max_backups: int, max_instances: Optional[int] = None) -> None:
        """Extension (training-data generation, deepxube_b200/updaters/updater_rl.py): the engine evaluates every generated
        child before the CLOSED filter, as the reference does, and logs Node.bellman_backup() of every expanded node."""
        if self._engine is not None:
            raise RuntimeError("enable_backups must precede the first add_instances")
        if self.dx_algo not in (None, "graph_v", "graph_q"):
            raise TypeError("Bellman backups are implemented for graph_v / graph_q")
        self._max_backups = int(max_backups)
        if max_instances:
            self._max_instances = max(self._max_instances or 0, int(max_instances))

    def drain_backups(self, mode: str = "bellman"):
        """-> (state rows [n, S], [graph_q: action ids [n],] backups float64 [n], instance slots [n]) since the last drain;
        mode "bellman" | "ub_solns" | "lhbl" (graph_v)"""
        return self._engine.drain_backups(mode)

    def backup_deepest(self):
        """-> (state rows [n_slots, S], depths [n_slots]): per instance slot, the state of the deepest node that has children among
        the logged expansions (the root's when there is none) -- the relabelling source of UpdateHER._get_her_goals
        (base/updater.py:508-527).  Reads the log: call it before `drain_backups`."""
        return self._engine.backup_deepest()

    # ---- PathFind protocol -----------------------------------------------------------------------------
    def make_instances(self, states: List[Any], goals: List[Any], inst_infos: Optional[List[Any]] = None,
                       compute_root_vals: bool = True, batch_size: Optional[int] = None, weight: Optional[float] = None,
                       eps: Optional[float] = None, beam_size: Optional[int] = None) -> List[InstanceGraph]:
        """`beam_size` is the name the reference's A* variant gives its per-call batch override
        (graph_search.py:142-145); `batch_size` is the Q* variant's (:161-166).  Both are accepted."""
        t0 = time.time()
        if inst_infos is None:
            inst_infos = [None] * len(states)
        b = batch_size if batch_size is not None else (beam_size if beam_size is not None else self.batch_size_default)
        w = self.weight_default if weight is None else weight
        e = self.eps_default if eps is None else eps
        self._check_eps(e)
        insts = [InstanceGraph(Node(s, g, 0.0, 0.0), info, batch_size=b, weight=w, eps=e)
                 for s, g, info in zip(states, goals, inst_infos, strict=True)]
        self.times.record_time("root", time.time() - t0)
        return insts

    def add_instances(self, instances: List[InstanceGraph]) -> None:
        if not instances:
            return
        t0 = time.time()
        if self._engine is None:
            self._make_engine(instances)
        elif not self.instances and self._engine.num_instances() > 0 and self._max_backups == 0:
            # every earlier instance was removed: recycle the engine's instance slots, node store and CLOSED table
            # (the reference's add / step / remove_finished_instances / add loop, base/updater.py:343-400)
            self._recycle()
        self._sync_weights()
        eng, dom = self._engine, self.domain
        states = [i.root_node.state for i in instances]
        goals = [i.root_node.goal for i in instances]
        rows, grows = dom.states_to_rows(states), dom.goals_to_rows(goals)
        root_vals = None
        if eng.heur_mode == _lib.DX_HEUR_HOST:           # root values through the host callable (graph_search.py:103-107)
            root_vals = self._call_host_fn(states, goals)
        base = eng.num_instances()
        eng.add_instances(rows, grows, [i.batch_size for i in instances], [i.weight for i in instances], root_vals)
        for k, inst in enumerate(instances):
            inst._owner, inst._slot = self, base + k
        self.instances.extend(instances)
        self._refresh()
        self.times.record_time("heur", time.time() - t0)

    def _recycle(self) -> None:
        for inst in list(self._detached):              # instances the caller still holds: keep their solutions readable
            if inst.goal_node is not None and inst.goal_node.parent is None and inst.goal_node.path_cost > 0:
                inst._materialise_path(inst.goal_node)
            inst._owner, inst._slot = None, -1
        self._detached = weakref.WeakSet()
        self._engine.reset()

    def remove_instances(self, test_rem) -> List[Any]:
        """PathFind.remove_instances (base/pathfinding.py:300-310): removed instances stop searching on the device too;
        their nodes stay in the engine (get_path works) until the slots are recycled, which happens when instances
        are added while none is left."""
        removed = [i for i in self.instances if test_rem(i)]
        if removed:
            self.instances = [i for i in self.instances if not test_rem(i)]
            if self._engine is not None:
                self._engine.remove_instances([i._slot for i in removed if i._owner is self and i._slot >= 0])
            for i in removed:
                self._detached.add(i)
        return removed

    def _call_host_fn(self, states: List[Any], goals: List[Any]) -> np.ndarray:
        if self.is_q:
            actions_l = self.domain.get_state_actions(states)
            vals = self._fn(states, goals, actions_l, [None] * len(states))
            out = np.asarray(vals, dtype=np.float32).reshape(len(states), self.domain.get_num_acts())
        else:
            vals = self._fn(states, goals, [None] * len(states))
            out = np.asarray(vals, dtype=np.float32).reshape(len(states))
        assert out.shape[0] == len(states), f"{out.shape[0]}, {len(states)}"
        if np.isnan(out).any():
            raise ValueError("heuristic function returned NaN")
        return out

    def _refresh(self) -> None:
        sts = self._engine.status()
        for inst in self.instances:
            if inst._owner is self:
                inst._update(sts[inst._slot])

    def step(self, verbose: bool = False) -> Tuple[List[Node], List[Any]]:
        """One lock-step iteration over all unfinished instances (base/pathfinding.py:338-400 / :471-525).
        Popped nodes / edges are not materialised on the host (only the training code of the reference reads
        them); use Instance.get_nodes() to look at the nodes the next step will expand."""
        self.itr += 1
        eng = self._engine
        if eng is None or all(i.finished() for i in self.instances):
            return [], []
        t0 = time.time()
        if eng.heur_mode == _lib.DX_HEUR_HOST:
            n = eng.step_begin()
            rows, grows = eng.get_pending(n)
            dom = self.domain
            t1 = time.time()
            if n:
                states = dom.rows_to_states(rows)
                goals = self._rows_to_goals(grows)
                vals = self._call_host_fn(states, goals)
            else:
                vals = np.zeros((0, dom.get_num_acts()) if self.is_q else (0,), np.float32)
            t2 = time.time()
            self.times.record_time("heur", t2 - t1)
            eng.step_end(vals)
            self.times.record_time("search", (t1 - t0) + (time.time() - t2))
        else:
            if self.stage_times_on and not self._stage_last:
                eng.set_profiling(1)
                self._stage_last = eng.stage_ms()
            eng.run(1)
            self._record_stage_times(time.time() - t0)
        self._refresh()
        if verbose:
            fs = [i.frontier_size() for i in self.instances]
            print(f"Itr: {self.itr}, Frontier sizes(Min/Max): {min(fs)}/{max(fs)}, "
                  f"%has_soln: {100.0 * np.mean([i.has_soln() for i in self.instances])}, "
                  f"%finished: {100.0 * np.mean([i.finished() for i in self.instances])}")
            print(f"Times - {self.times.get_time_str()}\n")
        return [], []

    # reference Times key <- engine stages (csrc/engine.cu ST_*): "cost" is the f-key / sort stage of the pushed batch
    # (GraphSearch._compute_costs + the ordering work the heap does at push time), "pushpop" the OPEN merge + pop
    _STAGE_KEYS = (("expand", ("expand",)), ("filt", ("filt", "scatter")), ("heur", ("heur",)), ("cost", ("sort",)),
                   ("pushpop", ("pushpop",)), ("up_inst", ("book",)))

    def _record_stage_times(self, wall: float) -> None:
        if not self.stage_times_on:
            self.times.record_time("search", wall)
            return
        now = self._engine.stage_ms()
        dev = 0.0
        for key, stages in self._STAGE_KEYS:
            dt = sum(now[s] - self._stage_last.get(s, 0.0) for s in stages) / 1e3
            self.times.record_time(key, dt)
            dev += dt
        self._stage_last = now
        self.times.record_time("host", max(wall - dev, 0.0))      # launch, status read-back, Python

    def run(self, max_itrs: int) -> int:
        """Extension: up to max_itrs steps without returning to Python between them (device heuristics only)."""
        eng = self._engine
        if eng is None:
            return 0
        if eng.heur_mode == _lib.DX_HEUR_HOST:
            done = 0
            while done < max_itrs and not all(i.finished() for i in self.instances):
                self.step()
                done += 1
            return done
        t0 = time.time()
        if self.stage_times_on and not self._stage_last:
            eng.set_profiling(1)
            self._stage_last = eng.stage_ms()
        done = eng.run(max_itrs)
        self.itr += done
        self._record_stage_times(time.time() - t0)
        self._refresh()
        return done

    def _rows_to_goals(self, grows: np.ndarray) -> List[Any]:
        # goal rows have the state layout; wrap them in the goal type of the first instance
        goal_t = type(self.instances[0].root_node.goal)
        dom = self.domain
        if hasattr(goal_t, "__init__") and dom.dx_name == "grid":
            return [goal_t(int(r[0]), int(r[1])) for r in grows]
        return [goal_t(r) for r in grows]

    def stage_times(self) -> Dict[str, float]:
        return self._engine.stage_ms() if self._engine is not None else {}

    def reset(self) -> None:
        """Extension: drop all instances but keep the engine (buffers, loaded weights, captured graphs) --
        what `solve` gets in the reference by constructing a fresh, cheap PathFind per instance."""
        if self._engine is not None:
            self._recycle()
        self.instances = []
        self.itr = 0

    def close(self) -> None:
        if self._engine is not None:
            self._engine.close()
            self._engine = None

    def __repr__(self) -> str:
        return f"{type(self).__name__}(batch_size={self.batch_size_default}, weight={self.weight_default}, eps={self.eps_default})"


@pathfinding_factory.register_class("graph_v")
class GraphSearchHeurNodeActsEnum(GraphSearch):
    is_q = False

    @staticmethod
    def domain_type() -> Type:
        return ActsEnum

    @staticmethod
    def pathfind_functions_type() -> Type:
        return PFNsHeurV

    @staticmethod
    def description() -> str:
        return "Graph search with heuristic that prioritizes nodes"


@pathfinding_factory.register_class("graph_q")
class GraphSearchHeurEdgeActsEnum(GraphSearch):
    is_q = True

    @staticmethod
    def domain_type() -> Type:
        return ActsEnum

    @staticmethod
    def pathfind_functions_type() -> Type:
        return PFNsHeurQ

    @staticmethod
    def description() -> str:
        return "Graph search with heuristic that prioritizes edges"


class GraphSearchParser(Parser):
    _alg = "graph"

    def parse(self, args_str: str) -> Dict[str, Any]:
        kwargs: Dict[str, Any] = {}
        for tok in args_str.split("_"):
            mb, mw, me = re.search(r"^(\S+)B$", tok), re.search(r"^(\S+)W", tok), re.search(r"^(\S+)E", tok)
            if mb is not None:
                kwargs["batch_size"] = int(mb.group(1))
            elif mw is not None:
                kwargs["weight"] = float(mw.group(1))
            elif me is not None:
                kwargs["eps"] = float(me.group(1))
            else:
                raise ValueError(f"Unexpected argument {tok!r}")
        return kwargs

    def help(self) -> str:
        return ("<int>B (batch size), <float>W (weight), <float>E (epsilon for chance to randomly pop node; must be 0 here).\n"
                f"E.g. {self._alg}.10B_0.5W")


@pathfinding_factory.register_parser("graph_v")
class GraphSearchNodeParser(GraphSearchParser):
    _alg = "graph_v"


@pathfinding_factory.register_parser("graph_q")
class GraphSearchEdgeParser(GraphSearchParser):
    _alg = "graph_q"
