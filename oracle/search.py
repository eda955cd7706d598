"""Oracle restatement of batch weighted A* (``graph_v``) and batch weighted Q* (``graph_q``).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Pure-Python containers (heapq / dict) around
numpy arrays, exactly the data structures the reference uses, so the semantics -- not just the
results -- are the reference's:

  InstanceGraph (OPEN heap of (f, push_count, item), CLOSED dict state -> g, lb/ub, finished)
        deepxube/pathfinding/graph_search.py:20-80
  InstanceNodeGraph seeds CLOSED[root] = 0.0 (A* only)       graph_search.py:119-121
  PathFindNode.step / _expand  (A* iteration driver)          deepxube/base/pathfinding.py:338-463
  PathFindEdge.step / _get_edge_nodes / _get_edges (Q*)       deepxube/base/pathfinding.py:471-588
  PathFindNodeStatic._get_next_nodes / PathFindEdgeStatic._get_next_edges   pathfinding.py:621-671
  GraphSearchHeurNode._compute_costs  f = w*g + h (float64)   graph_search.py:147-158
  GraphSearchHeurEdge._compute_costs  f = w*g(parent) + q     graph_search.py:168-179
  PathFindSetHeurV / PathFindSetHeurQ (h = min_a q)           pathfinding.py:701-744
  get_path (parent walk)                                      pathfinding.py:93-120

Nodes are rows of growing arrays instead of ``Node`` objects; node ids are assigned in creation
order per instance (root = 0).  eps > 0 (random pops, graph_search.py:58-61) is not supported.
"""
from heapq import heappop, heappush
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np

from .domains_np import OracleDomain

# heuristic callables take raw rows: (states [N,S] u8, goals [N,G] u8) -> float64 [N] (V) or [N,A] (Q)
HeurFn = Callable[[np.ndarray, np.ndarray], np.ndarray]


class _Store:
    """Growing per-instance node arrays."""

    def __init__(self, state_len: int, num_actions: int, with_q: bool):
        self.n = 0
        self.cap = 1024
        self.state = np.zeros((self.cap, state_len), np.uint8)
        self.g = np.zeros(self.cap, np.float64)
        self.h = np.zeros(self.cap, np.float64)
        self.parent = np.full(self.cap, -1, np.int64)
        self.action = np.full(self.cap, -1, np.int64)
        self.q = np.zeros((self.cap, num_actions), np.float64) if with_q else None

    def _grow(self, need: int) -> None:
        if self.n + need <= self.cap:
            return
        cap = max(self.cap * 2, self.n + need)
        for name in ("state", "g", "h", "parent", "action", "q"):
            arr = getattr(self, name)
            if arr is None:
                continue
            new = np.zeros((cap,) + arr.shape[1:], arr.dtype)
            new[: self.n] = arr[: self.n]
            setattr(self, name, new)
        self.cap = cap

    def append(self, states: np.ndarray, g: np.ndarray, parent: np.ndarray, action: np.ndarray) -> np.ndarray:
        k = states.shape[0]
        self._grow(k)
        ids = np.arange(self.n, self.n + k)
        self.state[ids] = states
        self.g[ids] = g
        self.parent[ids] = parent
        self.action[ids] = action
        self.n += k
        return ids


class OracleInstance:
    def __init__(self, domain: OracleDomain, state: np.ndarray, goal: np.ndarray, batch_size: int, weight: float,
                 is_q: bool):
        self.domain = domain
        self.goal = np.asarray(goal, np.uint8)
        self.batch_size = int(batch_size)
        self.weight = float(weight)
        self.store = _Store(domain.state_len, domain.num_actions, is_q)
        self.store.append(np.asarray(state, np.uint8)[None], np.zeros(1), np.array([-1]), np.array([-1]))
        self.open: List[Tuple[float, int, int]] = []          # (f, push_count, item)
        self.push_count = 0
        self.closed: Dict[bytes, float] = {}
        self.ub = np.inf
        self.lb = 0.0            # root heuristic is still 0.0 when the instance is built (graph_search.py:27)
        self.goal_node: Optional[int] = None
        self.nodes_curr = np.array([0], np.int64)
        self.itr = 0
        self.num_nodes_generated = 0
        self.expanded: set = set()      # ids of the nodes that have children (A*: expanded; Q*: at least one edge popped)
        if not is_q:
            self.closed[self.store.state[0].tobytes()] = 0.0   # graph_search.py:119-121

    def finished(self) -> bool:                               # graph_search.py:43-46
        case1 = (self.goal_node is not None) and (self.lb >= (self.weight * self.ub))
        case2 = (self.itr > 0) and (len(self.nodes_curr) == 0)
        return bool(case1 or case2)

    def record_goal(self, ids: np.ndarray, solved: np.ndarray) -> None:   # graph_search.py:35-41
        for nid, s in zip(ids.tolist(), solved.tolist()):
            if s and (self.ub > self.store.g[nid]):
                self.goal_node = nid
                self.ub = float(self.store.g[nid])

    def check_closed(self, states: np.ndarray, g: np.ndarray) -> np.ndarray:   # graph_search.py:73-80
        keep = np.zeros(states.shape[0], bool)
        closed = self.closed
        for i in range(states.shape[0]):
            key = states[i].tobytes()
            prev = closed.get(key)
            gi = float(g[i])
            if (prev is None) or (prev > gi):
                closed[key] = gi
                keep[i] = True
        return keep

    def push(self, items: np.ndarray, costs: np.ndarray) -> None:          # graph_search.py:48-51
        for it, c in zip(items.tolist(), costs.tolist()):
            heappush(self.open, (c, self.push_count, it))
            self.push_count += 1

    def pop(self) -> np.ndarray:                                           # graph_search.py:53-71
        k = min(self.batch_size, len(self.open))
        popped = [heappop(self.open) for _ in range(k)]
        if popped:
            self.lb = max(popped[0][0], self.lb)
        return np.array([p[2] for p in popped], np.int64)

    def path_cost(self) -> float:
        return np.inf if self.goal_node is None else float(self.store.g[self.goal_node])

    def has_soln(self) -> bool:
        return self.goal_node is not None

    def deepest_with_children(self) -> Tuple[np.ndarray, int]:            # base/updater.py:508-527
        """UpdateHER._get_her_goals' relabelling source: walk the root's descendants breadth first (Node.get_all_descendants,
        base/pathfinding.py:78-90: a FIFO, children in edge_dict insertion order = node-id order here), keep the FIRST node with
        children whose depth (the reference counts the states on its path, here the hops: same order) is strictly larger than any before;
        the root's state when no descendant has children.  "Has children" = was expanded (`self.expanded`), also when the CLOSED
        filter dropped every child; dropped children never enter the store, but they have no children of their own, so leaving
        them out changes neither the candidates nor their order.  -> (state row, depth)"""
        st = self.store
        kids: Dict[int, List[int]] = {}
        for nid in range(1, st.n):
            kids.setdefault(int(st.parent[nid]), []).append(nid)
        best, best_depth = 0, 0
        fifo = [(c, 1) for c in kids.get(0, [])]
        while fifo:
            nid, depth = fifo.pop(0)
            for c in kids.get(nid, []):
                fifo.append((c, depth + 1))
            if nid in self.expanded and depth > best_depth:
                best, best_depth = nid, depth
        return st.state[best].copy(), best_depth

    def get_path(self) -> Tuple[np.ndarray, List[int], float]:            # base/pathfinding.py:93-120
        assert self.goal_node is not None
        ids, acts = [], []
        nid = self.goal_node
        while self.store.parent[nid] >= 0:
            ids.append(nid)
            acts.append(int(self.store.action[nid]))
            nid = int(self.store.parent[nid])
        ids.append(nid)
        ids.reverse()
        acts.reverse()
        return self.store.state[ids].copy(), acts, float(self.store.g[self.goal_node])


class OracleSearch:
    """Lock-step search over many instances (ref: PathFind holds ``instances`` and one heuristic call
    per step covers all of them, base/pathfinding.py:339-349,402-426,702-718)."""

    def __init__(self, domain: OracleDomain, heur_fn: Optional[HeurFn], is_q: bool, batch_size: int = 1,
                 weight: float = 1.0, eval_kept_only: bool = False, trace: bool = False):
        self.domain = domain
        self.heur_fn = heur_fn
        self.is_q = is_q
        self.batch_size = batch_size
        self.weight = weight
        self.eval_kept_only = eval_kept_only
        self.instances: List[OracleInstance] = []
        self.itr = 0
        self.trace_on = trace
        self.trace: List[List[dict]] = []
        self.num_heur_evals = 0
        # Bellman backups of the popped nodes, per step: (states, backup values float64, instance index) -- what
        # UpdateHeurVRLKeepGoalABC._get_labels_no_rb (updaters/updater_rl.py:143-158) reads after each PathFind.step()
        self.backup_log: List[Tuple[np.ndarray, np.ndarray, np.ndarray]] = []
        self.expansions: List[dict] = []      # one entry per (step, instance), in backup_log order
        self.solved_mark: Dict[Tuple[int, int], bool] = {}     # Q*: Node.is_solved of the nodes that were popped (else None)

    # zero heuristic defaults (ref: base/pathfind_fns.py:205-210, 235-243)
    def _heur(self, states: np.ndarray, goals: np.ndarray) -> np.ndarray:
        n, a = states.shape[0], self.domain.num_actions
        self.num_heur_evals += n
        if self.heur_fn is None:
            return np.zeros((n, a)) if self.is_q else np.zeros(n)
        if n == 0:
            return np.zeros((0, a)) if self.is_q else np.zeros(0)
        out = np.asarray(self.heur_fn(states, goals), np.float64)
        assert out.shape == ((n, a) if self.is_q else (n,)), out.shape
        return out

    def add_instances(self, states: np.ndarray, goals: np.ndarray, batch_size: Optional[int] = None,
                      weight: Optional[float] = None) -> List[OracleInstance]:
        """make_instances + add_instances (graph_search.py:94-109,142-145,161-166): the root value is
        computed with one heuristic call over the new roots."""
        b = self.batch_size if batch_size is None else batch_size
        w = self.weight if weight is None else weight
        new = [OracleInstance(self.domain, s, g, b, w, self.is_q) for s, g in zip(states, goals)]
        if new:
            vals = self._heur(np.stack([i.store.state[0] for i in new]), np.stack([i.goal for i in new]))
            for inst, v in zip(new, vals):
                if self.is_q:
                    inst.store.q[0] = v
                    inst.store.h[0] = v.min()
                else:
                    inst.store.h[0] = v
        self.instances.extend(new)
        return new

    def all_finished(self) -> bool:
        return all(i.finished() for i in self.instances)

    def step(self) -> None:
        insts = [i for i in self.instances if not i.finished()]
        self.itr += 1
        if not insts:
            return
        if self.is_q:
            self._step_q(insts)
        else:
            self._step_v(insts)
        for i in insts:
            i.itr += 1

    # ---- batch weighted A* -----------------------------------------------------------------------
    def _step_v(self, insts: List[OracleInstance]) -> None:
        dom, a = self.domain, self.domain.num_actions
        per = []
        for inst in insts:
            ids = inst.nodes_curr
            st = inst.store
            pst = st.state[ids]
            goals = np.broadcast_to(inst.goal, (len(ids), dom.goal_len))
            inst.record_goal(ids, dom.is_solved(pst, goals))
            cst = dom.expand(pst)
            cg = np.repeat(st.g[ids], a) + 1.0
            cpar = np.repeat(ids, a)
            cact = np.tile(np.arange(a), len(ids))
            inst.num_nodes_generated += cst.shape[0]
            inst.expanded.update(int(x) for x in ids)
            per.append([inst, cst, cg, cpar, cact, None])
        if not self.eval_kept_only:
            allst = np.concatenate([p[1] for p in per])
            allgoal = np.concatenate([np.broadcast_to(p[0].goal, (p[1].shape[0], dom.goal_len)) for p in per])
            h_all = self._heur(allst, allgoal)
            off = 0
            for p in per:
                p[5] = h_all[off: off + p[1].shape[0]]
                off += p[1].shape[0]
            # Node.bellman_backup (base/pathfinding.py:41-47): 0 if solved else min_a (tc + heuristic(child_a)), tc = 1.0
            bst, bval, bidx = [], [], []
            for p in per:
                inst = p[0]
                ids = inst.nodes_curr
                pst = inst.store.state[ids]
                solved = dom.is_solved(pst, np.broadcast_to(inst.goal, (len(ids), dom.goal_len)))
                ch = np.asarray(p[5], np.float64).reshape(len(ids), a)
                mins = (1.0 + ch).min(axis=1) if len(ids) else np.zeros(0)
                bst.append(pst.copy())
                bval.append(np.where(solved, 0.0, mins))
                bidx.append(np.full(len(ids), self.instances.index(inst), np.int32))
                # what the other label options need (labels_v): the expanded nodes, their children's values / node ids
                p.append(dict(inst=self.instances.index(inst), node=np.asarray(ids).copy(), solved=solved.copy(), child_h=ch,
                              child_id=np.full((len(ids), a), -1, np.int64)))
            self.backup_log.append((np.concatenate(bst), np.concatenate(bval), np.concatenate(bidx)))
        keeps = []
        for p in per:
            inst, cst, cg = p[0], p[1], p[2]
            keeps.append(inst.check_closed(cst, cg))
        if self.eval_kept_only:
            kst = np.concatenate([p[1][k] for p, k in zip(per, keeps)])
            kgoal = np.concatenate([np.broadcast_to(p[0].goal, (int(k.sum()), dom.goal_len)) for p, k in zip(per, keeps)])
            h_k = self._heur(kst, kgoal)
            off = 0
            for p, k in zip(per, keeps):
                hh = np.zeros(p[1].shape[0])
                hh[k] = h_k[off: off + int(k.sum())]
                off += int(k.sum())
                p[5] = hh
        rec = []
        for p, keep in zip(per, keeps):
            inst, cst, cg, cpar, cact, h = p[:6]
            popped_prev = inst.nodes_curr
            new_ids = inst.store.append(cst[keep], cg[keep], cpar[keep], cact[keep])
            if len(p) > 6:
                p[6]["child_id"].reshape(-1)[keep] = new_ids
                self.expansions.append(p[6])
            inst.store.h[new_ids] = h[keep]
            costs = np.array(inst.weight) * cg[keep] + h[keep]          # float64, graph_search.py:153
            inst.push(new_ids, costs)
            inst.nodes_curr = inst.pop()
            if self.trace_on:
                rec.append(dict(popped=popped_prev.copy(), keep=keep.copy(), f=costs.copy(), next=inst.nodes_curr.copy(),
                                lb=inst.lb, ub=inst.ub, open=len(inst.open), closed=len(inst.closed)))
        if self.trace_on:
            self.trace.append(rec)

    def labels_v(self, mode: str) -> np.ndarray:
        """Labels of all expanded nodes so far, in backup_log order, for the updater's options (UpRLArgs, base/updater.py:669-672;
        UpdateHeurVRLKeepGoalABC._get_labels_no_rb, updaters/updater_rl.py:143-158):
          "bellman"   Node.bellman_backup() of every popped node
          "ub_solns"  + every solved popped node bounds its ancestors' values by the path cost down to it
                      (Node.upper_bound_parent_path, base/pathfinding.py:49-53)
          "lhbl"      Node.tree_backup() from the root (base/pathfinding.py:55-64)"""
        recs = []                                  # (inst, node, solved, child_h, child_id), flat, expansion order
        for e in self.expansions:
            for k in range(len(e["node"])):
                recs.append((e["inst"], int(e["node"][k]), bool(e["solved"][k]), e["child_h"][k], e["child_id"][k]))
        rec_of = {(r[0], r[1]): i for i, r in enumerate(recs)}
        vals = np.array([0.0 if r[2] else float((1.0 + r[3]).min()) for r in recs])
        if mode == "bellman":
            return vals
        if mode == "ub_solns":
            for i, r in enumerate(recs):
                if not r[2]:
                    continue
                store = self.instances[r[0]].store
                node, dist = r[1], 0.0
                while store.parent[node] >= 0:
                    node = int(store.parent[node])
                    dist += 1.0
                    j = rec_of.get((r[0], node))
                    if j is not None:
                        vals[j] = min(vals[j], dist)
            return vals
        assert mode == "lhbl"
        for i in range(len(recs) - 1, -1, -1):     # children are expanded after their parents
            inst, _, solved, ch, cid = recs[i]
            if solved:
                vals[i] = 0.0
                continue
            best = np.inf
            for a_i in range(len(ch)):
                j = rec_of.get((inst, int(cid[a_i]))) if cid[a_i] >= 0 else None
                best = min(best, 1.0 + (vals[j] if j is not None else max(float(ch[a_i]), 0.0)))
            vals[i] = best
        return vals

    def labels_q(self, mode: str) -> np.ndarray:
        """Q* labels of all popped edges so far, in backup_log order (UpdateHeurQRLKeepGoalABC._get_labels_no_rb,
        updaters/updater_rl.py:169-189): Node.backup_act(action) after the optional ub_heur_solns / lhbl passes."""
        recs = []                                  # (inst, parent, child, h(child)), flat, pop order
        for e in self.expansions:
            for k in range(len(e["child"])):
                recs.append((e["inst"], int(e["parent"][k]), int(e["child"][k]), float(e["child_h"][k])))
        sol = lambda i, n: self.solved_mark.get((i, n), False)
        if mode == "bellman":
            return np.array([0.0 if sol(i, p) else 1.0 + h for i, p, c, h in recs])
        if mode == "ub_solns":
            ub: Dict[Tuple[int, int], float] = {}
            for i, p, c, h in recs:
                if not sol(i, p):
                    continue
                store = self.instances[i].store
                node, dist = p, 0.0
                while True:
                    ub[(i, node)] = min(ub.get((i, node), np.inf), dist)
                    if store.parent[node] < 0:
                        break
                    node = int(store.parent[node])
                    dist += 1.0
            return np.array([0.0 if sol(i, p) else 1.0 + ub.get((i, c), h) for i, p, c, h in recs])
        assert mode == "lhbl"
        emin: Dict[Tuple[int, int], float] = {}
        vals = np.zeros(len(recs))
        for r in range(len(recs) - 1, -1, -1):
            i, p, c, h = recs[r]
            t = 0.0 if sol(i, c) else (emin[(i, c)] if (i, c) in emin else max(h, 0.0))
            vals[r] = 0.0 if sol(i, p) else 1.0 + t
            emin[(i, p)] = min(emin.get((i, p), np.inf), 1.0 + t)
        return vals

    # ---- batch weighted Q* -----------------------------------------------------------------------
    def _step_q(self, insts: List[OracleInstance]) -> None:
        dom, a = self.domain, self.domain.num_actions
        per = []
        for inst in insts:
            ids = inst.nodes_curr
            st = inst.store
            pst = st.state[ids]
            goals = np.broadcast_to(inst.goal, (len(ids), dom.goal_len))
            solved_now = dom.is_solved(pst, goals)
            inst.record_goal(ids, solved_now)
            for nid, sv in zip(ids, solved_now):                         # PathFind.set_is_solved (base/pathfinding.py:353)
                self.solved_mark[(self.instances.index(inst), int(nid))] = bool(sv)
            keep = inst.check_closed(pst, st.g[ids])                     # closed test at pop, graph_search.py:132-133
            kid = ids[keep]
            # edges: kept-node order x action id (base/pathfinding.py:568-588)
            e_item = (np.repeat(kid, a) * a + np.tile(np.arange(a), len(kid))).astype(np.int64)
            costs = np.array(inst.weight) * np.repeat(st.g[kid], a) + st.q[kid].reshape(-1)   # graph_search.py:175
            inst.push(e_item, costs)
            e_pop = inst.pop()
            par, act = e_pop // a, e_pop % a
            cst = dom.next_state(st.state[par], act) if len(par) else np.zeros((0, dom.state_len), np.uint8)
            cg = st.g[par] + 1.0                                          # base/pathfinding.py:542
            new_ids = st.append(cst, cg, par, act)
            inst.expanded.update(int(x) for x in par)
            inst.num_nodes_generated += len(new_ids)                      # base/pathfinding.py:562-563
            per.append((inst, new_ids, ids, keep, costs))
        allst = np.concatenate([p[0].store.state[p[1]] for p in per])
        allgoal = np.concatenate([np.broadcast_to(p[0].goal, (len(p[1]), dom.goal_len)) for p in per])
        q_all = self._heur(allst, allgoal)
        off = 0
        rec = []
        bq: List[Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]] = []
        for inst, new_ids, ids, keep, costs in per:
            q = q_all[off: off + len(new_ids)]
            off += len(new_ids)
            inst.store.q[new_ids] = q
            inst.store.h[new_ids] = q.min(axis=1) if len(new_ids) else 0.0
            # Node.backup_act (base/pathfinding.py:66-77) of the popped edges: 0 if the edge's node is solved else
            # tc + heuristic(node_next); node_next.backup_val is +inf at this point (updaters/updater_rl.py:181-186)
            par, act = inst.store.parent[new_ids], inst.store.action[new_ids]
            solved = dom.is_solved(inst.store.state[par], np.broadcast_to(inst.goal, (len(par), dom.goal_len)))
            hn = q.min(axis=1).astype(np.float64) if len(new_ids) else np.zeros(0)
            bq.append((inst.store.state[par].copy(), np.where(solved, 0.0, 1.0 + hn), np.full(len(par), self.instances.index(inst), np.int32),
                       np.asarray(act, np.int32)))
            self.expansions.append(dict(inst=self.instances.index(inst), parent=np.asarray(par).copy(), child=np.asarray(new_ids).copy(),
                                        child_h=hn.copy()))
            inst.nodes_curr = new_ids
            if self.trace_on:
                rec.append(dict(popped=ids.copy(), keep=keep.copy(), f=costs.copy(), next=new_ids.copy(), lb=inst.lb,
                                ub=inst.ub, open=len(inst.open), closed=len(inst.closed)))
        if self.trace_on:
            self.trace.append(rec)
        self.backup_log.append(tuple(np.concatenate([b[k] for b in bq]) for k in range(4)))


def solve_one(domain: OracleDomain, state: np.ndarray, goal: np.ndarray, heur_fn: Optional[HeurFn], is_q: bool,
              batch_size: int, weight: float, max_itrs: Optional[int] = None, eval_kept_only: bool = False) -> dict:
    """The ``solve`` loop for one instance (ref: deepxube/_solve.py:141-191)."""
    srch = OracleSearch(domain, heur_fn, is_q, batch_size, weight, eval_kept_only)
    inst = srch.add_instances(state[None], goal[None])[0]
    itrs = 0
    while not inst.finished():
        srch.step()
        itrs += 1
        if (max_itrs is not None) and (itrs == max_itrs):
            break
    out = dict(iterations=itrs, num_nodes_generated=inst.num_nodes_generated, path_cost=inst.path_cost(),
               solved=inst.goal_node is not None, actions=None, heur_evals=srch.num_heur_evals)
    if inst.goal_node is not None:
        _, acts, _ = inst.get_path()
        out["actions"] = acts
    return out


# ---- replay-buffer labels (test infrastructure, like everything in oracle/) -------------------------------------------
def vi_targets_v(domain: OracleDomain, heur_fn: HeurFn, states: np.ndarray, goals: np.ndarray,
                 solved: Optional[np.ndarray] = None) -> np.ndarray:
    """UpdateHeurVRL._get_labels_rb (updaters/updater_rl.py:41-69): expand the sampled states, evaluate the target network on
    every child, ctg = min_a (tc + h(child_a)) * not(is_solved); tc = 1.0, numpy float64."""
    n, a = states.shape[0], domain.num_actions
    if solved is None:
        solved = domain.is_solved(states, goals)
    if n == 0:
        return np.zeros(0)
    children = domain.expand(states)
    h = np.asarray(heur_fn(children, np.repeat(goals, a, axis=0)), np.float64)
    return (1.0 + h).reshape(n, a).min(axis=1) * np.logical_not(np.asarray(solved, bool))


def vi_targets_q(domain: OracleDomain, heurq_fn: HeurFn, states_next: np.ndarray, goals: np.ndarray,
                 solved: Optional[np.ndarray] = None) -> np.ndarray:
    """UpdateHeurQRL._get_labels_rb (updaters/updater_rl.py:102-120): ctg = (tc + min_a q_targ(next, a)) * not(is_solved of
    the edge's node); tc = 1.0."""
    n = states_next.shape[0]
    if n == 0:
        return np.zeros(0)
    q = np.asarray(heurq_fn(states_next, goals), np.float64)
    sol = np.zeros(n, bool) if solved is None else np.asarray(solved, bool)
    return (1.0 + q.min(axis=1)) * np.logical_not(sol)
