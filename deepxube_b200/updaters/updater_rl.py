"""Training-data generation from the search: the value-iteration targets of the heuristic network.

Mirror of ONE runner of the reference's RL updaters for `heurv` / `heurq`, with the goal kept or relabelled in hindsight
(HER, `UpdateHeurVRLHER` / `UpdateHeurQRLHER` below); labels without a
replay buffer: Bellman backups, `ub_heur_solns`, LHBL tree backups; with a replay buffer (`rb_size > 0`): the popped nodes /
edges of every emitted instance group go into an array-backed replay buffer, as many elements are sampled back, and their
value-iteration targets are recomputed with the (target) network on the device (`dx_vi_targets`):

    base/updater.py:343-400        update_runner: for each of `search_itrs` iterations -- add instances (new ones for the
                                   instances removed last iteration, same steps_gen), one PathFind.step(), remove the
                                   instances that finished or reached search_itrs, emit their data; finally emit the rest
    base/updater.py:781-791        data of an instance group = the nodes those instances popped + their labels
    updaters/updater_rl.py:143-158 V labels = Node.bellman_backup() of each popped node
    updaters/updater_rl.py:169-189 Q labels = Node.backup_act(action) of each popped edge
    base/updater.py:904-921, 938-965 -> [*nnet inputs of (state, goal[, action]), labels]
    base/updater.py:763-795        _rb_add_sample_train_data / _get_instance_data_rb: add, sample len(popped), label, to_np
    updaters/updater_rl.py:30-69, 85-120  replay data (is_solved [, tc, next state]) and the labels of sampled elements
    base/updater.py:330-341        one `update_runner` work item = one batch of searches; `generate()` runs several
    base/updater.py:485-583, 803-818  HER: instances that did not finish with a solution get the goal sampled from their deepest
                                   expanded node (found on the device, `dx_backup_deepest`), goal-keeping instances first

The search and the backups run on the device (graph_v engine created with max_backups > 0: every generated child is
evaluated before the CLOSED filter and each expanded node's backup is logged by k_bellman_backup); this module only owns the
instance bookkeeping.  The reference spreads runners over processes and feeds a training loop through queues; neither is part
of the hot path.
"""
from typing import Any, List, Optional, Tuple

import numpy as np

from deepxube_b200.pathfinding.graph_search import GraphSearch


_GOAL_CACHE: dict = {}       # goal row bytes -> goal object (replay-buffer mode hands goal OBJECTS back, like the reference)


class UpdateHeurVRLKeepGoal:
    is_q = False
    her = False

    def __init__(self, domain: Any, pathfind: GraphSearch, search_itrs: int, nnet_input: Optional[Any] = None,
                 ub_heur_solns: bool = False, lhbl: bool = False, rb_size: int = 0):
        """ub_heur_solns / lhbl: UpRLArgs (base/updater.py:669-672); lhbl wins when both are set, as in
        updaters/updater_rl.py:145-155, 171-179.  Both are computed on the device when the log is drained."""
        self.label_mode = "lhbl" if lhbl else ("ub_solns" if ub_heur_solns else "bellman")
        want = "graph_q" if self.is_q else "graph_v"
        if pathfind.is_q != self.is_q or pathfind.dx_algo not in (None, want):
            raise TypeError(f"{type(self).__name__} needs a {want} search (PathFindSetHeur{'Q' if self.is_q else 'V'}, "
                            "updaters/updater_rl.py:137-165)")
        if pathfind.instances:
            raise ValueError("the PathFind must be fresh (base/updater.py:364 builds one per batch)")
        self.domain, self.pathfind, self.search_itrs, self.nnet_input = domain, pathfind, int(search_itrs), nnet_input
        self.rb_size = int(rb_size)
        self.rb: Optional[Any] = None                   # _init_replay_buffer (base/updater.py:707-708), built on first use
        if self.her:
            if self.rb_size <= 0:
                raise NotImplementedError("Must use replay buffer if doing HER.")                     # base/updater.py:483
            if not hasattr(domain, "sample_goal_from_state"):
                raise TypeError("HER needs a GoalSampleableFromState domain (base/domain.py:421-431); of the device domains: grid")

    def _make_instances(self, steps_gen: List[int]) -> List[Any]:
        states, goals = self.domain.sample_problem_instances(steps_gen)                      # base/updater.py:679-685
        return self.pathfind.make_instances(states, goals, inst_infos=[(s,) for s in steps_gen], compute_root_vals=False)

    def get_instance_data(self, steps_gen: List[int], instances: Optional[List[Any]] = None) -> Tuple:
        """-> (states, goals, [Q: actions,] ctgs_backup) in the order the reference emits them: groups of removed instances
        in removal order, instances in order within a group, each instance's popped nodes / edges in pop order.
        `instances`: use these instead of sampling the first batch (tests)."""
        pf, n = self.pathfind, len(steps_gen)
        a = self.domain.get_num_acts()
        b = max(pf.batch_size_default, 1)
        # log capacity: one record per popped node; the iteration after an instance finishes its last pops still take (dropped) slots
        need = 2 * n * b * self.search_itrs + n * b
        if pf._engine is None:
            pf.enable_backups(max_backups=need, max_instances=n * self.search_itrs)
        else:                               # a later batch on the same engine (the reference builds a new PathFind per batch,
            if pf._max_backups < need:      # base/updater.py:364; here the buffers, weights and captured graphs are kept)
                raise ValueError(f"batch of {n} instances needs a backup log of {need} records, the engine has {pf._max_backups}")
            pf.reset()
        removed: List[Any] = []
        groups: List[List[Any]] = []
        log: List[Tuple[np.ndarray, np.ndarray, np.ndarray]] = []
        for it in range(self.search_itrs):
            if it == 0:
                new = instances if instances is not None else self._make_instances(list(steps_gen))
            else:
                new = self._make_instances([int(i.inst_info[0]) for i in removed]) if removed else []
            pf.add_instances(new)
            assert len(pf.instances) == n, f"Values were {len(pf.instances)} and {n}"
            pf.step()
            removed = pf.remove_finished_instances(self.search_itrs)
            if removed:
                groups.append(removed)
        if pf.instances:
            groups.append(list(pf.instances))
        deepest = pf.backup_deepest()[0] if self.her else None     # (before the drain empties the log)
        drained = pf.drain_backups(self.label_mode)
        rows, vals, slots = drained[0], drained[-2], drained[-1]
        if self.rb_size > 0:
            return self._emit_rb(groups, rows, slots, drained[1] if self.is_q else None, deepest)
        # regroup the device log (iteration-major) into the reference's emission order
        per_group = [np.concatenate([np.flatnonzero(slots == inst._slot) for inst in g]) if len(vals) else np.zeros(0, np.int64)
                     for g in groups]
        goal_of = {inst._slot: inst.root_node.goal for g in groups for inst in g}
        order = np.concatenate(per_group) if per_group else np.zeros(0, np.int64)
        states = self.domain.rows_to_states(rows[order])
        goals = [goal_of[int(s)] for s in slots[order]]
        if self.is_q:
            fixed = self.domain.get_actions_fixed()
            return states, goals, [fixed[int(k)] for k in drained[1][order]], vals[order]
        return states, goals, vals[order]

    # ---- replay-buffer mode (base/updater.py:763-795) -----------------------------------------------------------------
    def _init_replay_buffer(self, state_bytes: int, goal_bytes: int) -> None:
        from deepxube_b200.updaters.replay_buffer import ReplayBufferQ, ReplayBufferV
        self.rb = (ReplayBufferQ if self.is_q else ReplayBufferV)(self.rb_size, state_bytes, goal_bytes)

    def _group_goals(self, group: List[Any], deepest: Optional[np.ndarray]) -> Tuple[List[Any], List[Any]]:
        """-> the group's instances in emission order and the goal each one's data is labelled against: their own."""
        return list(group), [inst.root_node.goal for inst in group]

    def _emit_rb(self, groups: List[List[Any]], rows: np.ndarray, slots: np.ndarray, acts: Optional[np.ndarray],
                 deepest: Optional[np.ndarray]) -> Tuple:
        """For every emitted instance group, in emission order: popped data -> replay buffer, sample as many elements back,
        label them with the device network (`_get_instance_data_rb` -> `_rb_add_sample_train_data`).  Everything stays in packed
        rows until the end; -> (states, goals, [Q: actions,] labels) of the SAMPLED elements, groups concatenated."""
        from deepxube_b200 import _lib
        dom = self.domain
        fn_eng = self._device_fn()._engine()
        out_rows, out_grows, out_acts, out_labels = [], [], [], []
        goal_of: dict = {}
        grow_of: dict = {}
        for group in groups:
            insts, goals_g = self._group_goals(group, deepest)
            grows_g = dom.goals_to_rows(goals_g)
            for inst, goal, grow in zip(insts, goals_g, grows_g):
                goal_of[inst._slot], grow_of[inst._slot] = goal, grow
            idx = np.concatenate([np.flatnonzero(slots == inst._slot) for inst in insts]) if len(slots) else np.zeros(0, np.int64)
            if len(idx) == 0:
                continue
            r = rows[idx]
            g = np.stack([grow_of[int(k)] for k in slots[idx]])
            if self.rb is None:
                self._init_replay_buffer(r.shape[1], g.shape[1])
            solved = _lib.is_solved(dom.dx_name, dom.dx_dim, r, g)                       # Node.is_solved of the popped nodes
            if self.is_q:
                a = acts[idx]
                nxt = _lib.next_state(dom.dx_name, dom.dx_dim, r, a)                     # node.edge_dict[action] = (tc, node_next)
                self.rb.add(r, g, a, solved, np.ones(len(idx)), nxt)
                (sr, sg, sa), (ssol, stc, snxt) = self.rb.sample(len(idx))
                assert np.all(stc == 1.0), "the in-scope domains have unit transition costs"
                labels = fn_eng.vi_targets(snxt, sg, ssol)                               # (tc + min_a q_targ(next, a)) * not solved
                out_acts.append(sa)
            else:
                self.rb.add(r, g, solved)
                (sr, sg), ssol = self.rb.sample(len(idx))
                labels = fn_eng.vi_targets(sr, sg, ssol)                                 # min_a (tc + h_targ(child_a)) * not solved
            out_rows.append(sr); out_grows.append(sg); out_labels.append(labels)
        if not out_rows:
            return ([], [], [], np.zeros(0)) if self.is_q else ([], [], np.zeros(0))
        srows, sgrows = np.concatenate(out_rows), np.concatenate(out_grows)
        states = dom.rows_to_states(srows)
        goals = dom.rows_to_goals(sgrows) if hasattr(dom, "rows_to_goals") else self._goals_from_rows(sgrows, goal_of, grow_of)
        labels = np.concatenate(out_labels)
        if self.is_q:
            fixed = dom.get_actions_fixed()
            return states, goals, [fixed[int(k)] for k in np.concatenate(out_acts)], labels
        return states, goals, labels

    @staticmethod
    def _goals_from_rows(grows: np.ndarray, goal_of: dict, grow_of: dict) -> List[Any]:
        """goal objects of sampled goal rows: rows that came from this batch map back to their objects; rows from earlier batches
        (still in the buffer) are resolved through the cache the updater keeps of every goal row it has seen"""
        cache = _GOAL_CACHE
        for k, row in grow_of.items():
            cache.setdefault(row.tobytes(), goal_of[k])
        return [cache[row.tobytes()] for row in grows]

    def generate(self, num_batches: int, batch_size: int, step_probs: List[float], step_max: int) -> List[Tuple]:
        """`update_runner`'s inner loop over work items (base/updater.py:330-395): `num_batches` batches of `batch_size` searches,
        start-state step counts drawn from `step_probs` over 0..step_max (`_add_instances`, :418-423).  The engine, its weights
        and captured graphs are reused across batches; with `rb_size > 0` so is the replay buffer."""
        out = []
        for _ in range(num_batches):
            steps_gen = np.random.choice(step_max + 1, size=batch_size, p=np.array(step_probs)).tolist()
            out.append(self.get_instance_data(steps_gen))
        return out

    def get_update_data(self, steps_gen: List[int]) -> List[np.ndarray]:
        """[*inputs_nnet, ctgs] (base/updater.py:915-921, 958-965)"""
        data = self.get_instance_data(steps_gen)
        assert self.nnet_input is not None, "nnet_input needed to build network inputs"
        if self.is_q:                           # one action per row (`[[action] for action in ...]`, base/updater.py:960)
            return list(self.nnet_input.to_np(data[0], data[1], [[a] for a in data[2]])) + [np.asarray(data[3])]
        return list(self.nnet_input.to_np(data[0], data[1])) + [np.asarray(data[2])]


    def _device_fn(self) -> Any:
        from deepxube_b200.base.pathfind_fns import _DeviceHeur
        fn = self.pathfind._fn
        if not isinstance(fn, _DeviceHeur):
            raise TypeError("replay-buffer labels are computed by the device (target) network: the PathFind functions must "
                            "come from a resnet_fc network (heurv / heurq_fixout / heurq_in), not a host callable")
        return fn

    def _get_labels_rb(self, input_data: Tuple, replay_data: Any, times: Any = None) -> List[float]:
        """UpdateHeurVRL._get_labels_rb (updaters/updater_rl.py:41-69): input_data = (states, goals, contexts), replay_data =
        is_solved_l -> min_a (tc + h_targ(child_a)) * not(is_solved), expansion + network + min on the device."""
        states, goals = input_data[0], input_data[1]
        fn = self._device_fn()
        rows, grows = self.domain.states_to_rows(states), self.domain.goals_to_rows(goals)
        return fn._engine().vi_targets(rows, grows, None if replay_data is None else np.asarray(replay_data, bool)).tolist()


class UpdateHeurQRLKeepGoal(UpdateHeurVRLKeepGoal):
    """the same runner over graph_q: popped edges and Node.backup_act labels (updaters/updater_rl.py:162-189)"""
    is_q = True

    def _get_labels_rb(self, input_data: Tuple, replay_data: Any, times: Any = None) -> List[float]:
        """UpdateHeurQRL._get_labels_rb (updaters/updater_rl.py:102-120): input_data = (states, goals, actions, contexts),
        replay_data = (is_solved_l, tcs, states_next) -> (tc + min_a q_targ(next, a)) * not(is_solved)."""
        goals = input_data[1]
        is_solved_l, tcs, states_next = replay_data
        assert all(float(t) == 1.0 for t in tcs), "the in-scope domains have unit transition costs"
        fn = self._device_fn()
        rows, grows = self.domain.states_to_rows(states_next), self.domain.goals_to_rows(goals)
        return fn._engine().vi_targets(rows, grows, np.asarray(is_solved_l, bool)).tolist()



class _HERMixin:
    """UpdateHER._get_her_goals (base/updater.py:485-530) over engine-backed instances: an instance that finished with a solution
    keeps its goal; every other one is relabelled with `domain.sample_goal_from_state(start states, deepest states)`, the deepest
    state being that of its deepest node that has children (the root's when nothing deeper was expanded) -- found on the device by
    `dx_backup_deepest` (csrc/k_closed.cu: k_her_depth / k_her_pick).  The goal-keeping instances are emitted first, as in the
    reference.  `_get_her_node_data` / `_get_her_edge_data` (:532-583) are what `_emit_rb` does with the returned goals: the popped
    nodes / edges of each instance with the instance's (new) goal, is_solved recomputed against it."""
    her = True

    def _group_goals(self, group: List[Any], deepest: Optional[np.ndarray]) -> Tuple[List[Any], List[Any]]:
        np.random.uniform(0, 1, size=len(group))       # `rand_keeps` (:495): drawn and unused by the reference; keeps the RNG stream aligned
        keep = [i for i in group if i.finished() and i.has_soln()]
        relabel = [i for i in group if not (i.finished() and i.has_soln())]
        goals = [i.root_node.goal for i in keep]
        if relabel:
            starts = [i.root_node.state for i in relabel]
            deep = self.domain.rows_to_states(deepest[[i._slot for i in relabel]])
            goals = goals + list(self.domain.sample_goal_from_state(starts, deep))
        return keep + relabel, goals


class UpdateHeurVRLHER(_HERMixin, UpdateHeurVRLKeepGoal):
    """`up_her_v` (updaters/updater_rl.py:203-212, 253-264)"""


class UpdateHeurQRLHER(_HERMixin, UpdateHeurQRLKeepGoal):
    """`up_her_q` (updaters/updater_rl.py:215-224, 311-323)"""
