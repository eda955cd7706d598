"""Training-data generation from the search: the value-iteration targets of the heuristic network.

Mirror of ONE runner of the reference's RL updaters for `heurv` / `heurq` with the goal kept (no HER), no replay buffer;
labels: Bellman backups, `ub_heur_solns`, LHBL tree backups:

    base/updater.py:343-400        update_runner: for each of `search_itrs` iterations -- add instances (new ones for the
                                   instances removed last iteration, same steps_gen), one PathFind.step(), remove the
                                   instances that finished or reached search_itrs, emit their data; finally emit the rest
    base/updater.py:781-791        data of an instance group = the nodes those instances popped + their labels
    updaters/updater_rl.py:143-158 V labels = Node.bellman_backup() of each popped node
    updaters/updater_rl.py:169-189 Q labels = Node.backup_act(action) of each popped edge
    base/updater.py:904-921, 938-965 -> [*nnet inputs of (state, goal[, action]), labels]

The search and the backups run on the device (graph_v engine created with max_backups > 0: every generated child is
evaluated before the CLOSED filter and each expanded node's backup is logged by k_bellman_backup); this module only owns the
instance bookkeeping.  The reference spreads runners over processes and feeds a training loop through queues; neither is part
of the hot path.
"""
from typing import Any, List, Optional, Tuple

import numpy as np

from deepxube_b200.pathfinding.graph_search import GraphSearch


class UpdateHeurVRLKeepGoal:
    is_q = False

    def __init__(self, domain: Any, pathfind: GraphSearch, search_itrs: int, nnet_input: Optional[Any] = None,
                 ub_heur_solns: bool = False, lhbl: bool = False):
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
        drained = pf.drain_backups(self.label_mode)
        rows, vals, slots = drained[0], drained[-2], drained[-1]
        # regroup the device log (iteration-major) into the reference's emission order
        order = np.concatenate([np.flatnonzero(slots == inst._slot) for g in groups for inst in g]) if len(vals) else np.zeros(0, np.int64)
        goal_of = {inst._slot: inst.root_node.goal for g in groups for inst in g}
        states = self.domain.rows_to_states(rows[order])
        goals = [goal_of[int(s)] for s in slots[order]]
        if self.is_q:
            fixed = self.domain.get_actions_fixed()
            return states, goals, [fixed[int(k)] for k in drained[1][order]], vals[order]
        return states, goals, vals[order]

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
