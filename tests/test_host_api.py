"""CPU tests of the host-side mirror of the reference's registries / parsers / protocols (no GPU needed).
Mirrors tests/test_defaults.py and tests/test_compat.py of the reference for the in-scope names."""
import os

import numpy as np
import pytest

import deepxube_b200  # noqa: F401
from deepxube_b200 import compat_pickle
from deepxube_b200.base.factory import DelimParser, Factory
from deepxube_b200.factories.domain_factory import domain_factory, get_domain_from_arg
from deepxube_b200.factories.pathfind_fns_factory import get_path_fns_nnet_par_dict
from deepxube_b200.factories.pathfinding_factory import get_pathfind_from_arg, pathfinding_factory
from deepxube_b200.utils.command_line_utils import get_name_args
from deepxube_b200.utils.misc_utils import flatten, unflatten
from deepxube_b200.utils.timing_utils import Times

import golden_utils as gu


def test_name_args_split():
    assert get_name_args("cube3") == ("cube3", None)
    assert get_name_args("graph.10B_0.5W") == ("graph", "10B_0.5W")
    assert get_name_args("resnet_fc.1000H_4B_bn") == ("resnet_fc", "1000H_4B_bn")


def test_registries_hold_the_hot_path_names():
    assert set(domain_factory.get_all_class_names()) == {"cube3", "npuzzle", "grid", "lightsout"}
    assert set(pathfinding_factory.get_all_class_names()) == {"graph_v", "graph_q", "beam_v", "beam_q"}
    with pytest.raises(ValueError, match="Unknown domain"):
        get_domain_from_arg("sokoban")
    with pytest.raises(ValueError, match="already registered"):
        domain_factory.register_class("cube3")(object)


def test_domain_parsers():
    d, n = get_domain_from_arg("npuzzle.5")
    assert n == "npuzzle" and d.dim == 5 and d.get_num_acts() == 4
    d, n = get_domain_from_arg("grid.7d")
    assert n == "grid" and d.dim == 7
    with pytest.raises(ValueError):
        get_domain_from_arg("grid.7x")
    d, _ = get_domain_from_arg("cube3")
    assert d.get_num_acts() == 12 and [repr(a) for a in d.get_actions_fixed()][:4] == ["U-1", "U1", "D-1", "D1"]


def test_graph_search_parser_and_defaults():
    dom, name = get_domain_from_arg("cube3")
    pfv, _ = get_path_fns_nnet_par_dict(dom, name, ["heurv"])
    pfq, _ = get_path_fns_nnet_par_dict(dom, name, ["heurq_fixout"])
    p, pname, full = get_pathfind_from_arg(dom, pfv, "graph.100B_0.1W")
    assert pname == "graph_v" and full == "graph_v.100B_0.1W" and p.batch_size_default == 100 and p.weight_default == 0.1
    p, pname, _ = get_pathfind_from_arg(dom, pfq, "graph")
    assert pname == "graph_q" and (p.batch_size_default, p.weight_default, p.eps_default) == (1, 1.0, 0.0)   # test_defaults.py
    insts = p.make_instances(*dom.sample_problem_instances([0]))
    assert (insts[0].batch_size, insts[0].weight, insts[0].eps) == (1, 1.0, 0.0) and insts[0].lb == 0.0 and insts[0].ub == np.inf
    with pytest.raises(ValueError):
        get_pathfind_from_arg(dom, pfv, "graph.10X")
    with pytest.raises(ValueError, match="eps"):
        get_pathfind_from_arg(dom, pfv, "graph.10B_0.1E")
    pb, pname, _ = get_pathfind_from_arg(dom, pfv, "beam.10B")           # prefix resolution, like `graph` (pathfinding_factory.py:27-66)
    assert pname == "beam_v" and pb.beam_size_default == 10
    pb, pname, _ = get_pathfind_from_arg(dom, pfq, "beam.10B")
    assert pname == "beam_q" and pb.is_q and pb.dx_algo == "beam_q"
    with pytest.raises(ValueError, match="Could not find any compatible"):
        get_pathfind_from_arg(dom, pfv, "rollout.1.0T")
    with pytest.raises(TypeError):
        pathfinding_factory.get_type("graph_q")(dom, pfv)      # PFNsHeurV given to the Q* search


@pytest.mark.parametrize("domain_str,fn,ok", [("cube3", "heurv,resnet_fc.100H_1B_bn", True), ("cube3", "heurq_fixout,resnet_fc.100H_1B_bn", True),
                                              ("grid.7d", "heurv,resnet_fc.100H_1B_bn", True), ("grid.7d", "heurq_fixout,resnet_fc.100H_1B_bn", True),
                                              ("npuzzle.4", "heurv,resnet_fc.100H_1B_bn", True), ("npuzzle.4", "heurq_fixout,resnet_fc.100H_1B_bn", False)])
def test_fn_resolution(domain_str, fn, ok):
    """test_compat.py: every (fn x graph) string combination resolves; npuzzle has no fixed-action input."""
    dom, name = get_domain_from_arg(domain_str)
    if not ok:
        with pytest.raises(ValueError, match="Cannot build fn"):
            get_path_fns_nnet_par_dict(dom, name, [fn])
        return
    pfns, pars = get_path_fns_nnet_par_dict(dom, name, [fn])
    p, pname, _ = get_pathfind_from_arg(dom, pfns, "graph.10B")
    assert pname == ("graph_q" if fn.startswith("heurq") else "graph_v")
    nnet = (pfns.heurq if fn.startswith("heurq") else pfns.heurv).nnet
    assert nnet.res_dim == 100 and nnet.num_blocks == 1 and nnet.batch_norm
    assert nnet.out_dim == (dom.get_num_acts() if fn.startswith("heurq") else 1)


def test_resnet_fc_state_dict_roundtrip(tmp_path):
    dom, name = get_domain_from_arg("cube3")
    pfns, _ = get_path_fns_nnet_par_dict(dom, name, ["heurv,resnet_fc.200H_2B_bn"])
    nnet = pfns.heurv.nnet
    nnet.load_state_dict({"module." + k: v for k, v in gu.state_dict_np("cube3v").items()})    # tutorial weights, `module.` prefix
    f = str(tmp_path / "heur.pt")
    nnet.save(f)
    pf2, _ = get_path_fns_nnet_par_dict(dom, name, ["heurv,resnet_fc.200H_2B_bn"], nnet_files=[f])
    for k, v in nnet.state_dict().items():
        assert np.array_equal(pf2.heurv.nnet.state_dict()[k], v)
    with pytest.raises(RuntimeError, match="size mismatch|missing"):
        get_path_fns_nnet_par_dict(dom, name, ["heurv,resnet_fc.100H_2B_bn"], nnet_files=[f])
    with pytest.raises(NotImplementedError):
        get_path_fns_nnet_par_dict(dom, name, ["heurv,resnet_fc.100H_2B_wn"])


def test_delim_parser_longest_suffix_case_insensitive():
    class P(DelimParser):
        def __init__(self):
            super().__init__()
            self.add_argument("H", "h", int, "hidden")
            self.add_argument("bn", "bn", None, "batch norm")
            self.add_argument("n", "n", int, "n")

        @property
        def delim(self):
            return "_"
    assert P().parse("1000h_BN_3n") == {"h": 1000, "bn": True, "n": 3}
    with pytest.raises(ValueError):
        P().parse("12q")
    assert P().parse("") == {}


def test_times_and_flatten():
    t = Times()
    t.record_time("heur", 1.5)
    t.record_time("heur", 0.5)
    t.record_time("expand", 0.25, path=["sub"])
    assert t.get_time_str().startswith("heur: 2.00, ->sub: 0.25, Tot: 2.25")
    flat, split = flatten([[1, 2], [], [3]])
    assert flat == [1, 2, 3] and split == [2, 2] and unflatten(flat, split) == [[1, 2], [], [3]]


def test_reads_reference_pickles():
    d = compat_pickle.load(os.path.join(gu.GOLDEN, "ref_cube3_hard.pkl"))
    k = gu.load("kat_cube3")
    assert type(d["states"][0]).__module__ == "deepxube_b200.domains.cube3"
    assert np.array_equal(np.stack([s.colors for s in d["states"]]), k["hard_states"])
    g = compat_pickle.load(os.path.join(gu.GOLDEN, "ref_grid_custom_insts.pkl"))
    assert type(g["states"][0]).__name__ == "GridState" and hasattr(g["goals"][0], "robot_x")
    r = compat_pickle.load(os.path.join(gu.GOLDEN, "ref_cube3_results_heur_hard.pkl"))
    assert r["path_costs"] == k["hard_path_costs"].tolist() and [a.action for a in r["actions"][0]][:3] == [1, 8, 1]


def test_problem_instance_generation_is_host_side():
    """sample_problem_instances only produces INPUTS for the search and runs from the host move tables."""
    import random
    random.seed(0)
    from oracle.domains_np import OracleDomain
    for dstr, oname, dim in (("cube3", "cube3", 0), ("npuzzle.4", "npuzzle", 4), ("grid.7d", "grid", 7)):
        dom, _ = get_domain_from_arg(dstr)
        states, goals = dom.sample_problem_instances([0, 1, 7, 30])
        assert len(states) == len(goals) == 4
        rows, grows = dom.states_to_rows(states), dom.goals_to_rows(goals)
        assert OracleDomain(oname, dim).is_solved(rows[:1], grows[:1])[0]          # 0 steps -> solved
        walked, acts, costs = dom.random_walk(states[:2], [5, 3])
        o = OracleDomain(oname, dim)
        cur = rows[:2].copy()
        for i, al in enumerate(acts):
            for a in al:
                cur[i:i + 1] = o.next_state(cur[i:i + 1], np.array([dom.action_index(a)]))
        assert np.array_equal(dom.states_to_rows(walked), cur) and costs == [5.0, 3.0]


def test_beam_v_parser_and_rejections():
    """beam_v.<int>B (deepxube/pathfinding/beam_search.py:256-283); the sampling options are rejected, not ignored"""
    from deepxube_b200.factories.pathfinding_factory import get_pathfind_from_arg
    from deepxube_b200.pathfind_fns.fns_core import PFNsHeurVC
    dom, _ = get_domain_from_arg("cube3")
    pfns = PFNsHeurVC(heurv=lambda s, g, c: [0.0] * len(s))
    pathfind, name, _ = get_pathfind_from_arg(dom, pfns, "beam_v.25B")
    assert name == "beam_v" and pathfind.beam_size_default == 25 and pathfind.dx_algo == "beam_v"
    assert repr(pathfind) == "BeamSearchHeurNodeActsEnum(beam_size=25, temp=0.0, eps=0.0, rollout=False)"
    for bad in ("beam_v.10B_1.0T", "beam_v.10B_0.1E"):
        with pytest.raises(ValueError, match="not supported"):
            get_pathfind_from_arg(dom, pfns, bad)
    with pytest.raises(ValueError, match="Error parsing 10X"):
        get_pathfind_from_arg(dom, pfns, "beam_v.10X")


def test_updater_rl_rejects_what_it_does_not_implement():
    """the training-data slice covers heurv + graph_v only"""
    from deepxube_b200.updaters.updater_rl import UpdateHeurVRLKeepGoal
    dom, name = get_domain_from_arg("cube3")
    pfq, _ = get_path_fns_nnet_par_dict(dom, name, ["heurq_fixout"])
    pq, _, _ = get_pathfind_from_arg(dom, pfq, "graph.1B")
    with pytest.raises(TypeError, match="graph_v"):
        UpdateHeurVRLKeepGoal(dom, pq, 5)
    from deepxube_b200.updaters.updater_rl import UpdateHeurQRLKeepGoal
    assert UpdateHeurQRLKeepGoal(dom, pq, 3).is_q
    pfv, _ = get_path_fns_nnet_par_dict(dom, name, ["heurv"])
    pb, _, _ = get_pathfind_from_arg(dom, pfv, "beam_v.4B")
    with pytest.raises(TypeError, match="graph_v"):
        UpdateHeurVRLKeepGoal(dom, pb, 5)
    pv, _, _ = get_pathfind_from_arg(dom, pfv, "graph.1B_0.8W")
    up = UpdateHeurVRLKeepGoal(dom, pv, 5)
    assert up.search_itrs == 5 and pv._max_backups == 0
    with pytest.raises(TypeError, match="graph_q"):
        UpdateHeurQRLKeepGoal(dom, pv, 5)
    with pytest.raises(TypeError, match="graph_v / graph_q"):
        pb.enable_backups(10)


def test_writes_pickles_the_reference_opens(tmp_path):
    """SURVEY 8f.2, the other direction: files written with ref_classes=True carry the reference's class paths.  Checked
    structurally everywhere (no deepxube_b200 path left in the stream, round trip through our reader) and, where the
    reference is present (the build container), by loading them with the UNMODIFIED reference and using its domain on them."""
    import pickle
    import random
    random.seed(2)
    for dstr, ref_mod in (("cube3", "deepxube.domains.cube3"), ("npuzzle.4", "deepxube.domains.npuzzle"), ("grid.7d", "deepxube.domains.grid"),
                          ("lightsout.5", "deepxube.domains.lightsout")):
        dom, _ = get_domain_from_arg(dstr)
        states, goals = dom.sample_problem_instances([0, 3, 10])
        walked, acts, _ = dom.random_walk(states, [2, 2, 2])
        data = {"states": states, "goals": goals, "actions": acts, "states_on_path": [[s, w] for s, w in zip(states, walked)]}
        blob = compat_pickle.dumps(data, ref_classes=True)
        assert b"deepxube_b200" not in blob and ("c" + ref_mod + "\n").encode() in blob
        back = compat_pickle.loads(blob)                         # our reader maps the paths back
        assert np.array_equal(dom.states_to_rows(back["states"]), dom.states_to_rows(states))
        assert [dom.action_index(a) for al in back["actions"] for a in al] == [dom.action_index(a) for al in acts for a in al]
        assert b"deepxube_b200" in compat_pickle.dumps(data)     # default: our own class paths
        f = tmp_path / "d.pkl"
        compat_pickle.dump(data, str(f), ref_classes=True)
        assert f.read_bytes() == blob
    from oracle.ref_shim import import_reference, reference_available
    if not reference_available():
        return
    import_reference()
    from deepxube.factories.domain_factory import get_domain_from_arg as ref_get_domain
    for dstr in ("cube3", "npuzzle.4", "grid.7d", "lightsout.5"):
        dom, _ = get_domain_from_arg(dstr)
        rdom, _ = ref_get_domain(dstr)
        states, goals = dom.sample_problem_instances([0, 3, 10])
        walked, acts, _ = dom.random_walk(states, [1, 1, 1])
        blob = compat_pickle.dumps({"states": states, "goals": goals, "actions": acts, "walked": walked}, ref_classes=True)
        d = pickle.loads(blob)                                   # the plain unpickler of the reference's tooling
        assert type(d["states"][0]).__module__.startswith("deepxube.domains")
        assert rdom.is_solved(d["states"], d["goals"])[0]            # the 0-step instance
        nxt, tcs = rdom.next_state(d["states"], [al[0] for al in d["actions"]])       # viz_step's replay (deepxube/_cli.py:246-249)
        assert all(a == b for a, b in zip(nxt, d["walked"])) and all(t == 1.0 for t in tcs)


def test_replay_buffers_match_the_reference_deque():
    """the array-backed replay buffers behave like the reference's deque(maxlen) buffers
    (deepxube/updaters/utils/replay_buffer_utils.py:28-84): same survivors after overflowing adds, same elements for the same
    np.random seed.  Checked against a deque restatement everywhere and, in the build container, against the reference's classes."""
    from collections import deque
    from deepxube_b200.updaters.replay_buffer import ReplayBufferQ, ReplayBufferV
    rng = np.random.default_rng(0)
    S = 5
    for cap in (7, 64):
        rv, rq = ReplayBufferV(cap, S, S), ReplayBufferQ(cap, S, S)
        dv, dq = deque([], cap), deque([], cap)
        for n in (3, 5, 1, 9, 70, 2):                     # (includes an add larger than the whole buffer)
            st, gl, nx = (rng.integers(0, 250, size=(n, S)).astype(np.uint8) for _ in range(3))
            sol, act, tc = rng.integers(0, 2, size=n).astype(bool), rng.integers(0, 12, size=n), np.ones(n)
            rv.add(st, gl, sol)
            rq.add(st, gl, act, sol, tc, nx)
            dv.extend(zip(map(bytes, st), map(bytes, gl), sol.tolist()))
            dq.extend(zip(map(bytes, st), map(bytes, gl), act.tolist(), sol.tolist(), tc.tolist(), map(bytes, nx)))
            assert rv.size() == len(dv) == rq.size() and rv.max_size() == cap
            np.random.seed(n)
            idxs = np.random.randint(0, len(dv), size=11).tolist()
            np.random.seed(n)
            (a, b), c = rv.sample(11)
            assert [(bytes(x), bytes(y), bool(z)) for x, y, z in zip(a, b, c)] == [dv[i] for i in idxs]
            np.random.seed(n)
            (a, b, k), (c, t, x) = rq.sample(11)
            assert [(bytes(p), bytes(q), int(r), bool(s_), float(u), bytes(v)) for p, q, r, s_, u, v in zip(a, b, k, c, t, x)] == [dq[i] for i in idxs]
    with pytest.raises(AssertionError, match="Replay buffer size"):
        ReplayBufferV(4, S, S).sample(1)
    from oracle.ref_shim import import_reference, reference_available
    if not reference_available():
        return
    import_reference()
    from deepxube.updaters.utils.replay_buffer_utils import ReplayBufferV as RefV
    ref, mine = RefV(10), ReplayBufferV(10, S, S)
    for n in (4, 9, 3):
        st, gl = (rng.integers(0, 250, size=(n, S)).astype(np.uint8) for _ in range(2))
        sol = rng.integers(0, 2, size=n).astype(bool)
        ref.add(([bytes(r) for r in st], [bytes(r) for r in gl], [None] * n), sol.tolist())
        mine.add(st, gl, sol)
        np.random.seed(5)
        (rs, rg, _), rsol = ref.sample(6)
        np.random.seed(5)
        (ms, mg), msol = mine.sample(6)
        assert rs == [bytes(r) for r in ms] and rg == [bytes(r) for r in mg] and rsol == msol.tolist()


def test_her_goal_relabelling_order_and_rng_stream():
    """UpdateHER._get_her_goals (base/updater.py:485-530) in the updater mirror, without a device: instances that finished with a
    solution keep their goal and come first, the others get `sample_goal_from_state(start states, deepest states)` with the deepest
    rows looked up by instance slot, and one `np.random.uniform(size=len(instances))` is drawn per group like the reference's unused
    `rand_keeps` (the streams of a seeded reference run and of the mirror stay aligned)."""
    from deepxube_b200.domains.grid import Grid, GridGoal, GridState
    from deepxube_b200.updaters.updater_rl import UpdateHeurQRLHER, UpdateHeurVRLHER, _HERMixin

    class _Inst:
        def __init__(self, slot, start, goal, fin, soln):
            self._slot, self._fin, self._soln = slot, fin, soln
            self.root_node = type("N", (), {"state": start, "goal": goal})()

        def finished(self):
            return self._fin

        def has_soln(self):
            return self._soln

    dom = Grid(7)
    calls = []
    orig = dom.sample_goal_from_state
    dom.sample_goal_from_state = lambda starts, deep: (calls.append((list(starts), list(deep))), orig(starts, deep))[1]
    up = object.__new__(UpdateHeurVRLHER)
    up.domain = dom
    group = [_Inst(3, GridState(0, 0), GridGoal(6, 6), False, False), _Inst(0, GridState(1, 1), GridGoal(1, 2), True, True),
             _Inst(5, GridState(2, 2), GridGoal(5, 5), True, False), _Inst(1, GridState(3, 3), GridGoal(3, 3), True, True)]
    deepest = np.array([[9, 9], [8, 8], [7, 7], [0, 4], [6, 6], [2, 5]], np.uint8)          # rows by instance slot
    np.random.seed(11)
    insts, goals = up._group_goals(group, deepest)
    after = np.random.uniform()
    np.random.seed(11)
    np.random.uniform(0, 1, size=4)
    assert after == np.random.uniform()                                                       # exactly one draw of len(group) values
    assert [i._slot for i in insts] == [0, 1, 3, 5]                                           # goal-keeping first, order kept within each part
    assert [(g.robot_x, g.robot_y) for g in goals] == [(1, 2), (3, 3), (0, 4), (2, 5)]
    assert len(calls) == 1 and calls[0][0] == [GridState(0, 0), GridState(2, 2)] and calls[0][1] == [GridState(0, 4), GridState(2, 5)]
    calls.clear()
    insts, goals = up._group_goals(group[1:2], deepest)                                       # nobody to relabel: the domain is not asked
    assert calls == [] and [i._slot for i in insts] == [0] and (goals[0].robot_x, goals[0].robot_y) == (1, 2)
    assert UpdateHeurVRLHER.her and UpdateHeurQRLHER.her and UpdateHeurQRLHER.is_q and issubclass(UpdateHeurQRLHER, _HERMixin)
