"""Pins the CPU oracle (oracle/) to outputs of the unmodified reference (tests/golden) and to the
reference's own committed known-answer files.  CPU only."""
import numpy as np
import pytest

from oracle.domains_np import OracleDomain, cube3_perm, npuzzle_swap_lut
from oracle.nnet_torch import resnet_fc_forward, heurv_from_state_dict, heurq_from_state_dict
from oracle.pseudo_heur import pseudo_v, pseudo_v_fine, make_pseudo_q, make_pseudo_q_fine
from oracle.search import OracleSearch, solve_one

import golden_utils as gu


def test_cube3_perm_matches_reference():
    assert np.array_equal(cube3_perm(), gu.load("domains")["cube3_perm"])


@pytest.mark.parametrize("d", [3, 4, 5])
def test_npuzzle_lut_matches_reference(d):
    assert np.array_equal(npuzzle_swap_lut(d), gu.load("domains")[f"npuzzle{d}_lut"])


@pytest.mark.parametrize("name,dim,key", [("cube3", 0, "cube3"), ("npuzzle", 3, "npuzzle3"), ("npuzzle", 4, "npuzzle4"),
                                          ("npuzzle", 5, "npuzzle5"), ("grid", 7, "grid7")])
def test_expand_matches_reference(name, dim, key):
    z = gu.load("domains")
    dom = OracleDomain(name, dim)
    assert np.array_equal(dom.expand(z[key + "_states"]), z[key + "_children"])
    if key + "_solved" in z:
        assert np.array_equal(dom.is_solved(z[key + "_states"], z[key + "_goals"]), z[key + "_solved"])
    if key + "_solved_wild" in z:
        assert np.array_equal(dom.is_solved(z[key + "_states"], z[key + "_goals_wild"]), z[key + "_solved_wild"])


@pytest.mark.parametrize("d", [3, 5, 7])
def test_lightsout_matches_reference(d):
    """move matrix, expand (a border move lists its own cell twice but toggles it once, lightsout.py:186), is_solved"""
    z = gu.load("lightsout")
    dom = OracleDomain("lightsout", d)
    assert np.array_equal(dom.moves, z[f"lo{d}_moves"])
    assert np.array_equal(dom.expand(z[f"lo{d}_states"]), z[f"lo{d}_children"])
    assert np.array_equal(dom.is_solved(z[f"lo{d}_states"], z[f"lo{d}_goals"]), z[f"lo{d}_solved"])
    assert dom.input_info() == (d * d, 1)


@pytest.mark.parametrize("tag,q", [("lo5v", False), ("lo5q", True)])
def test_lightsout_nnet_matches_reference(tag, q):
    """one-hot depth 1: the tile values themselves are the network input (pytorch_models.py:15-24)"""
    from oracle.nnet_torch import heurq_from_state_dict, heurv_from_state_dict
    z = gu.load("lightsout")
    sd = gu.state_dict_torch(tag)
    fn = (heurq_from_state_dict if q else heurv_from_state_dict)(sd, 1)
    np.testing.assert_allclose(fn(z[tag + "_states"].astype(np.int64)), z[tag + "_out"], rtol=0, atol=2e-5)


@pytest.mark.parametrize("tag,dname,dim", [("qin_cube3", "cube3", 0), ("qin_lo5", "lightsout", 5), ("qin_grid7", "grid", 7)])
def test_heurq_in_nnet_matches_reference(tag, dname, dim):
    """action-as-input Q networks: one forward per (state, action), the action index one-hot encoded with depth A"""
    from oracle.nnet_torch import heurq_in_from_state_dict
    z = gu.load("heurq_in")
    dom = OracleDomain(dname, dim)
    fn = heurq_in_from_state_dict(gu.state_dict_torch(tag), dom.input_info()[1], dom.num_actions)
    x = dom.nnet_input(z[tag + "_states"], z[tag + "_goals"])
    np.testing.assert_allclose(fn(x), z[tag + "_out"], rtol=0, atol=2e-5)


def test_expand_edge_cases():
    dom = OracleDomain("cube3")
    assert dom.expand(np.zeros((0, 54), np.uint8)).shape == (0, 54)
    # every move is a permutation and a^1 is its inverse (ref: cube3.py:560-574)
    ident = np.arange(54, dtype=np.uint8)[None]
    for a in range(12):
        s1 = dom.next_state(ident, np.array([a]))
        assert sorted(s1[0].tolist()) == list(range(54))
        assert np.array_equal(dom.next_state(s1, np.array([a ^ 1])), ident)


@pytest.mark.parametrize("tag,dname,dim,depth,q", [("cube3v", "cube3", 0, 6, False), ("np4v", "npuzzle", 4, 16, False),
                                                   ("cube3q", "cube3", 0, 6, True), ("cube3v_nobn", "cube3", 0, 6, False),
                                                   ("np5v", "npuzzle", 5, 25, False)])
def test_nnet_matches_reference(tag, dname, dim, depth, q):
    z = gu.load("nnet")
    sd = gu.state_dict_torch(tag)
    fn = (heurq_from_state_dict if q else heurv_from_state_dict)(sd, depth)
    out = fn(z[tag + "_states"].astype(np.int64))
    assert (z[tag + "_out"] > 0).mean() > 0.25, "fixture pins nothing: (almost) every reference output is clipped to 0"
    # same torch ops on the same machine class -> expect (near) bit equality; allow fp32 reassociation
    np.testing.assert_allclose(out, z[tag + "_out"], rtol=0, atol=2e-5)
    # chunked evaluation (nnet_batch_size) must not change the result materially
    out2 = (heurq_from_state_dict if q else heurv_from_state_dict)(sd, depth, batch_size=7)(z[tag + "_states"].astype(np.int64))
    np.testing.assert_allclose(out2, out, rtol=0, atol=2e-5)


def test_grid_nets_match_reference():
    z = gu.load("nnet")
    dom = OracleDomain("grid", 7)
    x = dom.nnet_input(z["gridv_states"], z["gridv_goals"])
    np.testing.assert_allclose(heurv_from_state_dict(gu.state_dict_torch("gridv"), 7)(x), z["gridv_out"], atol=2e-5, rtol=0)
    np.testing.assert_allclose(heurq_from_state_dict(gu.state_dict_torch("gridq"), 7)(x), z["gridq_out"], atol=2e-5, rtol=0)


def _heur_for(cfg, dom):
    heur = str(cfg[2])
    if heur in ("v_zero", "q_zero"):
        return None
    if heur == "v_pseudo":
        return pseudo_v
    if heur == "v_pseudo_fine":
        return pseudo_v_fine
    if heur == "q_pseudo_fine":
        return make_pseudo_q_fine(dom.num_actions)
    if heur == "q_pseudo":
        return make_pseudo_q(dom.num_actions)
    if heur == "q_in_net":          # action-as-input Q network of the heurq_in trace
        from oracle.nnet_torch import heurq_in_from_state_dict
        f = heurq_in_from_state_dict(gu.state_dict_torch("qin_cube3"), dom.input_info()[1], dom.num_actions)
        return lambda s, g: f(dom.nnet_input(s, g))
    prefix = {"v_net": "gridv", "q_net": "gridq"}[heur]
    depth = dom.input_info()[1]
    mk = heurq_from_state_dict if heur == "q_net" else heurv_from_state_dict
    f = mk(gu.state_dict_torch(prefix), depth)
    return lambda s, g: f(dom.nnet_input(s, g))


@pytest.mark.parametrize("tag", gu.trace_tags())
@pytest.mark.parametrize("kept_only", [False, True])
def test_search_trace_matches_reference(tag, kept_only):
    c = gu.trace_case(tag)
    name, dim = gu.parse_domain(str(c["cfg"][0]))
    b, w = gu.parse_pathfind(str(c["cfg"][1]))
    is_q = str(c["cfg"][2]).startswith("q")
    if is_q and kept_only:
        pytest.skip("kept-only evaluation is an A* option")
    dom = OracleDomain(name, dim)
    srch = OracleSearch(dom, _heur_for(c["cfg"], dom), is_q, b, w, eval_kept_only=kept_only)
    insts = srch.add_instances(c["states"], c["goals"])
    n_steps = c["itr"].shape[0]
    for t in range(n_steps):
        assert not srch.all_finished()
        srch.step()
        for j, inst in enumerate(insts):
            ids = inst.nodes_curr
            assert len(ids) == c["n_next"][t, j], (t, j)
            assert gu.digest(inst.store.state[ids], inst.store.g[ids]) == c["dig_next"][t, j], (t, j)
            assert len(inst.open) == c["open"][t, j] and len(inst.closed) == c["closed"][t, j], (t, j)
            if kept_only and str(c["cfg"][2]).endswith("net"):
                # a different NN batch composition changes MKL's summation order by ~1 ulp (fp32)
                assert np.isclose(inst.lb, c["lb"][t, j], rtol=1e-6, atol=0) and inst.ub == c["ub"][t, j], (t, j)
            else:
                assert inst.lb == c["lb"][t, j] and inst.ub == c["ub"][t, j], (t, j)
            assert inst.num_nodes_generated == c["gen"][t, j] and inst.itr == c["itr"][t, j], (t, j)
            assert inst.finished() == c["finished"][t, j], (t, j)
    off = 0
    for j, inst in enumerate(insts):
        assert inst.path_cost() == c["path_cost"][j]
        ln = int(c["actions_len"][j])
        if ln >= 0:
            assert inst.get_path()[1] == c["actions_flat"][off: off + ln].tolist()
            off += ln


@pytest.mark.parametrize("tag", gu.beam_tags())
def test_beam_search_matches_reference(tag):
    """beam_v / beam_q: the oracle calls the same np.argpartition as the reference, so even the order of the
    selected nodes is reproduced."""
    from oracle.beam import OracleBeamQ, OracleBeamV
    c = gu.beam_case(tag)
    name, dim = gu.parse_domain(str(c["cfg"][0]))
    b, _ = gu.parse_pathfind(str(c["cfg"][1]))
    dom = OracleDomain(name, dim)
    is_q = str(c["cfg"][1]).startswith("beam_q")
    srch = OracleBeamQ(dom, make_pseudo_q_fine(dom.num_actions), b) if is_q else OracleBeamV(dom, pseudo_v_fine, b)
    insts = srch.add_instances(c["states"], c["goals"])
    for t in range(c["itr"].shape[0]):
        srch.step()
        for j, inst in enumerate(insts):
            ids = inst.nodes_curr
            if t == 0 or not c["finished"][t - 1, j]:
                assert len(ids) == c["n_next"][t, j], (t, j)
                assert gu.digest(inst.store.state[ids], inst.store.g[ids]) == c["dig_next"][t, j], (t, j)
                assert gu.sorted_digest(inst.store.state[ids], inst.store.g[ids]) == c["setdig_next"][t, j], (t, j)
            assert inst.num_nodes_generated == c["gen"][t, j] and inst.itr == c["itr"][t, j], (t, j)
            assert inst.finished() == c["finished"][t, j], (t, j)
    off = 0
    for j, inst in enumerate(insts):
        assert inst.path_cost() == c["path_cost"][j]
        ln = int(c["actions_len"][j])
        if ln >= 0:
            assert inst.get_path()[1] == c["actions_flat"][off: off + ln].tolist()
            off += ln


def test_kat_brute_easy():
    """tutorial/cube3/results_brute_easy (zero heuristic, graph.1B_1.0W): committed by the reference authors."""
    k = gu.load("kat_cube3")
    dom = OracleDomain("cube3")
    off = 0
    for i in range(10):
        if k["easy_nodes"][i] > 3000:
            off += int(k["easy_actions_len"][i])
            continue  # two 12k-node instances: covered by the slower KAT below
        r = solve_one(dom, k["easy_states"][i], k["easy_goals"][i], None, False, 1, 1.0)
        ln = int(k["easy_actions_len"][i])
        assert r["path_cost"] == k["easy_path_costs"][i] and r["num_nodes_generated"] == k["easy_nodes"][i]
        assert r["iterations"] == k["easy_iterations"][i] and r["actions"] == k["easy_actions_flat"][off: off + ln].tolist()
        off += ln


@pytest.mark.parametrize("idx", [4, 8])
def test_kat_heur_hard(idx):
    """tutorial/cube3/results_heur_hard (trained resnet_fc.200H_2B_bn, graph.100B_0.1W)."""
    k = gu.load("kat_cube3")
    dom = OracleDomain("cube3")
    f = heurv_from_state_dict(gu.state_dict_torch("cube3v"), 6)
    r = solve_one(dom, k["hard_states"][idx], k["hard_goals"][idx], lambda s, g: f(s.astype(np.int64)), False, 100, 0.1)
    off = int(k["hard_actions_len"][:idx].sum())
    ln = int(k["hard_actions_len"][idx])
    assert r["path_cost"] == k["hard_path_costs"][idx]
    assert r["num_nodes_generated"] == k["hard_nodes"][idx] and r["iterations"] == k["hard_iterations"][idx]
    assert r["actions"] == k["hard_actions_flat"][off: off + ln].tolist()


@pytest.mark.parametrize("tag", gu.backup_tags())
def test_bellman_backups_match_reference(tag):
    """SURVEY 8f.1 slice: Node.bellman_backup() of the popped nodes of every step (updaters/updater_rl.py:143-158)."""
    c = gu.backup_case(tag)
    name, dim = gu.parse_domain(str(c["cfg"][0]))
    b, w = gu.parse_pathfind(str(c["cfg"][1]))
    dom = OracleDomain(name, dim)
    is_q = tag.startswith("q_")          # graph_q: Node.backup_act of the popped edges (updaters/updater_rl.py:169-189)
    srch = OracleSearch(dom, _heur_for(c["cfg"], dom), is_q, b, w)
    insts = srch.add_instances(c["states"], c["goals"])
    for _ in range(len(c["bk_counts"])):
        srch.step()
    assert [len(x[1]) for x in srch.backup_log] == c["bk_counts"].tolist()
    st = np.concatenate([x[0] for x in srch.backup_log])
    val = np.concatenate([x[1] for x in srch.backup_log])
    idx = np.concatenate([x[2] for x in srch.backup_log])
    assert np.array_equal(st, c["bk_states"]) and np.array_equal(idx, c["bk_inst"])
    if is_q:
        assert np.array_equal(np.concatenate([x[3] for x in srch.backup_log]), c["bk_acts"])
    # the updater's other label options: ub_heur_solns, lhbl (updaters/updater_rl.py:145-155, 171-179)
    if True:
        for mode, key in (("bellman", "bk_vals"), ("ub_solns", "bk_vals_ub"), ("lhbl", "bk_vals_lhbl")):
            got = srch.labels_q(mode) if is_q else srch.labels_v(mode)
            if str(c["cfg"][2]).endswith("net"):
                np.testing.assert_allclose(got, c[key], rtol=0, atol=2e-5)
            else:
                assert np.array_equal(got, c[key]), mode
    if str(c["cfg"][2]).endswith("net"):
        np.testing.assert_allclose(val, c["bk_vals"], rtol=0, atol=2e-5)
    else:
        assert np.array_equal(val, c["bk_vals"])
    assert [i.finished() for i in insts] == c["finished"].tolist()


@pytest.mark.parametrize("tag", gu.her_tags())
def test_her_deepest_nodes_and_relabelled_data_match_reference(tag):
    """SURVEY 8f.1, hindsight relabelling: UpdateHER._get_her_goals / _get_her_node_data / _get_her_edge_data
    (base/updater.py:485-583) run by the unmodified reference (tools/gen_golden.py gen_her): which instances keep their goal, the
    deepest node with children of every other one, and -- grid, the in-scope domain that can sample a goal from a state -- the
    relabelled goals, the popped nodes / edges regrouped per instance and is_solved against the new goals."""
    c = gu.her_case(tag)
    name, dim = gu.parse_domain(str(c["cfg"][0]))
    b, w = gu.parse_pathfind(str(c["cfg"][1]))
    is_q = str(c["cfg"][2]) == "q"
    dom = OracleDomain(name, dim)
    srch = OracleSearch(dom, make_pseudo_q_fine(dom.num_actions) if is_q else pseudo_v_fine, is_q, b, w)
    insts = srch.add_instances(c["states"], c["goals"])
    for _ in range(int(c["cfg"][3])):
        srch.step()
    keep = [k for k, i in enumerate(insts) if i.finished() and i.has_soln()]
    relabel = [k for k, i in enumerate(insts) if not (i.finished() and i.has_soln())]
    assert keep + relabel == c["order"].tolist() and len(keep) == int(c["n_keep"])
    deepest = np.stack([insts[k].deepest_with_children()[0] for k in relabel]) if relabel else np.zeros((0, dom.state_len), np.uint8)
    assert np.array_equal(deepest, c["deepest"])
    if "goals_her" not in c:
        return
    goals_her = np.concatenate([c["goals"][keep], deepest])          # Grid.sample_goal_from_state: the goal IS the robot position
    assert np.array_equal(goals_her, c["goals_her"])
    inst_of = np.concatenate([x[2] for x in srch.backup_log])
    st_all = np.concatenate([x[0] for x in srch.backup_log])
    idx = np.concatenate([np.flatnonzero(inst_of == k) for k in keep + relabel])
    g_of = {k: goals_her[j] for j, k in enumerate(keep + relabel)}
    grows = np.stack([g_of[int(k)] for k in inst_of[idx]])
    assert np.array_equal(st_all[idx], c["her_states"]) and np.array_equal(grows, c["her_goals"])
    assert np.array_equal(dom.is_solved(st_all[idx], grows), c["her_solved"])
    if is_q:
        acts = np.concatenate([x[3] for x in srch.backup_log])[idx]
        assert np.array_equal(acts, c["her_acts"])
        assert np.array_equal(dom.next_state(st_all[idx], acts), c["her_next"])


VI_NETS = {"v_cube3_net": "cube3v", "v_np4_net": "np4v", "v_grid7_net": "gridv", "q_cube3_net": "cube3q", "q_grid7_net": "gridq"}


@pytest.mark.parametrize("tag", gu.vi_tags())
def test_replay_buffer_labels_match_reference(tag):
    """UpdateHeurVRL._get_labels_rb / UpdateHeurQRL._get_labels_rb (updaters/updater_rl.py:41-69, 102-120), golden = the
    reference's own methods."""
    from oracle.search import vi_targets_q, vi_targets_v
    c = gu.vi_case(tag)
    name, dim = gu.parse_domain(str(c["cfg"][0]))
    dom = OracleDomain(name, dim)
    if str(c["cfg"][2]) == "pseudo":
        got = vi_targets_v(dom, pseudo_v_fine, c["states"], c["goals"], c["solved"])
        assert np.array_equal(got, c["labels"])
        assert np.array_equal(vi_targets_v(dom, pseudo_v_fine, c["states"], c["goals"]), c["labels"])     # is_solved computed
        return
    depth = dom.input_info()[1]
    sd = gu.state_dict_torch(VI_NETS[tag])
    if tag.startswith("q_"):
        f = heurq_from_state_dict(sd, depth)
        got = vi_targets_q(dom, lambda s, g: f(dom.nnet_input(s, g)), c["states_next"], c["goals"], c["solved"])
    else:
        f = heurv_from_state_dict(sd, depth)
        got = vi_targets_v(dom, lambda s, g: f(dom.nnet_input(s, g)), c["states"], c["goals"], c["solved"])
    np.testing.assert_allclose(got, c["labels"], rtol=0, atol=2e-5)
    assert (got[c["solved"]] == 0.0).all()
