"""GPU tests through the reference-facing Python API (registries -> PathFind -> libdxb200).
Mirrors of the reference's own tests: tests/test_bruteforce.py (search correctness with the zero
heuristic, several instances in one PathFind), tests/test_domains.py (expand == next_state over the
actions; reverse actions), plus the `solve` entry point on the tutorial known-answer set."""
import os
import pickle
import random
import sys

import numpy as np
import pytest

import deepxube_b200  # noqa: F401
from deepxube_b200 import compat_pickle
from deepxube_b200.base.pathfind_fns import PFNsHeurQ, PFNsHeurV
from deepxube_b200.base.pathfinding import get_path
from deepxube_b200.factories.domain_factory import get_domain_from_arg
from deepxube_b200.factories.pathfind_fns_factory import get_path_fns_nnet_par_dict
from deepxube_b200.factories.pathfinding_factory import get_pathfind_from_arg
from deepxube_b200.pathfind_fns.fns_core import PFNsHeurQC, PFNsHeurVC
from deepxube_b200.pathfinding.utils import is_valid_soln

import golden_utils as gu
from oracle.pseudo_heur import make_pseudo_q, pseudo_v

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fn_name", ["heurv", "heurq_fixout"])
@pytest.mark.parametrize("pathfind_arg", ["graph", "graph.10B"])
def test_bruteforce(fn_name, pathfind_arg):
    """reference tests/test_bruteforce.py:26-58"""
    random.seed(0)
    dom, name = get_domain_from_arg("cube3")
    states, goals = dom.sample_problem_instances([0, 1, 2] * 3)
    pfns, _ = get_path_fns_nnet_par_dict(dom, name, [fn_name])
    pathfind, _, _ = get_pathfind_from_arg(dom, pfns, pathfind_arg)
    pathfind.add_instances(pathfind.make_instances(states, goals, None, True))
    for _ in range(200):
        pathfind.step()
    solved = []
    for inst, s, g in zip(pathfind.instances, states, goals):
        if inst.goal_node is not None:
            _, actions, _, cost = get_path(inst.goal_node)
            solved.append(is_valid_soln(s, g, actions, dom))
            assert cost == len(actions) <= 2
        else:
            solved.append(False)
    assert 100 * np.mean(solved) == 100
    pathfind.close()


@pytest.mark.parametrize("domain_str", ["cube3", "npuzzle.4", "npuzzle.3", "grid.7d", "lightsout.5"])
def test_expand_equals_next_state(domain_str):
    """reference tests/test_domains.py::test_expand"""
    random.seed(1)
    dom, _ = get_domain_from_arg(domain_str)
    states, goals = dom.sample_problem_instances(list(range(0, 20)))
    children, actions_l, tcs_l = dom.expand(states)
    assert dom.is_solved(states[:1], goals[:1]) == [True]
    for s, ch, acts, tcs in zip(states, children, actions_l, tcs_l):
        nxt, tc2 = dom.next_state([s] * len(acts), acts)
        assert {c: t for c, t in zip(ch, tcs)} == {c: t for c, t in zip(nxt, tc2)}
        assert all(a == b for a, b in zip(ch, nxt))


def test_actsrev_cube3():
    """reference tests/test_domains.py::test_actsrev: next_state(sample_rev_state(s)) == s"""
    random.seed(2)
    dom, _ = get_domain_from_arg("cube3")
    states, _ = dom.sample_problem_instances([10] * 20)
    prev, acts_rev, tcs = dom.sample_rev_state(states)
    back, _ = dom.next_state(prev, acts_rev)
    assert all(a == b for a, b in zip(back, states)) and tcs == [1.0] * 20


def test_user_callable_heuristic_host_mode():
    """Any HeurVFn / HeurQFn callable is accepted (base/pathfind_fns.py:20-31): user lambdas go through the
    host-callback mode and must give the oracle's results for the same function."""
    from oracle.domains_np import OracleDomain
    from oracle.search import OracleSearch
    random.seed(3)
    dom, _ = get_domain_from_arg("cube3")
    states, goals = dom.sample_problem_instances([20, 40, 60])
    rows, grows = dom.states_to_rows(states), dom.goals_to_rows(goals)
    for is_q in (False, True):
        if is_q:
            fq = make_pseudo_q(12)
            pfns = PFNsHeurQC(heurq=lambda s, g, a, c: fq(dom.states_to_rows(s)).tolist())
        else:
            pfns = PFNsHeurVC(heurv=lambda s, g, c: pseudo_v(dom.states_to_rows(s)).tolist())
        pathfind, pname, _ = get_pathfind_from_arg(dom, pfns, "graph.200B_0.6W")
        assert pname == ("graph_q" if is_q else "graph_v")
        pathfind.add_instances(pathfind.make_instances(states, goals))
        orc = OracleSearch(OracleDomain("cube3"), make_pseudo_q(12) if is_q else pseudo_v, is_q, 200, 0.6)
        oi = orc.add_instances(rows, grows)
        for _ in range(6):
            pathfind.step()
            orc.step()
        for inst, o in zip(pathfind.instances, oi):
            assert (inst.num_nodes_generated, inst.itr, inst.lb, inst.ub) == (o.num_nodes_generated, o.itr, o.lb, o.ub)
            got = dom.states_to_rows([n.state for n in inst.get_nodes()])
            assert np.array_equal(got, o.store.state[o.nodes_curr])
        pathfind.close()


def test_solve_cli_reproduces_tutorial_kat(tmp_path):
    """`solve --domain cube3 --fn heurv,resnet_fc.200H_2B_bn,<heur.pt> --pathfind graph.100B_0.1W --file hard.pkl`
    (the exact command of tutorial/cube3/results_heur_hard/output.txt:1) against the reference's results.pkl."""
    from deepxube_b200 import _cli
    dom, name = get_domain_from_arg("cube3")
    pfns, _ = get_path_fns_nnet_par_dict(dom, name, ["heurv,resnet_fc.200H_2B_bn"])
    pfns.heurv.nnet.load_state_dict(gu.state_dict_np("cube3v"))
    pt = str(tmp_path / "heur.pt")
    pfns.heurv.nnet.save(pt)
    res_dir = str(tmp_path / "results")
    argv, stdout = sys.argv, sys.stdout
    sys.argv = ["deepxube_b200", "solve", "--domain", "cube3", "--fn", f"heurv,resnet_fc.200H_2B_bn,{pt}", "--pathfind",
                "graph.100B_0.1W", "--file", os.path.join(gu.GOLDEN, "ref_cube3_hard.pkl"), "--results", res_dir, "--redo"]
    try:
        _cli.main()
    finally:
        sys.argv, sys.stdout = argv, stdout
    mine = pickle.load(open(os.path.join(res_dir, "results.pkl"), "rb"))
    ref = compat_pickle.load(os.path.join(gu.GOLDEN, "ref_cube3_results_heur_hard.pkl"))
    assert mine["path_costs"] == ref["path_costs"] and mine["num_nodes_generated"] == ref["num_nodes_generated"]
    assert mine["iterations"] == ref["iterations"] and mine["solved"] == ref["solved"]
    assert [[a.action for a in al] for al in mine["actions"]] == [[a.action for a in al] for al in ref["actions"]]
    assert all(np.array_equal(np.stack([s.colors for s in m]), np.stack([s.colors for s in r]))
               for m, r in zip(mine["states_on_path"], ref["states_on_path"]))
    out = open(os.path.join(res_dir, "output.txt")).read()
    assert "State: 9, SolnCost: 48.00, # Nodes Gen: 93,756, Itrs: 80" in out and "Means - SolnCost: 56.70" in out
    # resume: a second run without --redo has nothing left to do and keeps the results
    sys.argv = [a for a in sys.argv if a != "--redo"]
    argv2 = ["deepxube_b200", "solve", "--domain", "cube3", "--fn", f"heurv,resnet_fc.200H_2B_bn,{pt}", "--pathfind",
             "graph.100B_0.1W", "--file", os.path.join(gu.GOLDEN, "ref_cube3_hard.pkl"), "--results", res_dir]
    sys.argv = argv2
    try:
        _cli.main()
    finally:
        sys.argv, sys.stdout = argv, stdout
    again = pickle.load(open(os.path.join(res_dir, "results.pkl"), "rb"))
    assert again["path_costs"] == ref["path_costs"] and len(again["actions"]) == 10
    # --ref_pickle: results.pkl with the reference's class paths (what `deepxube viz --soln` would open); our reader maps them back
    res_ref = str(tmp_path / "results_ref")
    sys.argv = argv2[:-1] + [res_ref, "--ref_pickle", "--max_itrs", "200"]
    try:
        _cli.main()
    finally:
        sys.argv, sys.stdout = argv, stdout
    blob = open(os.path.join(res_ref, "results.pkl"), "rb").read()
    assert b"deepxube_b200" not in blob and b"cdeepxube.domains.cube3\nCube3State\n" in blob
    back = compat_pickle.loads(blob)
    assert back["path_costs"] == ref["path_costs"]
    assert [[a.action for a in al] for al in back["actions"]] == [[a.action for a in al] for al in ref["actions"]]


def test_device_heur_callable_matches_reference_values():
    """The heurv callable built by `--fn heurv,resnet_fc...` is usable on its own (HeurVFn protocol)."""
    dom, name = get_domain_from_arg("grid.7d")
    z = gu.load("nnet")
    pfns, _ = get_path_fns_nnet_par_dict(dom, name, ["heurq_fixout,resnet_fc.100H_1B_bn"], nnet_batch_size=17)
    pfns.heurq.nnet.load_state_dict(gu.state_dict_np("gridq"))
    states = dom.rows_to_states(z["gridv_states"])
    from deepxube_b200.domains.grid import GridGoal
    goals = [GridGoal(int(r[0]), int(r[1])) for r in z["gridv_goals"]]
    q = np.array(pfns.heurq(states, goals, dom.get_state_actions(states), [None] * len(states)))
    np.testing.assert_allclose(q, z["gridq_out"], rtol=0, atol=2e-4)


@pytest.mark.parametrize("tag,dname,dim,q,atol", [("cube3v", "cube3", 0, False, 0.08), ("gridv", "grid", 7, False, 0.05),
                                                  ("gridq", "grid", 7, True, 0.05), ("np4v", "npuzzle", 4, False, 0.15),
                                                  ("cube3q", "cube3", 0, True, 0.01),
                                                  # 24-puzzle, one-hot width 625: the non-fused input path (k_onehot_bf16 + TMA-fed GEMM)
                                                  ("np5v", "npuzzle", 5, False, 0.15)])
def test_bf16_network_within_tolerance(tag, dname, dim, q, atol):
    """tcgen05 bf16 path vs the reference's fp32 outputs.  Stated tolerance: |h_bf16 - h_ref| <= 0.08 for the
    trained cube3 network (outputs up to ~25, i.e. ~0.3 % relative; bf16 has an 8-bit mantissa and the residual
    stream is kept in bf16)."""
    from deepxube_b200 import _lib as L
    z = gu.load("nnet")
    eng = L.Engine(dname, dim, "graph_q" if q else "graph_v", L.DX_HEUR_NET_BF16, max_instances=4, max_pop_total=256, max_nodes=4096)
    blob, info = L.pack_state_dict(gu.state_dict_np(tag))
    eng.load_weights(eng.in_count, eng.depth, info["res_dim"], info["num_blocks"], info["out_dim"], info["batch_norm"], blob)
    skey = "gridv" if tag.startswith("grid") else tag
    out = eng.heur_eval(z[skey + "_states"], z["gridv_goals"] if tag.startswith("grid") else None, info["out_dim"])
    ref = z[tag + "_out"]
    np.testing.assert_allclose(out if q else out[:, 0], ref, rtol=0, atol=atol)
    eng.close()


@pytest.mark.parametrize("switch", [{}, {"DXB200_FUSE_IN": "0"}, {"DXB200_FOLD_BIAS": "0"}, {"DXB200_FUSE_OUT": "0"},
                                    {"DXB200_GEMM_CLUSTER": "1"}, {"DXB200_GEMM_CLUSTER": "4"},
                                    {"DXB200_FUSE_IN": "0", "DXB200_FOLD_BIAS": "0", "DXB200_FUSE_OUT": "0"}])
def test_bf16_kernel_variants_agree(switch, monkeypatch):
    """Every variant of the tensor-core path (one-hot generated inside the first GEMM or read from a one-hot matrix, bias
    folded into the GEMM or added in the epilogue, output layer folded into the last GEMM or separate, single CTAs /
    CTA pairs / clusters of two pairs with multicast A tiles) computes the same network: each stays within the bf16
    tolerance of the fp32 torch reference on 3000 random states (DeepCubeA-sized resnet_fc.1000H_4B_bn), on 5000 rows so
    that several row tiles and every cluster rank are exercised."""
    import torch
    from deepxube_b200 import _lib as L
    from oracle.nnet_torch import random_state_dict, resnet_fc_forward
    for k, v in switch.items():
        monkeypatch.setenv(k, v)
    rng = np.random.default_rng(5)
    st = rng.integers(0, 6, size=(5000, 54), dtype=np.uint8)
    sd = random_state_dict(324, 1000, 4, 1, True, seed=3, randomize_bn=True)
    ref = np.maximum(np.asarray(resnet_fc_forward(sd, st.astype(np.int64), 6)).ravel(), 0)
    eng = L.Engine("cube3", 0, "graph_v", L.DX_HEUR_NET_BF16, max_instances=4, max_pop_total=512, max_nodes=8192)
    blob, info = L.pack_state_dict({k: v.numpy() for k, v in sd.items()})
    eng.load_weights(54, 6, 1000, 4, 1, True, blob)
    out = eng.heur_eval(st, None, 1).ravel()
    out2 = eng.heur_eval(st, None, 1).ravel()
    eng.close()
    assert np.array_equal(out, out2), "evaluation is not repeatable"
    np.testing.assert_allclose(out, ref, rtol=0, atol=0.01)


@pytest.mark.parametrize("switch", [{}, {"DXB200_QOUT": "0"}, {"DXB200_GEMM_CLUSTER": "1"}])
def test_bf16_q_output_layer_variants_agree(switch, monkeypatch):
    """Q* network (out_dim = 12): the output layer as one more tensor-core GEMM with fp32 q-values written by the
    epilogue, or the warp-per-state kernel -- both within the bf16 tolerance of the fp32 torch reference, and
    node.heuristic = min_a q (deepxube/base/pathfinding.py:722-744)."""
    from deepxube_b200 import _lib as L
    from oracle.nnet_torch import random_state_dict, resnet_fc_forward
    for k, v in switch.items():
        monkeypatch.setenv(k, v)
    rng = np.random.default_rng(6)
    st = rng.integers(0, 6, size=(3000, 54), dtype=np.uint8)
    sd = random_state_dict(324, 1000, 4, 12, True, seed=4, randomize_bn=True)
    ref = np.maximum(np.asarray(resnet_fc_forward(sd, st.astype(np.int64), 6)), 0)
    eng = L.Engine("cube3", 0, "graph_q", L.DX_HEUR_NET_BF16, max_instances=4, max_pop_total=4096, max_nodes=8192)
    blob, info = L.pack_state_dict({k: v.numpy() for k, v in sd.items()})
    eng.load_weights(54, 6, 1000, 4, 12, True, blob)
    out = eng.heur_eval(st, None, 12)
    eng.close()
    assert out.shape == (3000, 12)
    np.testing.assert_allclose(out, ref, rtol=0, atol=0.01)


@pytest.mark.parametrize("fn,tag,q", [("heurv,resnet_fc.96H_2B_bn", "lo5v", False), ("heurq_fixout,resnet_fc.128H_1B_bn", "lo5q", True)])
def test_lightsout_through_the_registries(fn, tag, q):
    """`--domain lightsout.5 --fn heurv|heurq_fixout,resnet_fc...` resolves like in the reference (flat_sg / flat_sg_actfix
    inputs with one-hot depth 1) and the device network reproduces the reference's outputs; a short search runs through
    the PathFind API."""
    z = gu.load("lightsout")
    dom, name = get_domain_from_arg("lightsout.5")
    pfns, _ = get_path_fns_nnet_par_dict(dom, name, [fn])
    heur = pfns.heurq if q else pfns.heurv
    heur.nnet.load_state_dict(gu.state_dict_np(tag))
    states = dom.rows_to_states(z[tag + "_states"])
    goals = dom.sample_goalstate_goal_pairs(len(states))[1]
    if q:
        out = np.array(heur(states, goals, dom.get_state_actions(states), [None] * len(states)))
    else:
        out = np.array(heur(states, goals, [None] * len(states)))
    np.testing.assert_allclose(out, z[tag + "_out"], rtol=0, atol=2e-4)
    pathfind, pname, _ = get_pathfind_from_arg(dom, pfns, "graph.50B_0.8W")
    assert pname == ("graph_q" if q else "graph_v")
    random.seed(4)
    st, gl = dom.sample_problem_instances([3, 6, 9])
    pathfind.add_instances(pathfind.make_instances(st, gl))
    for _ in range(8):
        pathfind.step()
    assert all(i.num_nodes_generated > 0 for i in pathfind.instances)
    for inst, s0, g0 in zip(pathfind.instances, st, gl):
        if inst.goal_node is not None:
            _, actions, _, cost = get_path(inst.goal_node)
            assert is_valid_soln(s0, g0, actions, dom) and cost == len(actions)
    pathfind.close()


@pytest.mark.parametrize("domain_str,tag", [("cube3", "qin_cube3"), ("lightsout.5", "qin_lo5"), ("grid.7d", "qin_grid7")])
def test_heurq_in_through_the_registries(domain_str, tag):
    """`--fn heurq_in,resnet_fc...` (action-as-input Q network, fns_core.py:99-133) resolves to the `flat_sga` /
    `grid_flat_in_actin` input, reproduces the reference's q-values and drives graph_q."""
    z = gu.load("heurq_in")
    dom, name = get_domain_from_arg(domain_str)
    sd = gu.state_dict_np(tag)
    h, nb = sd["heur.0.weight"].shape[0], sum(1 for k in sd if k.endswith("layers.0.0.weight"))
    pfns, _ = get_path_fns_nnet_par_dict(dom, name, [f"heurq_in,resnet_fc.{h}H_{nb}B_bn"])
    assert isinstance(pfns, PFNsHeurQ) and pfns.heurq.nnet.act_depth == dom.get_num_acts() and pfns.heurq.nnet.out_dim == 1
    pfns.heurq.nnet.load_state_dict(sd)
    states = dom.rows_to_states(z[tag + "_states"])
    if domain_str.startswith("grid"):
        from deepxube_b200.domains.grid import GridGoal
        goals = [GridGoal(int(r[0]), int(r[1])) for r in z[tag + "_goals"]]
    else:
        goals = dom.sample_goalstate_goal_pairs(len(states))[1]
    q = np.array(pfns.heurq(states, goals, dom.get_state_actions(states), [None] * len(states)))
    np.testing.assert_allclose(q, z[tag + "_out"], rtol=0, atol=2e-4)
    pathfind, pname, _ = get_pathfind_from_arg(dom, pfns, "graph.20B_0.7W")
    assert pname == "graph_q"
    random.seed(5)
    st, gl = dom.sample_problem_instances([2, 5, 9])
    pathfind.add_instances(pathfind.make_instances(st, gl))
    for _ in range(6):
        pathfind.step()
    assert all(i.num_nodes_generated > 0 for i in pathfind.instances)
    pathfind.close()


def test_beam_v_through_the_registries():
    """`--pathfind beam_v.<B>B` (beam_search.py:202-214) through the PathFind API with a device network: instances finish as
    soon as a goal is popped, returned paths are valid, and nothing but the beam survives an iteration."""
    random.seed(7)
    dom, name = get_domain_from_arg("cube3")
    pfns, _ = get_path_fns_nnet_par_dict(dom, name, ["heurv,resnet_fc.200H_2B_bn"])
    pfns.heurv.nnet.load_state_dict(gu.state_dict_np("cube3v"))
    pathfind, pname, _ = get_pathfind_from_arg(dom, pfns, "beam_v.100B")
    assert pname == "beam_v"
    states, goals = dom.sample_problem_instances([3, 6, 9, 12])
    pathfind.add_instances(pathfind.make_instances(states, goals))
    for _ in range(40):
        if all(i.finished() for i in pathfind.instances):
            break
        pathfind.step()
    solved = 0
    for inst, s0, g0 in zip(pathfind.instances, states, goals):
        assert inst.frontier_size() <= 100
        if inst.goal_node is not None:
            _, actions, _, cost = get_path(inst.goal_node)
            assert is_valid_soln(s0, g0, actions, dom) and cost == len(actions) and inst.finished()
            solved += 1
    assert solved >= 3          # the trained tutorial network solves short scrambles greedily
    pathfind.close()


@pytest.mark.parametrize("out_dim", [1, 12])
def test_bf16_odd_row_counts(out_dim):
    """Row counts around the 128 / 256-row tile and CTA-pair boundaries (1, 127, 129, 255, 257, 1000 states): every row of
    the tensor-core path must agree with the fp32 device path, and a row's value must not depend on what else is in the batch."""
    from deepxube_b200 import _lib as L
    from oracle.nnet_torch import random_state_dict
    rng = np.random.default_rng(9)
    st = rng.integers(0, 6, size=(1000, 54), dtype=np.uint8)
    sd = random_state_dict(324, 1000, 4, out_dim, True, seed=8, randomize_bn=True)
    blob, info = L.pack_state_dict({k: v.numpy() for k, v in sd.items()})
    outs = {}
    for mode in (L.DX_HEUR_NET_FP32, L.DX_HEUR_NET_BF16):
        eng = L.Engine("cube3", 0, "graph_q" if out_dim > 1 else "graph_v", mode, max_instances=2, max_pop_total=1024, max_nodes=4096)
        eng.load_weights(54, 6, 1000, 4, out_dim, True, blob)
        outs[mode] = {n: eng.heur_eval(st[:n], None, out_dim) for n in (1, 127, 129, 255, 257, 1000)}
        eng.close()
    full = outs[L.DX_HEUR_NET_BF16][1000]
    for n, o in outs[L.DX_HEUR_NET_BF16].items():
        assert np.array_equal(o, full[:n]), n                                  # batch-independent, bit for bit
        np.testing.assert_allclose(o, outs[L.DX_HEUR_NET_FP32][n], rtol=0, atol=0.01)


def test_bf16_search_deviation_is_reported():
    """bf16 heuristics cannot promise identical node counts: solutions must stay valid and the
    deviation of cost / nodes from the fp32 KAT is bounded and printed."""
    from deepxube_b200 import _lib as L
    from oracle.domains_np import OracleDomain
    k = gu.load("kat_cube3")
    dom = OracleDomain("cube3")
    eng = L.Engine("cube3", 0, "graph_v", L.DX_HEUR_NET_BF16, max_instances=1, max_pop_total=100, max_nodes=600000)
    blob, info = L.pack_state_dict(gu.state_dict_np("cube3v"))
    eng.load_weights(54, 6, info["res_dim"], info["num_blocks"], 1, True, blob)
    dc, dn = [], []
    for i in range(10):
        eng.reset()
        eng.add_instances(k["hard_states"][i:i + 1], k["hard_goals"][i:i + 1], 100, 0.1)
        eng.run(100000)
        s = eng.status()[0]
        acts, _ = eng.get_path(0)
        cur = k["hard_states"][i:i + 1]
        for a in acts:
            cur = dom.next_state(cur, np.array([a]))
        assert s.finished and dom.is_solved(cur, k["hard_goals"][i:i + 1])[0]
        dc.append(len(acts) - k["hard_path_costs"][i])
        dn.append(s.num_nodes_generated / k["hard_nodes"][i] - 1.0)
    print(f"bf16 vs fp32 KAT: cost deviation {dc}, nodes-generated deviation {[round(100 * x, 1) for x in dn]} %")
    # measured on B200: cost deviation 0 on all ten instances; nodes generated between -33 % and +44 % (W = 0.1
    # makes the termination iteration very sensitive to h).  The bound below is what the test enforces.
    assert max(abs(x) for x in dc) <= 2 and max(abs(x) for x in dn) <= 0.75
    eng.close()


@pytest.mark.parametrize("is_q,label_mode", [(False, "bellman"), (True, "bellman"), (False, "lhbl"), (False, "ub_solns"), (True, "lhbl"), (True, "ub_solns")])
def test_updater_rl_bellman_targets_vs_oracle(is_q, label_mode):
    """SURVEY 8f.1 slice: one runner of the heurv / heurq RL updater (instances with batch size 1, replaced when they finish,
    base/updater.py:343-400) -- emitted (state, goal, [action,] backup) records equal the oracle's for the same instances."""
    from deepxube_b200.updaters.updater_rl import UpdateHeurQRLKeepGoal, UpdateHeurVRLKeepGoal
    from oracle.domains_np import OracleDomain
    from oracle.pseudo_heur import make_pseudo_q_fine, pseudo_v_fine
    from oracle.search import OracleSearch
    random.seed(5)
    dom, _ = get_domain_from_arg("cube3")
    fq = make_pseudo_q_fine(12)
    if is_q:
        pfns = PFNsHeurQC(heurq=lambda s, g, a, c: fq(dom.states_to_rows(s)).tolist())
    else:
        pfns = PFNsHeurVC(heurv=lambda s, g, c: pseudo_v_fine(dom.states_to_rows(s)).tolist())
    pathfind, pname, _ = get_pathfind_from_arg(dom, pfns, "graph.1B_0.8W")
    assert pname == ("graph_q" if is_q else "graph_v")
    made = []

    class Rec(UpdateHeurQRLKeepGoal if is_q else UpdateHeurVRLKeepGoal):
        def _make_instances(self, steps_gen):
            insts = super()._make_instances(steps_gen)
            made.append(insts)
            return insts

    steps_gen = [0, 1, 2, 1, 0, 3, 30, 40, 2, 1, 0, 25]
    search_itrs = 6
    up = Rec(dom, pathfind, search_itrs, lhbl=label_mode == "lhbl", ub_heur_solns=label_mode == "ub_solns")
    data = up.get_instance_data(steps_gen)
    states, goals, ctgs = data[0], data[1], data[-1]
    assert len(made) >= 2 and [len(m) for m in made][0] == len(steps_gen)          # finished instances were replaced

    orc = OracleSearch(OracleDomain("cube3"), fq if is_q else pseudo_v_fine, is_q, 1, 0.8)
    groups, live = [], []
    it_made = iter(made)
    removed = None
    for it in range(search_itrs):
        batch = next(it_made) if (it == 0 or removed) else []
        if batch:
            live += orc.add_instances(dom.states_to_rows([i.root_node.state for i in batch]),
                                      dom.goals_to_rows([i.root_node.goal for i in batch]))
        orc.step()
        removed = [i for i in live if i.finished() or i.itr >= search_itrs]
        live = [i for i in live if not (i.finished() or i.itr >= search_itrs)]
        if removed:
            groups.append(removed)
    if live:
        groups.append(live)
    st = np.concatenate([x[0] for x in orc.backup_log]); val = np.concatenate([x[1] for x in orc.backup_log])
    if label_mode != "bellman":
        val = orc.labels_q(label_mode) if is_q else orc.labels_v(label_mode)
    idx = np.concatenate([x[2] for x in orc.backup_log])
    order = np.concatenate([np.flatnonzero(idx == orc.instances.index(i)) for g in groups for i in g])
    assert np.array_equal(dom.states_to_rows(states), st[order])
    assert np.array_equal(ctgs, val[order]) and (ctgs[:1] == 0.0).all()            # first emitted: a solved root
    if is_q:
        acts = np.concatenate([x[3] for x in orc.backup_log])
        assert [dom.action_index(a) for a in data[2]] == acts[order].tolist()
    exp_goals = np.stack([orc.instances[k].goal for k in idx[order]])
    assert np.array_equal(dom.goals_to_rows(goals)[:, :exp_goals.shape[1]], exp_goals)
    assert len(ctgs) == sum(i.itr for i in orc.instances)                          # batch size 1: one popped node / edge per step
    # a second batch on the same updater / engine (buffers, weights and graphs are reused)
    made.clear()
    data2 = up.get_instance_data(steps_gen[:8])
    assert len(made[0]) == 8 and len(data2[-1]) >= 8 and len(pathfind.instances) <= 8
    with pytest.raises(ValueError, match="backup log"):
        up.get_instance_data(steps_gen * 4)
    pathfind.close()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_updater_rl_with_device_network(precision):
    """the updater mirror with the trained tutorial network evaluated on the device (the target network of the reference's
    update): [*inputs_nnet, ctgs] as base/updater.py:915-921 returns them, labels recomputed independently from the
    network values of each state's children."""
    from deepxube_b200 import _lib as L
    from deepxube_b200.base.nnet_input import get_nnet_input_from_arg
    from deepxube_b200.updaters.updater_rl import UpdateHeurVRLKeepGoal
    random.seed(11)
    dom, name = get_domain_from_arg("cube3")
    pfns, _ = get_path_fns_nnet_par_dict(dom, name, ["heurv,resnet_fc.200H_2B_bn"])
    pfns.heurv.nnet.load_state_dict(gu.state_dict_np("cube3v"))
    pfns.heurv.precision = precision
    pathfind, pname, _ = get_pathfind_from_arg(dom, pfns, "graph.1B_0.8W")
    up = UpdateHeurVRLKeepGoal(dom, pathfind, 8, nnet_input=get_nnet_input_from_arg(dom, name, "flat_sg"))
    steps_gen = [0, 1, 2, 3, 5, 8, 13, 21, 30, 30, 1, 2] * 4
    data = up.get_update_data(steps_gen)
    rows, ctgs = data[0], data[-1]
    assert len(data) == 2 and rows.shape == (len(ctgs), 54) and ctgs.dtype == np.float64 and len(ctgs) >= len(steps_gen)
    goal_obj = dom.sample_problem_instances([0])[1][0]
    goal = dom.goals_to_rows([goal_obj])
    solved = L.is_solved("cube3", 0, rows, np.repeat(goal, len(rows), axis=0))
    children = L.expand("cube3", 0, rows).reshape(-1, 54)
    st_c = dom.rows_to_states(children)
    h = np.array(pfns.heurv(st_c, [goal_obj] * len(st_c), [None] * len(st_c))).reshape(len(rows), 12)
    want = np.where(solved, 0.0, 1.0 + h.min(axis=1))
    np.testing.assert_allclose(ctgs, want, rtol=0, atol=2e-4 if precision == "fp32" else 0.1)
    assert (ctgs[solved] == 0.0).all() and solved.sum() >= 4
    pathfind.close()


@pytest.mark.parametrize("tag", ["v_cube3_net", "q_cube3_net", "v_grid7_net", "q_grid7_net"])
def test_updater_rl_replay_labels_through_the_registries(tag):
    """the mirror's _get_labels_rb with host State / Goal objects == the reference's labels (tests/golden/vi.npz)"""
    from deepxube_b200.updaters.updater_rl import UpdateHeurQRLKeepGoal, UpdateHeurVRLKeepGoal
    c = gu.vi_case(tag)
    is_q = tag.startswith("q_")
    dom, name = get_domain_from_arg(str(c["cfg"][0]))
    pfns, _ = get_path_fns_nnet_par_dict(dom, name, [str(c["cfg"][1])])
    heur = pfns.heurq if is_q else pfns.heurv
    heur.nnet.load_state_dict(gu.state_dict_np({"v_cube3_net": "cube3v", "q_cube3_net": "cube3q", "v_grid7_net": "gridv",
                                                "q_grid7_net": "gridq"}[tag]))
    pathfind, _, _ = get_pathfind_from_arg(dom, pfns, "graph.1B_0.8W")
    up = (UpdateHeurQRLKeepGoal if is_q else UpdateHeurVRLKeepGoal)(dom, pathfind, 4)
    states = dom.rows_to_states(c["states"])
    grows = c["goals"]
    goal_t = type(dom.sample_problem_instances([0])[1][0])
    goals = [goal_t(int(r[0]), int(r[1])) for r in grows] if name == "grid" else [goal_t(r) for r in grows]
    solved = c["solved"].tolist()
    if is_q:
        fixed = dom.get_actions_fixed()
        acts = [fixed[int(a)] for a in c["acts"]]
        nxt = dom.rows_to_states(c["states_next"])
        lab = up._get_labels_rb((states, goals, acts, [None] * len(states)), (solved, [1.0] * len(states), nxt))
    else:
        lab = up._get_labels_rb((states, goals, [None] * len(states)), solved)
    np.testing.assert_allclose(np.array(lab), c["labels"], rtol=0, atol=2e-4)
    host = (UpdateHeurVRLKeepGoal)(dom, get_pathfind_from_arg(dom, PFNsHeurVC(heurv=lambda s, g, c_: [0.0] * len(s)), "graph_v.1B")[0], 4)
    with pytest.raises(TypeError, match="device"):
        host._get_labels_rb((states, goals, None), solved)
    pathfind.close()


@pytest.mark.parametrize("is_q", [False, True])
def test_updater_rl_replay_buffer_mode(is_q):
    """SURVEY 8f.1, replay-buffer mode (base/updater.py:763-795 `_get_instance_data_rb` -> `_rb_add_sample_train_data`): the popped
    nodes / edges of every emitted instance group are added to the buffer, as many are sampled back and labelled with the device
    (target) network.  Checked: every sampled element is one the searches popped (state, goal, action), the buffer holds the newest
    `rb_size` of them, the labels equal an independent evaluation of the reference's formulas (updaters/updater_rl.py:41-69,
    102-120) with the same network through the public functions, and a second batch reuses engine, weights and buffer."""
    from deepxube_b200 import _lib as L
    from deepxube_b200.updaters.updater_rl import UpdateHeurQRLKeepGoal, UpdateHeurVRLKeepGoal
    random.seed(3)
    np.random.seed(3)
    dom, name = get_domain_from_arg("cube3")
    fn = "heurq_fixout,resnet_fc.128H_1B_bn" if is_q else "heurv,resnet_fc.200H_2B_bn"      # (the shapes of the golden nets)
    pfns, _ = get_path_fns_nnet_par_dict(dom, name, [fn])
    heur = pfns.heurq if is_q else pfns.heurv
    heur.nnet.load_state_dict(gu.state_dict_np("cube3q" if is_q else "cube3v"))
    pathfind, _, _ = get_pathfind_from_arg(dom, pfns, "graph.1B_0.8W")
    rb_size = 150
    up = (UpdateHeurQRLKeepGoal if is_q else UpdateHeurVRLKeepGoal)(dom, pathfind, 6, rb_size=rb_size)
    steps_gen = [0, 1, 2, 1, 0, 3, 30, 40, 2, 1, 0, 25] * 3
    total = 0
    for batch in range(2):
        data = up.get_instance_data(steps_gen)
        states, goals, labels = data[0], data[1], np.asarray(data[-1])
        assert len(states) == len(goals) == len(labels) >= len(steps_gen) and labels.dtype == np.float64
        total += len(states)
        assert up.rb.size() == min(rb_size, total) and up.rb.max_size() == rb_size
        rows, grows = dom.states_to_rows(states), dom.goals_to_rows(goals)
        solved = L.is_solved("cube3", 0, rows, grows)
        if is_q:
            acts = np.array([dom.action_index(a) for a in data[2]])
            nxt = L.next_state("cube3", 0, rows, acts)
            st_n = dom.rows_to_states(nxt)
            q = np.array(pfns.heurq(st_n, goals, [dom.get_actions_fixed()] * len(st_n), [None] * len(st_n)))
            want = np.where(solved, 0.0, 1.0 + q.min(axis=1))
        else:
            children = dom.rows_to_states(L.expand("cube3", 0, rows))
            h = np.array(pfns.heurv(children, [g for g in goals for _ in range(12)], [None] * len(children))).reshape(len(rows), 12)
            want = np.where(solved, 0.0, 1.0 + h.min(axis=1))
        np.testing.assert_allclose(labels, want, rtol=0, atol=2e-4)
        assert (labels[solved] == 0.0).all()
    out = up.generate(2, 8, [0.25, 0.25, 0.25, 0.25], 3)            # the runner loop over work items (base/updater.py:330-395)
    assert len(out) == 2 and all(len(o[0]) >= 8 for o in out)
    pathfind.close()


@pytest.mark.parametrize("is_q", [False, True])
def test_updater_her_relabels_goals_from_the_deepest_nodes(is_q):
    """SURVEY 8f.1, hindsight experience replay (`up_her_v` / `up_her_q`, base/updater.py:485-583, 803-818): instances that finish
    with a solution keep their goal and are emitted first; every other instance's popped nodes / edges are relabelled with the goal
    sampled from its deepest expanded node (dx_backup_deepest, pinned to the reference in tests/test_gpu_engine.py); is_solved and
    the labels are recomputed against the new goals.  Checked here: the plumbing through the registries -- which instances are
    relabelled, what `sample_goal_from_state` is asked, what enters the replay buffer, the labels of what is sampled back."""
    from deepxube_b200 import _lib as L
    from deepxube_b200.updaters.updater_rl import UpdateHeurQRLHER, UpdateHeurVRLHER, UpdateHeurVRLKeepGoal
    random.seed(5)
    np.random.seed(5)
    dom, name = get_domain_from_arg("grid.7d")
    fn = str(gu.vi_case("q_grid7_net" if is_q else "v_grid7_net")["cfg"][1])
    pfns, _ = get_path_fns_nnet_par_dict(dom, name, [fn])
    heur = pfns.heurq if is_q else pfns.heurv
    heur.nnet.load_state_dict(gu.state_dict_np("gridq" if is_q else "gridv"))
    pathfind, _, _ = get_pathfind_from_arg(dom, pfns, "graph.1B_1.0W")
    cls = UpdateHeurQRLHER if is_q else UpdateHeurVRLHER
    with pytest.raises(NotImplementedError, match="replay buffer"):
        cls(dom, pathfind, 4)                                                   # base/updater.py:483
    calls = []
    orig = dom.sample_goal_from_state

    relabelling = [False]

    def recording(starts, deepest):                 # (the domain also draws the goals of new problem instances through it)
        if relabelling[0]:
            calls.append((list(starts), list(deepest)))
        return orig(starts, deepest)

    def watch(up_):
        inner = up_._group_goals

        def group_goals(group, deepest):
            relabelling[0] = True
            try:
                return inner(group, deepest)
            finally:
                relabelling[0] = False
        up_._group_goals = group_goals
        return up_

    dom.sample_goal_from_state = recording
    itrs = 5
    up = watch(cls(dom, pathfind, itrs, rb_size=100000))
    # search_itrs iterations of batch-size-1 A*: far-away goals are not reached, near ones are
    steps_gen = [0, 1, 2, 12, 12, 11, 3, 10, 12, 1, 9, 12] * 4
    data = up.get_instance_data(steps_gen)
    states, goals, labels = data[0], data[1], np.asarray(data[-1])
    assert len(calls) >= 1 and len(states) == len(goals) == len(labels) == up.rb.size() > len(steps_gen)
    n_relabelled = sum(len(c[0]) for c in calls)
    assert 0 < n_relabelled
    for starts, deepest in calls:
        for s, d in zip(starts, deepest):            # the deepest expanded node is at most search_itrs - 1 moves from the start
            assert abs(s.robot_x - d.robot_x) + abs(s.robot_y - d.robot_y) <= itrs - 1
    # what entered the buffer: every relabelled goal is a deepest position, and each relabelled instance that expanded anything
    # below its root popped that deepest node -> a record whose state IS the new goal (solved against it)
    rb_state, rb_goal, rb_solved = up.rb._cols["state"][:up.rb.size()], up.rb._cols["goal"][:up.rb.size()], up.rb._cols["solved"][:up.rb.size()]
    deep_rows = {bytes(dom.states_to_rows([d])[0]) for _, dl in calls for d in dl}
    n_on_goal = int((rb_state == rb_goal).all(axis=1).sum())
    assert n_on_goal >= len(deep_rows) and np.array_equal(rb_solved, (rb_state == rb_goal).all(axis=1))
    assert deep_rows <= {bytes(r) for r in rb_goal}
    # labels of the sampled elements: the reference's formulas against the goals they carry
    rows, grows = dom.states_to_rows(states), dom.goals_to_rows(goals)
    solved = L.is_solved("grid", 7, rows, grows)
    if is_q:
        acts = np.array([dom.action_index(a) for a in data[2]])
        st_n = dom.rows_to_states(L.next_state("grid", 7, rows, acts))
        q = np.array(pfns.heurq(st_n, goals, [dom.get_actions_fixed()] * len(st_n), [None] * len(st_n)))
        want = np.where(solved, 0.0, 1.0 + q.min(axis=1))
    else:
        children = dom.rows_to_states(L.expand("grid", 7, rows))
        h = np.array(pfns.heurv(children, [g for g in goals for _ in range(4)], [None] * len(children))).reshape(len(rows), 4)
        want = np.where(solved, 0.0, 1.0 + h.min(axis=1))
    np.testing.assert_allclose(labels, want, rtol=0, atol=2e-4)
    assert solved.sum() > 0 and (labels[solved] == 0.0).all()
    # one expansion per instance: nothing below the root has children, the relabelled goal is the start position itself
    calls.clear()
    pathfind.reset()
    up1 = watch(cls(dom, pathfind, 1, rb_size=1000))
    d1 = up1.get_instance_data([12] * 10 + [0] * 3)
    assert len(calls) == 1 and len(d1[0]) == 13
    assert all(s == d for s, d in zip(*calls[0])) and len(calls[0][0]) >= 9
    on_goal = (dom.states_to_rows(d1[0]) == dom.goals_to_rows(d1[1])).all(axis=1)
    assert on_goal.all() and (np.asarray(d1[-1]) == 0.0).all()        # every record is its instance's root = its (new) goal
    dom.sample_goal_from_state = orig
    pathfind.close()
    cube, cname = get_domain_from_arg("cube3")
    pf_c, _ = get_path_fns_nnet_par_dict(cube, cname, ["heurv,resnet_fc.200H_2B_bn"])
    pathfind_c, _, _ = get_pathfind_from_arg(cube, pf_c, "graph.1B_1.0W")
    with pytest.raises(TypeError, match="GoalSampleableFromState"):
        UpdateHeurVRLHER(cube, pathfind_c, 3, rb_size=10)
    assert UpdateHeurVRLKeepGoal(cube, pathfind_c, 3, rb_size=10).her is False
    pathfind_c.close()
