"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) in the build
container.  The reference cannot travel to the GPU box, so its outputs are committed as fixtures.

    python tools/gen_golden.py            # rewrites tests/golden/

Fixtures (all plain numpy arrays, no pickled reference classes):
  domains.npz        move tables + expand / is_solved outputs of the reference domains
  nnet.npz           weights (trained tutorial nets + small random nets) and reference network outputs
  kat_cube3.npz      tutorial/cube3/{easy,hard}.pkl states and the committed results.pkl of
                     results_brute_easy / results_heur_hard (costs, nodes, iterations, actions)
  traces.npz         per-iteration digests of reference graph_v / graph_q runs (zero heuristic,
                     injected pseudo-heuristic, trained nets), several instances in lock-step
"""
import os
import pickle
import random
import sys
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_shim import import_reference, REF_ROOT  # noqa: E402
from oracle.pseudo_heur import pseudo_v, pseudo_v_fine, make_pseudo_q, make_pseudo_q_fine  # noqa: E402

import_reference()
from deepxube.factories.domain_factory import get_domain_from_arg  # noqa: E402
from deepxube.factories.pathfinding_factory import get_pathfind_from_arg  # noqa: E402
from deepxube.factories.pathfind_fns_factory import get_path_fns_nnet_par_dict  # noqa: E402
from deepxube.pathfind_fns.fns_core import PFNsHeurVC, PFNsHeurQC  # noqa: E402
from deepxube.base.pathfinding import get_path  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def seed_all(s: int) -> None:
    random.seed(s)
    np.random.seed(s)
    torch.manual_seed(s)


def state_np(s) -> np.ndarray:
    if hasattr(s, "colors"):
        return np.asarray(s.colors, np.uint8)
    if hasattr(s, "tiles"):
        return np.asarray(s.tiles, np.uint8)
    return np.array([s.robot_x, s.robot_y], np.uint8)


def states_np(sl) -> np.ndarray:
    return np.stack([state_np(s) for s in sl]) if len(sl) else np.zeros((0, 1), np.uint8)


def digest(states: np.ndarray, g: np.ndarray) -> int:
    return zlib.crc32(np.ascontiguousarray(g, np.float64).tobytes(), zlib.crc32(np.ascontiguousarray(states, np.uint8).tobytes()))


# ------------------------------------------------------------------------------------------------
def gen_domains() -> None:
    out = {}
    seed_all(0)
    cube, _ = get_domain_from_arg("cube3")
    from deepxube.domains.cube3 import Cube3State
    ident = Cube3State(np.arange(54, dtype=np.uint8))
    out["cube3_perm"] = np.stack([state_np(cube.next_state([ident], [a])[0][0]) for a in cube.get_actions_fixed()]).astype(np.int64)
    st, goals = cube.sample_problem_instances(list(np.random.randint(0, 40, size=48)))
    out["cube3_states"] = states_np(st)
    out["cube3_goals"] = states_np(goals) if hasattr(goals[0], "tiles") else np.stack([g.colors for g in goals]).astype(np.uint8)
    ch, acts, tcs = cube.expand(st)
    out["cube3_children"] = np.stack([state_np(c) for row in ch for c in row])
    assert all(a.action == i for row in acts for i, a in enumerate(row)) and all(t == 1.0 for row in tcs for t in row)
    out["cube3_solved"] = np.array(cube.is_solved(st, goals))

    for d in (3, 4, 5):
        dom, _ = get_domain_from_arg(f"npuzzle.{d}")
        out[f"npuzzle{d}_lut"] = np.asarray(dom.swap_zero_idxs, np.int64)
        st, goals = dom.sample_problem_instances(list(np.random.randint(0, 60, size=40)))
        out[f"npuzzle{d}_states"] = states_np(st)
        out[f"npuzzle{d}_goals"] = np.stack([g.tiles for g in goals]).astype(np.uint8)
        ch, acts, tcs = dom.expand(st)
        out[f"npuzzle{d}_children"] = np.stack([state_np(c) for row in ch for c in row])
        assert all(t == 1.0 for row in tcs for t in row)
        out[f"npuzzle{d}_solved"] = np.array(dom.is_solved(st, goals))
        # wildcard goal (npuzzle.py:154-158)
        from deepxube.domains.npuzzle import NPGoal
        wg = np.stack([g.tiles for g in goals]).astype(np.uint8)
        wg[:, ::2] = d * d
        out[f"npuzzle{d}_goals_wild"] = wg
        out[f"npuzzle{d}_solved_wild"] = np.array(dom.is_solved(st, [NPGoal(x) for x in wg]))

    grid, _ = get_domain_from_arg("grid.7d")
    from deepxube.domains.grid import GridState
    cells = [GridState(x, y) for x in range(7) for y in range(7)]
    out["grid7_states"] = states_np(cells)
    ch, acts, tcs = grid.expand(cells)
    out["grid7_children"] = np.stack([state_np(c) for row in ch for c in row])
    np.savez_compressed(os.path.join(OUT, "domains.npz"), **out)
    print("domains.npz", {k: v.shape for k, v in out.items()})


# ------------------------------------------------------------------------------------------------
def sd_to_np(sd, prefix: str) -> dict:
    return {prefix + k.replace("module.", ""): v.numpy() for k, v in sd.items()}


def ref_nnet_raw(domain, domain_name, fn_str, nnet_file, states, goals):
    """Raw reference network outputs (before clipping) + processed outputs through the reference's
    own heurv / heurq callable."""
    pfns, par = get_path_fns_nnet_par_dict(domain, domain_name, [fn_str], torch.device("cpu"), nnet_files=[nnet_file])
    if hasattr(pfns, "heurv"):
        vals = np.array(pfns.heurv(states, goals, [None] * len(states)), np.float64)
    else:
        acts = domain.get_state_actions(states)
        vals = np.array(pfns.heurq(states, goals, acts, [None] * len(states)), np.float64)
    return pfns, vals


def gen_nnet() -> None:
    out = {}
    seed_all(1)
    # trained cube3 heurv net (tutorial)
    cube, _ = get_domain_from_arg("cube3")
    f = os.path.join(REF_ROOT, "tutorial/cube3/models/heur.pt")
    out.update(sd_to_np(torch.load(f, map_location="cpu", weights_only=True), "cube3v/"))
    st, goals = cube.sample_problem_instances(list(np.random.randint(0, 30, size=96)))
    _, vals = ref_nnet_raw(cube, "cube3", "heurv,resnet_fc.200H_2B_bn", f, st, goals)
    out["cube3v_states"], out["cube3v_out"] = states_np(st), vals

    # trained grid nets (tutorial/grid_tut == built-in grid semantics)
    grid, _ = get_domain_from_arg("grid.7d")
    from deepxube.domains.grid import GridState, GridGoal
    cells = [GridState(x, y) for x in range(7) for y in range(7)]
    gl = [GridGoal((3 * i) % 7, (5 * i + 1) % 7) for i in range(len(cells))]
    fv = os.path.join(REF_ROOT, "tutorial/grid_tut/flatin_v/heur.pt")
    fq = os.path.join(REF_ROOT, "tutorial/grid_tut/flatin_qfix/heur.pt")
    out.update(sd_to_np(torch.load(fv, map_location="cpu", weights_only=True), "gridv/"))
    out.update(sd_to_np(torch.load(fq, map_location="cpu", weights_only=True), "gridq/"))
    _, vals = ref_nnet_raw(grid, "grid", "heurv,resnet_fc.100H_1B_bn", fv, cells, gl)
    out["gridv_states"], out["gridv_goals"], out["gridv_out"] = states_np(cells), np.array([[g.robot_x, g.robot_y] for g in gl], np.uint8), vals
    _, vals = ref_nnet_raw(grid, "grid", "heurq_fixout,resnet_fc.100H_1B_bn", fq, cells, gl)
    out["gridq_out"] = vals

    # small random nets with non-trivial BN statistics, saved through the reference's own modules
    import tempfile
    from oracle.nnet_torch import random_state_dict
    for tag, dname, dstr, fn, in_dim, depth, odim in (("np4v", "npuzzle", "npuzzle.4", "heurv,resnet_fc.96H_2B_bn", 16, 16, 1),
                                                        ("cube3q", "cube3", "cube3", "heurq_fixout,resnet_fc.128H_1B_bn", 54, 6, 12),
                                                        ("cube3v_nobn", "cube3", "cube3", "heurv,resnet_fc.64H_1B", 54, 6, 1),
                                                        # 24-puzzle: one-hot width 625 > 8 k-blocks of 64 -> the tensor-core path's
                                                        # separate one-hot kernel (BASELINE configs[2], npuzzle.py:160-164)
                                                        ("np5v", "npuzzle", "npuzzle.5", "heurv,resnet_fc.200H_2B_bn", 25, 25, 1)):
        dom, _ = get_domain_from_arg(dstr)
        h = int(fn.split("resnet_fc.")[1].split("H")[0])
        nb = int(fn.split("H_")[1].split("B")[0])
        sd = random_state_dict(in_dim * depth, h, nb, odim, batch_norm=fn.endswith("bn"), seed=7, randomize_bn=True)
        st, goals = dom.sample_problem_instances(list(np.random.randint(0, 30, size=64)))
        if tag in ("np4v", "cube3v_nobn", "np5v"):
            # these random nets put every sampled state below zero -- clipped to 0 by process_outputs, a fixture of zeros pins
            # nothing -- so their output layer is rescaled (std 3) and shifted (10th percentile at 0: a few values still clip)
            from oracle.nnet_torch import resnet_fc_forward
            raw = np.asarray(resnet_fc_forward(sd, states_np(st).astype(np.int64), depth)).ravel()
            gain = float(np.float32(3.0 / raw.std()))
            shift = float(np.float32(-np.percentile(raw * gain, 10)))
            sd["heur.2.weight"] = sd["heur.2.weight"] * gain
            sd["heur.2.bias"] = sd["heur.2.bias"] * gain + shift
        with tempfile.NamedTemporaryFile(suffix=".pt") as tf:
            torch.save(sd, tf.name)
            _, vals = ref_nnet_raw(dom, dname, fn, tf.name, st, goals)
        out.update(sd_to_np(sd, tag + "/"))
        out[tag + "_states"], out[tag + "_out"] = states_np(st), vals
    np.savez_compressed(os.path.join(OUT, "nnet.npz"), **out)
    print("nnet.npz", len(out), "arrays")


# ------------------------------------------------------------------------------------------------
def gen_kat() -> None:
    out = {}
    for name in ("easy", "hard"):
        d = pickle.load(open(os.path.join(REF_ROOT, f"tutorial/cube3/{name}.pkl"), "rb"))
        out[f"{name}_states"] = states_np(d["states"])
        out[f"{name}_goals"] = np.stack([g.colors for g in d["goals"]]).astype(np.uint8)
    for name, rdir in (("easy", "results_brute_easy"), ("hard", "results_heur_hard")):
        r = pickle.load(open(os.path.join(REF_ROOT, f"tutorial/cube3/{rdir}/results.pkl"), "rb"))
        out[f"{name}_path_costs"] = np.array(r["path_costs"], np.float64)
        out[f"{name}_nodes"] = np.array(r["num_nodes_generated"], np.int64)
        out[f"{name}_iterations"] = np.array(r["iterations"], np.int64)
        out[f"{name}_solved"] = np.array(r["solved"])
        acts = [[a.action for a in al] for al in r["actions"]]
        out[f"{name}_actions_flat"] = np.array([x for al in acts for x in al], np.int64)
        out[f"{name}_actions_len"] = np.array([len(al) for al in acts], np.int64)
    np.savez_compressed(os.path.join(OUT, "kat_cube3.npz"), **out)
    print("kat_cube3.npz", {k: v.shape for k, v in out.items()})


# ------------------------------------------------------------------------------------------------
def run_trace(domain, pfns, pathfind_str, states, goals, max_steps):
    """Runs the reference PathFind on all instances in lock-step and records per-step digests."""
    pathfind, name, _ = get_pathfind_from_arg(domain, pfns, pathfind_str)
    insts = pathfind.make_instances(states, goals, None, True)
    pathfind.add_instances(insts)
    n = len(insts)
    rec = {k: [] for k in ("n_next", "dig_next", "open", "closed", "lb", "ub", "gen", "itr", "finished")}
    steps = 0
    while (not all(i.finished() for i in insts)) and steps < max_steps:
        pathfind.step()
        steps += 1
        for k in rec:
            rec[k].append([])
        for inst in insts:
            nodes = inst._nodes_curr
            st = states_np([x.state for x in nodes]) if nodes else np.zeros((0, 1), np.uint8)
            g = np.array([x.path_cost for x in nodes], np.float64)
            rec["n_next"][-1].append(len(nodes))
            rec["dig_next"][-1].append(digest(st, g))
            rec["open"][-1].append(len(inst.open_set))
            rec["closed"][-1].append(len(inst.closed_dict))
            rec["lb"][-1].append(inst.lb)
            rec["ub"][-1].append(inst.ub)
            rec["gen"][-1].append(inst.num_nodes_generated)
            rec["itr"][-1].append(inst.itr)
            rec["finished"][-1].append(inst.finished())
    res = {k: np.array(v) for k, v in rec.items()}
    res["path_cost"] = np.array([i.path_cost() for i in insts], np.float64)
    acts = []
    for i in insts:
        acts.append([a.action for a in get_path(i.goal_node)[1]] if i.goal_node is not None else [])
    res["actions_flat"] = np.array([x for al in acts for x in al], np.int64)
    res["actions_len"] = np.array([len(al) if i.goal_node is not None else -1 for al, i in zip(acts, insts)], np.int64)
    res["pathfind_name"] = np.array(name)
    return res


def wrap_v(fn):
    return PFNsHeurVC(heurv=lambda states, goals, ctxs: fn(states_np(states)).tolist())


def wrap_q(fn):
    return PFNsHeurQC(heurq=lambda states, goals, acts, ctxs: fn(states_np(states)).tolist())


def goals_np(goals) -> np.ndarray:
    g0 = goals[0]
    if hasattr(g0, "colors"):
        return np.stack([g.colors for g in goals]).astype(np.uint8)
    if hasattr(g0, "tiles"):
        return np.stack([g.tiles for g in goals]).astype(np.uint8)
    return np.array([[g.robot_x, g.robot_y] for g in goals], np.uint8)


def gen_traces() -> None:
    out = {}

    def add(tag, domain_str, pathfind_str, heur, steps_l, max_steps, nnet=None):
        seed_all(zlib.crc32(tag.encode()) % 1000)
        domain, dname = get_domain_from_arg(domain_str)
        states, goals = domain.sample_problem_instances(list(steps_l))
        is_q = heur.startswith("q")
        if heur in ("v_zero", "q_zero"):
            pfns, _ = get_path_fns_nnet_par_dict(domain, dname, ["heurv" if not is_q else "heurq_fixout"], torch.device("cpu"))
        elif heur == "v_pseudo":
            pfns = wrap_v(pseudo_v)
        elif heur == "q_pseudo":
            pfns = wrap_q(make_pseudo_q(domain.get_num_acts()))
        else:
            fn_str, f = nnet
            pfns, _ = get_path_fns_nnet_par_dict(domain, dname, [fn_str], torch.device("cpu"), nnet_files=[f])
        res = run_trace(domain, pfns, pathfind_str, states, goals, max_steps)
        out[tag + "/states"] = states_np(states)
        out[tag + "/goals"] = goals_np(goals)
        out[tag + "/cfg"] = np.array([domain_str, pathfind_str, heur, "" if nnet is None else nnet[0]])
        for k, v in res.items():
            out[tag + "/" + k] = v
        print(tag, res["pathfind_name"], "steps", res["itr"].shape[0], "cost", res["path_cost"], "gen", res["gen"][-1])

    add("cube3_v_zero_10B", "cube3", "graph.10B", "v_zero", [0, 1, 2] * 3, 200)
    add("cube3_v_zero_1B", "cube3", "graph", "v_zero", [0, 1, 2, 3], 400)
    add("cube3_v_pseudo", "cube3", "graph.100B_0.6W", "v_pseudo", [30, 50, 70, 100], 12)
    add("cube3_v_pseudo_big", "cube3", "graph.1000B_0.3W", "v_pseudo", [40, 60], 8)
    add("cube3_q_zero_10B", "cube3", "graph.10B", "q_zero", [0, 1, 2] * 2, 200)
    add("cube3_q_pseudo", "cube3", "graph.50B_0.6W", "q_pseudo", [30, 50, 70], 12)
    add("np4_v_pseudo", "npuzzle.4", "graph.50B_0.8W", "v_pseudo", [10, 30, 60, 90], 20)
    add("np4_v_zero", "npuzzle.4", "graph.20B", "v_zero", [0, 3, 6], 300)
    add("np3_v_zero", "npuzzle.3", "graph.5B", "v_zero", [0, 8, 20], 600)
    add("grid7_v_zero", "grid.7d", "graph", "v_zero", [0, 3, 9, 20, 50], 400)
    add("grid7_q_zero", "grid.7d", "graph.2B", "q_zero", [0, 3, 9, 20, 50], 400)
    add("grid7_v_net", "grid.7d", "graph.1B_1.0W", "v_net", list(np.arange(12) * 7), 400,
        ("heurv,resnet_fc.100H_1B_bn", os.path.join(REF_ROOT, "tutorial/grid_tut/flatin_v/heur.pt")))
    add("grid7_q_net", "grid.7d", "graph.1B_1.0W", "q_net", list(np.arange(12) * 7), 400,
        ("heurq_fixout,resnet_fc.100H_1B_bn", os.path.join(REF_ROOT, "tutorial/grid_tut/flatin_qfix/heur.pt")))
    np.savez_compressed(os.path.join(OUT, "traces.npz"), **out)
    print("traces.npz", len(out), "arrays")


def gen_lightsout() -> None:
    """lightsout.npz: the fourth domain, everything in one file so that the other fixtures stay untouched:
    move matrices, expand / is_solved outputs, a random-BN network through the reference's own modules (one-hot depth 1:
    the 0 / 1 tile value is the feature), and search traces with the zero / pseudo heuristics."""
    out = {}
    seed_all(3)
    for d in (3, 5, 7):
        dom, _ = get_domain_from_arg(f"lightsout.{d}")
        out[f"lo{d}_moves"] = np.asarray(dom.move_matrix, np.int64)
        st, goals = dom.sample_problem_instances(list(np.random.randint(0, 30, size=40)))
        out[f"lo{d}_states"] = states_np(st)
        out[f"lo{d}_goals"] = np.stack([g.tiles for g in goals]).astype(np.uint8)
        ch, acts, tcs = dom.expand(st)
        out[f"lo{d}_children"] = np.stack([state_np(c) for row in ch for c in row])
        assert all(a.action == i for row in acts for i, a in enumerate(row)) and all(t == 1.0 for row in tcs for t in row)
        out[f"lo{d}_solved"] = np.array(dom.is_solved(st, goals))
    # networks (random weights, non-trivial BN statistics) evaluated by the reference
    import tempfile
    from oracle.nnet_torch import random_state_dict
    dom, _ = get_domain_from_arg("lightsout.5")
    for tag, fn, odim in (("lo5v", "heurv,resnet_fc.96H_2B_bn", 1), ("lo5q", "heurq_fixout,resnet_fc.128H_1B_bn", 25)):
        h = int(fn.split("resnet_fc.")[1].split("H")[0])
        nb = int(fn.split("H_")[1].split("B")[0])
        sd = random_state_dict(25, h, nb, odim, batch_norm=True, seed=9, randomize_bn=True)
        with tempfile.NamedTemporaryFile(suffix=".pt") as tf:
            torch.save(sd, tf.name)
            st, goals = dom.sample_problem_instances(list(np.random.randint(0, 30, size=64)))
            _, vals = ref_nnet_raw(dom, "lightsout", fn, tf.name, st, goals)
        out.update(sd_to_np(sd, tag + "/"))
        out[tag + "_states"], out[tag + "_out"] = states_np(st), vals

    def add(tag, domain_str, pathfind_str, heur, steps_l, max_steps):
        seed_all(zlib.crc32(tag.encode()) % 1000)
        domain, dname = get_domain_from_arg(domain_str)
        states, goals = domain.sample_problem_instances(list(steps_l))
        is_q = heur.startswith("q")
        if heur in ("v_zero", "q_zero"):
            pfns, _ = get_path_fns_nnet_par_dict(domain, dname, ["heurv" if not is_q else "heurq_fixout"], torch.device("cpu"))
        elif heur == "v_pseudo":
            pfns = wrap_v(pseudo_v)
        else:
            pfns = wrap_q(make_pseudo_q(domain.get_num_acts()))
        res = run_trace(domain, pfns, pathfind_str, states, goals, max_steps)
        out[tag + "/states"] = states_np(states)
        out[tag + "/goals"] = goals_np(goals)
        out[tag + "/cfg"] = np.array([domain_str, pathfind_str, heur, ""])
        for k, v in res.items():
            out[tag + "/" + k] = v
        print(tag, res["pathfind_name"], "steps", res["itr"].shape[0], "cost", res["path_cost"], "gen", res["gen"][-1])

    add("lo3_v_zero", "lightsout.3", "graph.4B", "v_zero", [0, 2, 5, 9], 300)
    add("lo5_v_pseudo", "lightsout.5", "graph.40B_0.7W", "v_pseudo", [3, 8, 15, 30], 10)
    add("lo5_q_zero", "lightsout.5", "graph.20B", "q_zero", [0, 1, 2, 3], 60)
    add("lo7_q_pseudo", "lightsout.7", "graph.30B_0.6W", "q_pseudo", [5, 12, 25], 8)
    np.savez_compressed(os.path.join(OUT, "lightsout.npz"), **out)
    print("lightsout.npz", len(out), "arrays")


def gen_qin() -> None:
    """heurq_in.npz: action-as-input Q networks (heurq_in, fns_core.py:99-133) evaluated by the reference -- random weights
    with non-trivial BN statistics -- and one reference graph_q search driven by such a network."""
    import tempfile
    from oracle.nnet_torch import random_state_dict
    out = {}
    seed_all(5)
    files = {}
    for tag, dstr, dname, in_dim, h, nb in (("qin_cube3", "cube3", "cube3", 324 + 12, 64, 1), ("qin_lo5", "lightsout.5", "lightsout", 25 + 25, 96, 2),
                                            ("qin_grid7", "grid.7d", "grid", 28 + 4, 48, 1)):
        dom, _ = get_domain_from_arg(dstr)
        fn = f"heurq_in,resnet_fc.{h}H_{nb}B_bn"
        sd = random_state_dict(in_dim, h, nb, 1, batch_norm=True, seed=11, randomize_bn=True)
        tf = tempfile.NamedTemporaryFile(suffix=".pt", delete=False)
        torch.save(sd, tf.name)
        files[tag] = (fn, tf.name)
        st, goals = dom.sample_problem_instances(list(np.random.randint(0, 30, size=48)))
        _, vals = ref_nnet_raw(dom, dname, fn, tf.name, st, goals)
        out.update(sd_to_np(sd, tag + "/"))
        out[tag + "_states"], out[tag + "_goals"], out[tag + "_out"] = states_np(st), goals_np(goals), vals
        print(tag, vals.shape)

    seed_all(6)
    domain, dname = get_domain_from_arg("cube3")
    states, goals = domain.sample_problem_instances([3, 20, 50])
    fn, f = files["qin_cube3"]
    pfns, _ = get_path_fns_nnet_par_dict(domain, dname, [fn], torch.device("cpu"), nnet_files=[f])
    res = run_trace(domain, pfns, "graph.30B_0.7W", states, goals, 8)
    tag = "qin_cube3_trace"
    out[tag + "/states"], out[tag + "/goals"] = states_np(states), goals_np(goals)
    out[tag + "/cfg"] = np.array(["cube3", "graph.30B_0.7W", "q_in_net", fn])
    for k, v in res.items():
        out[tag + "/" + k] = v
    print(tag, res["pathfind_name"], "steps", res["itr"].shape[0], "cost", res["path_cost"], "gen", res["gen"][-1])
    np.savez_compressed(os.path.join(OUT, "heurq_in.npz"), **out)
    print("heurq_in.npz", len(out), "arrays")


def sorted_digest(states: np.ndarray, g: np.ndarray) -> int:
    """order-insensitive digest of a node set: rows sorted lexicographically together with their path costs"""
    if len(states) == 0:
        return digest(states, g)
    key = np.concatenate([np.asarray(states, np.uint8), np.asarray(g, np.float64)[:, None].astype(np.uint8)], axis=1)
    order = np.lexsort(key.T[::-1])
    return digest(np.asarray(states)[order], np.asarray(g)[order])


def gen_beam() -> None:
    """beam.npz: reference beam_v runs (beam_search.py) with the 24-bit pseudo heuristic (practically no exact cost ties between different states): per step the
    selected nodes as an ordered digest (the oracle calls the same np.argpartition) and as a set digest (the engine orders
    its selection by cost), counters, and the final solution."""
    out = {}

    def add(tag, domain_str, pathfind_str, steps_l, max_steps):
        seed_all(zlib.crc32(tag.encode()) % 1000)
        domain, dname = get_domain_from_arg(domain_str)
        states, goals = domain.sample_problem_instances(list(steps_l))
        pfns = wrap_q(make_pseudo_q_fine(domain.get_num_acts())) if pathfind_str.startswith("beam_q") else wrap_v(pseudo_v_fine)
        pathfind, name, _ = get_pathfind_from_arg(domain, pfns, pathfind_str)
        insts = pathfind.make_instances(states, goals, None, True)
        pathfind.add_instances(insts)
        rec = {k: [] for k in ("n_next", "dig_next", "setdig_next", "gen", "itr", "finished")}
        steps = 0
        while (not all(i.finished() for i in insts)) and steps < max_steps:
            pathfind.step()
            steps += 1
            for k in rec:
                rec[k].append([])
            for inst in insts:
                nodes = inst._nodes_curr
                st = states_np([x.state for x in nodes]) if nodes else np.zeros((0, 1), np.uint8)
                g = np.array([x.path_cost for x in nodes], np.float64)
                rec["n_next"][-1].append(len(nodes))
                rec["dig_next"][-1].append(digest(st, g))
                rec["setdig_next"][-1].append(sorted_digest(st, g))
                rec["gen"][-1].append(inst.num_nodes_generated)
                rec["itr"][-1].append(inst.itr)
                rec["finished"][-1].append(inst.finished())
        for k, v in rec.items():
            out[tag + "/" + k] = np.array(v)
        out[tag + "/states"], out[tag + "/goals"] = states_np(states), goals_np(goals)
        out[tag + "/cfg"] = np.array([domain_str, pathfind_str, "v_pseudo_fine", name])
        out[tag + "/path_cost"] = np.array([i.path_cost() for i in insts], np.float64)
        acts = [[a.action for a in get_path(i.goal_node)[1]] if i.goal_node is not None else [] for i in insts]
        out[tag + "/actions_flat"] = np.array([x for al in acts for x in al], np.int64)
        out[tag + "/actions_len"] = np.array([len(al) if i.goal_node is not None else -1 for al, i in zip(acts, insts)], np.int64)
        print(tag, name, "steps", steps, "cost", out[tag + "/path_cost"], "gen", rec["gen"][-1])

    add("beam_cube3_20B", "cube3", "beam_v.20B", [2, 4, 30, 60], 12)
    add("beam_cube3_500B", "cube3", "beam_v.500B", [5, 40], 8)
    add("beam_np4_30B", "npuzzle.4", "beam_v.30B", [4, 20, 60], 25)
    add("beam_lo5_10B", "lightsout.5", "beam_v.10B", [1, 3, 12], 10)
    add("beamq_cube3_20B", "cube3", "beam_q.20B", [2, 4, 30, 60], 12)
    add("beamq_cube3_400B", "cube3", "beam_q.400B", [5, 40], 8)
    add("beamq_lo5_15B", "lightsout.5", "beam_q.15B", [1, 3, 12], 10)
    np.savez_compressed(os.path.join(OUT, "beam.npz"), **out)
    print("beam.npz", len(out), "arrays")


def gen_backup() -> None:
    """backup.npz (the training-time slice): the labels UpdateHeurVRLKeepGoalABC._get_labels_no_rb
    (updaters/updater_rl.py:143-158, non-LHBL, no ub_heur_solns) reads off a graph_v search -- after every PathFind.step()
    the popped nodes it returns get node.bellman_backup() (base/pathfinding.py:41-47).  Stored flat: the popped states, the
    backup values (float64), the index of the owning instance and the number of records per step."""
    out = {}

    def add(tag, domain_str, pathfind_str, heur, steps_l, max_steps, nnet=None):
        seed_all(zlib.crc32(tag.encode()) % 1000)
        domain, dname = get_domain_from_arg(domain_str)
        states, goals = domain.sample_problem_instances(list(steps_l))
        if heur == "v_zero":
            pfns, _ = get_path_fns_nnet_par_dict(domain, dname, ["heurv"], torch.device("cpu"))
        elif heur == "v_pseudo_fine":
            pfns = wrap_v(pseudo_v_fine)
        else:
            pfns, _ = get_path_fns_nnet_par_dict(domain, dname, [nnet[0]], torch.device("cpu"), nnet_files=[nnet[1]])
        pathfind, name, _ = get_pathfind_from_arg(domain, pfns, pathfind_str)
        insts = pathfind.make_instances(states, goals, None, True)
        pathfind.add_instances(insts)
        st_l, val_l, idx_l, cnt_l, popped_l = [], [], [], [], []
        steps = 0
        while (not all(i.finished() for i in insts)) and steps < max_steps:
            owners = [k for k, inst in enumerate(insts) if not inst.finished() for _ in inst._nodes_curr]
            popped, _ = pathfind.step()
            assert len(popped) == len(owners)
            steps += 1
            vals = [node.bellman_backup() for node in popped]
            popped_l.append(popped)
            assert all(isinstance(v, float) for v in vals), set(type(v) for v in vals)
            st_l.append(states_np([n.state for n in popped]) if popped else np.zeros((0, states_np(states).shape[1]), np.uint8))
            val_l.append(np.array(vals, np.float64))
            idx_l.append(np.array(owners, np.int32))
            cnt_l.append(len(popped))
        # the updater's other label options over the same search (updaters/updater_rl.py:143-158), evaluated at the end of the
        # instances' lives like the reference does: ub_heur_solns, then lhbl (tree_backup recomputes every value from scratch)
        flat_nodes = [n for ns in popped_l for n in ns]
        for node in flat_nodes:
            node.bellman_backup()
        for node in flat_nodes:
            if node.is_solved:
                node.upper_bound_parent_path(0.0)
        out[tag + "/bk_vals_ub"] = np.array([n.backup_val for n in flat_nodes], np.float64)
        for inst in insts:
            inst.root_node.tree_backup()
        out[tag + "/bk_vals_lhbl"] = np.array([n.backup_val for n in flat_nodes], np.float64)
        out[tag + "/states"] = states_np(states)
        out[tag + "/goals"] = goals_np(goals)
        out[tag + "/cfg"] = np.array([domain_str, pathfind_str, heur, "" if nnet is None else nnet[0]])
        out[tag + "/bk_states"] = np.concatenate(st_l)
        out[tag + "/bk_vals"] = np.concatenate(val_l)
        out[tag + "/bk_inst"] = np.concatenate(idx_l)
        out[tag + "/bk_counts"] = np.array(cnt_l, np.int64)
        out[tag + "/finished"] = np.array([i.finished() for i in insts])
        out[tag + "/path_cost"] = np.array([i.path_cost() for i in insts], np.float64)
        print(tag, name, "steps", steps, "records", int(sum(cnt_l)), "finished", out[tag + "/finished"].astype(int),
              "zero-backups", int((out[tag + "/bk_vals"] == 0).sum()), "ub differs", int((out[tag + "/bk_vals_ub"] != out[tag + "/bk_vals"]).sum()),
              "lhbl differs", int((out[tag + "/bk_vals_lhbl"] != out[tag + "/bk_vals"]).sum()))

    def add_q(tag, domain_str, pathfind_str, heur, steps_l, max_steps, nnet=None):
        """graph_q: labels of the popped edges, computed the way UpdateHeurQRLKeepGoalABC._get_labels_no_rb does at the END
        of the instances' lives (updaters/updater_rl.py:169-189: per instance, edges in pop order, node.backup_act(action),
        node.backup_val overwritten as it goes), then stored in step order like the V cases."""
        seed_all(zlib.crc32(tag.encode()) % 1000)
        domain, dname = get_domain_from_arg(domain_str)
        states, goals = domain.sample_problem_instances(list(steps_l))
        if heur == "q_zero":
            pfns, _ = get_path_fns_nnet_par_dict(domain, dname, ["heurq_fixout"], torch.device("cpu"))
        elif heur == "q_pseudo_fine":
            pfns = wrap_q(make_pseudo_q_fine(domain.get_num_acts()))
        else:
            pfns, _ = get_path_fns_nnet_par_dict(domain, dname, [nnet[0]], torch.device("cpu"), nnet_files=[nnet[1]])
        pathfind, name, _ = get_pathfind_from_arg(domain, pfns, pathfind_str)
        insts = pathfind.make_instances(states, goals, None, True)
        pathfind.add_instances(insts)
        per_step, owners_l = [], []
        steps = 0
        while (not all(i.finished() for i in insts)) and steps < max_steps:
            live = [k for k, inst in enumerate(insts) if not inst.finished()]
            before = {k: len(insts[k].get_edges_popped()) for k in live}
            _, edges = pathfind.step()
            steps += 1
            owners = [k for k in live for _ in range(len(insts[k].get_edges_popped()) - before[k])]
            assert len(owners) == len(edges)
            per_step.append(edges)
            owners_l.append(owners)
        def labels(mode):
            """_get_labels_no_rb of the Q updater, verbatim order of operations, from a clean slate (backup_val = inf)"""
            for inst in insts:
                for n in [inst.root_node] + inst.root_node.get_all_descendants():
                    n.backup_val = np.inf
            res = {}
            if mode == "ub":
                for inst in insts:
                    for edge in inst.get_edges_popped():
                        if edge.node.is_solved:
                            edge.node.upper_bound_parent_path(0.0)
            elif mode == "lhbl":
                for inst in insts:
                    inst.root_node.tree_backup()
            for inst in insts:
                for edge in inst.get_edges_popped():
                    v = edge.node.backup_act(edge.action)
                    edge.node.backup_val = v
                    res[id(edge)] = v
            return res

        flat = [e for es in per_step for e in es]
        label = labels("bellman")
        lab_ub, lab_lhbl = labels("ub"), labels("lhbl")
        out[tag + "/bk_vals_ub"] = np.array([lab_ub[id(e)] for e in flat], np.float64)
        out[tag + "/bk_vals_lhbl"] = np.array([lab_lhbl[id(e)] for e in flat], np.float64)
        out[tag + "/states"] = states_np(states)
        out[tag + "/goals"] = goals_np(goals)
        out[tag + "/cfg"] = np.array([domain_str, pathfind_str, heur, "" if nnet is None else nnet[0]])
        out[tag + "/bk_states"] = states_np([e.node.state for e in flat])
        out[tag + "/bk_acts"] = np.array([e.action.action for e in flat], np.int32)
        out[tag + "/bk_vals"] = np.array([label[id(e)] for e in flat], np.float64)
        out[tag + "/bk_inst"] = np.array([k for o in owners_l for k in o], np.int32)
        out[tag + "/bk_counts"] = np.array([len(es) for es in per_step], np.int64)
        out[tag + "/finished"] = np.array([i.finished() for i in insts])
        out[tag + "/path_cost"] = np.array([i.path_cost() for i in insts], np.float64)
        print(tag, name, "steps", steps, "records", len(flat), "finished", out[tag + "/finished"].astype(int),
              "zero-backups", int((out[tag + "/bk_vals"] == 0).sum()), "ub differs", int((out[tag + "/bk_vals_ub"] != out[tag + "/bk_vals"]).sum()),
              "lhbl differs", int((out[tag + "/bk_vals_lhbl"] != out[tag + "/bk_vals"]).sum()))

    add_q("q_cube3_zero", "cube3", "graph.10B", "q_zero", [0, 1, 2, 1], 40)
    add_q("q_cube3_fine", "cube3", "graph.20B_0.6W", "q_pseudo_fine", [20, 40, 3, 60], 10)
    add_q("q_cube3_fine_short", "cube3", "graph.60B_1.0W", "q_pseudo_fine", [1, 2, 2, 3, 1, 2], 12)
    add_q("q_lo5_fine", "lightsout.5", "graph.8B_0.5W", "q_pseudo_fine", [3, 8, 1, 15], 10)
    add_q("q_grid7_net", "grid.7d", "graph.3B_1.0W", "q_net", list(np.arange(8) * 5), 60,
          ("heurq_fixout,resnet_fc.100H_1B_bn", os.path.join(REF_ROOT, "tutorial/grid_tut/flatin_qfix/heur.pt")))
    add("cube3_zero", "cube3", "graph.10B", "v_zero", [0, 1, 2, 3, 1, 2], 100)
    add("cube3_fine", "cube3", "graph.20B_0.6W", "v_pseudo_fine", [20, 40, 3, 60], 10)
    add("cube3_fine_short", "cube3", "graph.60B_1.0W", "v_pseudo_fine", [1, 2, 2, 3, 1, 2], 12)
    add("np4_fine", "npuzzle.4", "graph.30B_0.8W", "v_pseudo_fine", [10, 30, 2, 60, 90], 12)
    add("lo5_fine", "lightsout.5", "graph.8B_0.5W", "v_pseudo_fine", [3, 8, 15], 10)
    add("grid7_net", "grid.7d", "graph.3B_1.0W", "v_net", list(np.arange(8) * 5), 60,
        ("heurv,resnet_fc.100H_1B_bn", os.path.join(REF_ROOT, "tutorial/grid_tut/flatin_v/heur.pt")))
    np.savez_compressed(os.path.join(OUT, "backup.npz"), **out)
    print("backup.npz", len(out), "arrays")


def gen_vi() -> None:
    """vi.npz: replay-buffer labels straight out of the reference's own UpdateHeurVRL._get_labels_rb / UpdateHeurQRL._get_labels_rb
    (updaters/updater_rl.py:41-69, 102-120).  The updater objects are created without their process / queue plumbing
    (`__new__`), with the target-network function and the PathFind they would build injected."""
    import tempfile
    from deepxube.factories.updater_factory import updater_factory
    from deepxube.utils.timing_utils import Times
    from oracle.nnet_torch import random_state_dict
    out = {}
    z = np.load(os.path.join(OUT, "nnet.npz"))

    def net_file(tag, in_dim, h, nb, odim):
        """the random-BN nets of nnet.npz, written back to a .pt file for the reference's loader"""
        sd = {k[len(tag) + 1:]: torch.from_numpy(np.array(z[k])) for k in z.files if k.startswith(tag + "/")}
        assert sd["heur.0.weight"].shape == (h, in_dim) and sd["heur.2.weight"].shape[0] == odim
        tf = tempfile.NamedTemporaryFile(suffix=".pt", delete=False)
        torch.save(sd, tf.name)
        return tf.name

    def add(tag, domain_str, fn_str, nnet_file, steps_l, pseudo=False):
        seed_all(zlib.crc32(tag.encode()) % 1000)
        domain, dname = get_domain_from_arg(domain_str)
        is_q = fn_str.startswith("heurq")
        states, goals = domain.sample_problem_instances(list(steps_l))
        if pseudo:
            pfns = wrap_v(pseudo_v_fine)
        else:
            pfns, _ = get_path_fns_nnet_par_dict(domain, dname, [fn_str], torch.device("cpu"), nnet_files=[nnet_file])
        pathfind, _, _ = get_pathfind_from_arg(domain, pfns, "graph.1B_0.8W")
        cls = updater_factory.get_type("up_rl_q" if is_q else "up_rl_v")
        obj = cls.__new__(cls)
        obj.domain = domain
        obj.get_pathfind = lambda: pathfind
        solved = domain.is_solved(states, goals)
        ctxs = [None] * len(states)
        out[tag + "/cfg"] = np.array([domain_str, fn_str, "pseudo" if pseudo else "net"])
        out[tag + "/states"] = states_np(states)
        out[tag + "/goals"] = goals_np(goals)
        out[tag + "/solved"] = np.array(solved)
        if is_q:
            obj._get_targ_heurq_fn = lambda: pfns.heurq
            acts = [random.choice(al) for al in domain.get_state_actions(states)]
            nxt, tcs = domain.next_state(states, acts)
            assert all(t == 1.0 for t in tcs)
            lab = obj._get_labels_rb((states, goals, acts, ctxs), (solved, tcs, nxt), Times())
            out[tag + "/acts"] = np.array([a.action for a in acts], np.int32)
            out[tag + "/states_next"] = states_np(nxt)
        else:
            obj._get_targ_heurv_fn = lambda: pfns.heurv
            lab = obj._get_labels_rb((states, goals, ctxs), solved, Times())
        out[tag + "/labels"] = np.array(lab, np.float64)
        print(tag, "n", len(lab), "solved", int(np.sum(solved)), "labels", np.round(out[tag + "/labels"][:6], 3))

    rnd = lambda n, hi: [0, 0, 1, 1, 2] + list(np.random.RandomState(n).randint(0, hi, size=n))
    add("v_cube3_pseudo", "cube3", "heurv", None, rnd(120, 30), pseudo=True)
    add("v_cube3_net", "cube3", "heurv,resnet_fc.200H_2B_bn", os.path.join(REF_ROOT, "tutorial/cube3/models/heur.pt"), rnd(200, 30))
    add("v_np4_net", "npuzzle.4", "heurv,resnet_fc.96H_2B_bn", net_file("np4v", 256, 96, 2, 1), rnd(150, 40))
    add("v_grid7_net", "grid.7d", "heurv,resnet_fc.100H_1B_bn", os.path.join(REF_ROOT, "tutorial/grid_tut/flatin_v/heur.pt"), rnd(90, 12))
    add("q_cube3_net", "cube3", "heurq_fixout,resnet_fc.128H_1B_bn", net_file("cube3q", 324, 128, 1, 12), rnd(160, 30))
    add("q_grid7_net", "grid.7d", "heurq_fixout,resnet_fc.100H_1B_bn", os.path.join(REF_ROOT, "tutorial/grid_tut/flatin_qfix/heur.pt"), rnd(90, 12))
    np.savez_compressed(os.path.join(OUT, "vi.npz"), **out)
    print("vi.npz", len(out), "arrays")


def gen_her() -> None:
    """her.npz (hindsight experience replay): UpdateHER._get_her_goals / _get_her_node_data / _get_her_edge_data
    (deepxube/base/updater.py:485-583) run by the UNMODIFIED reference on its own searches.  The deepest-node rule is domain-agnostic,
    so it is also recorded for cube3 / npuzzle by handing the method a domain whose sample_goal_from_state returns the deepest states
    themselves; grid uses the reference's own Grid.sample_goal_from_state."""
    from deepxube.base.updater import UpdateHER
    from deepxube.utils.timing_utils import Times
    out = {}

    class _Stub:
        pass

    def add(tag, domain_str, pathfind_str, is_q, steps_l, max_steps):
        seed_all(zlib.crc32(tag.encode()) % 1000)
        domain, dname = get_domain_from_arg(domain_str)
        states, goals = domain.sample_problem_instances(list(steps_l))
        pfns = wrap_q(make_pseudo_q_fine(domain.get_num_acts())) if is_q else wrap_v(pseudo_v_fine)
        pathfind, name, _ = get_pathfind_from_arg(domain, pfns, pathfind_str)
        insts = pathfind.make_instances(states, goals, None, True)
        pathfind.add_instances(insts)
        steps = 0
        while (not all(i.finished() for i in insts)) and steps < max_steps:
            pathfind.step()
            steps += 1
        stub = _Stub()
        captured = {}

        class _Dom:
            def sample_goal_from_state(self, starts, deepest):
                captured["starts"], captured["deepest"] = starts, deepest
                if hasattr(domain, "sample_goal_from_state"):
                    try:
                        return domain.sample_goal_from_state(starts, deepest)
                    except Exception:
                        pass
                return list(deepest)

            def is_solved(self, st, gl):
                return domain.is_solved(st, gl)

        stub.domain = _Dom()
        order_in = list(insts)
        insts_her, goals_her = UpdateHER._get_her_goals(stub, order_in, Times())
        order = [order_in.index(i) for i in insts_her]
        n_keep = sum(1 for i in insts if i.finished() and i.has_soln())
        out[tag + "/cfg"] = np.array([domain_str, pathfind_str, "q" if is_q else "v", str(steps)])
        out[tag + "/states"] = states_np(states)
        out[tag + "/goals"] = goals_np(goals)
        out[tag + "/order"] = np.array(order, np.int32)                       # goal-keeping instances first, then the relabelled ones
        out[tag + "/n_keep"] = np.array(n_keep)
        deepest = captured.get("deepest", [])
        out[tag + "/deepest"] = states_np(deepest) if deepest else np.zeros((0, states_np(states).shape[1]), np.uint8)
        if domain_str.startswith("grid"):
            out[tag + "/goals_her"] = goals_np(goals_her)
            if is_q:
                st, gl, ac, _, sol, tcs, nxt = UpdateHER._get_her_edge_data(stub, insts_her, goals_her, Times())
                out[tag + "/her_acts"] = np.array([a.action for a in ac], np.int32)
                out[tag + "/her_next"] = states_np(nxt)
                assert all(t == 1.0 for t in tcs)
            else:
                st, gl, _, sol = UpdateHER._get_her_node_data(stub, insts_her, goals_her, Times())
            out[tag + "/her_states"] = states_np(st)
            out[tag + "/her_goals"] = goals_np(gl)
            out[tag + "/her_solved"] = np.array(sol, bool)
        print(tag, name, "steps", steps, "keep", n_keep, "relabel", len(deepest), "finished", [int(i.finished()) for i in insts])

    add("grid7_v", "grid.7d", "graph.2B_1.0W", False, [0, 3, 6, 9, 12, 2, 8, 11], 3)
    add("grid7_v_long", "grid.7d", "graph.1B_0.8W", False, [4, 10, 12, 12, 7, 9], 5)
    add("grid7_q", "grid.7d", "graph.2B_1.0W", True, [0, 3, 6, 9, 12, 2, 8, 11], 4)
    add("cube3_v", "cube3", "graph.6B_0.6W", False, [1, 2, 30, 40, 3, 25], 6)
    add("np4_v", "npuzzle.4", "graph.5B_0.8W", False, [2, 40, 60, 1, 50], 7)
    add("cube3_q", "cube3", "graph.8B_0.6W", True, [1, 30, 40, 2], 6)
    add("np3_q", "npuzzle.3", "graph.4B_0.9W", True, [3, 20, 30, 1, 25, 18], 9)
    add("lo3_v", "lightsout.3", "graph.3B_1.0W", False, [1, 5, 7, 9, 2], 5)
    add("grid7_q_long", "grid.7d", "graph.1B_0.7W", True, [12, 12, 11, 10, 9, 12, 8], 9)
    np.savez_compressed(os.path.join(OUT, "her.npz"), **out)
    print("her.npz", len(out), "arrays")


if __name__ == "__main__":
    which = sys.argv[1:] or ["domains", "nnet", "kat", "traces", "lightsout", "qin", "beam", "backup", "vi", "her"]
    if "her" in which:
        gen_her()
    if "vi" in which:
        gen_vi()
    if "backup" in which:
        gen_backup()
    if "beam" in which:
        gen_beam()
    if "qin" in which:
        gen_qin()
    if "lightsout" in which:
        gen_lightsout()
    if "domains" in which:
        gen_domains()
    if "nnet" in which:
        gen_nnet()
    if "kat" in which:
        gen_kat()
    if "traces" in which:
        gen_traces()
