"""HDA*-style hash-distributed single search: W ranks emulated in one process on one GPU (the engines run
one after the other; the NCCL transport of the same driver is exercised by `tools/hda_dist.py` on >= 2 GPUs)."""
import numpy as np
import pytest

import golden_utils as gu
from oracle.domains_np import OracleDomain

pytestmark = pytest.mark.gpu


def _scramble(dom, k, seed):
    rng = np.random.default_rng(seed)
    s = dom.canonical_goal()[None].copy()
    for _ in range(k):
        s = dom.next_state(s, rng.integers(0, dom.num_actions, size=1))
    return s[0]


def _valid(dom, state, goal, acts):
    cur = state[None]
    for a in acts:
        cur = dom.next_state(cur, np.array([a]))
    return bool(dom.is_solved(cur, goal[None])[0])


def test_world1_is_bit_identical_to_the_plain_engine():
    from deepxube_b200 import _lib as L
    from deepxube_b200.hda import HdaSearch
    dom = OracleDomain("cube3")
    goal = dom.canonical_goal()
    st = _scramble(dom, 4, 1)
    hda = HdaSearch("cube3", 0, 1, batch_size=50, weight=1.0, max_nodes_per_rank=400000, capacity_slack=1.0)
    r = hda.solve(st, goal)
    eng = L.Engine("cube3", 0, "graph_v", L.DX_HEUR_ZERO, max_instances=1, max_pop_total=50, max_nodes=400000)
    eng.add_instances(st[None], goal[None], 50, 1.0)
    itrs = eng.run(10000)
    s = eng.status()[0]
    assert r["iterations"] == itrs and r["num_nodes_generated"] == s.num_nodes_generated and r["path_cost"] == s.ub
    assert r["actions"] == eng.get_path(0)[0]
    hda.close(); eng.close()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_zero_heuristic_is_optimal_and_deterministic(world):
    """With h = 0 and w = 1 batch A* is uniform-cost search: any W must return an optimal-cost valid path."""
    from deepxube_b200.hda import HdaSearch
    dom = OracleDomain("npuzzle", 3)
    goal = dom.canonical_goal()
    st = _scramble(dom, 40, 2)
    ref = HdaSearch("npuzzle", 3, 1, batch_size=200, weight=1.0, max_nodes_per_rank=400000)
    r1 = ref.solve(st, goal)
    ref.close()
    runs = []
    for _ in range(2):
        h = HdaSearch("npuzzle", 3, world, batch_size=200, weight=1.0, max_nodes_per_rank=400000)
        runs.append(h.solve(st, goal))
        h.close()
    assert runs[0] == runs[1]                                   # deterministic
    assert runs[0]["solved"] and runs[0]["path_cost"] == r1["path_cost"] == len(runs[0]["actions"])
    assert _valid(dom, st, goal, runs[0]["actions"])


@pytest.mark.parametrize("world", [1, 4])
def test_device_resident_driver_equals_host_driven(world):
    """The device-resident iteration (fixed-size segments with count headers, scalars reduced by a kernel, one host
    read per iteration) and the host-driven one (variable-size exchange, counts through the host) are the same search."""
    from deepxube_b200.hda import HdaSearch
    dom = OracleDomain("cube3")
    goal = dom.canonical_goal()
    st = _scramble(dom, 30, 3)
    out = []
    for dev, exch in ((True, "p2p"), (True, "nccl"), (False, "nccl")):
        h = HdaSearch("cube3", 0, world, batch_size=400, weight=0.8, max_nodes_per_rank=400000, device_resident=dev, exchange=exch)
        out.append(h.solve(st, goal, max_itrs=6))
        h.close()
    assert out[0] == out[1] == out[2]
    assert out[0]["iterations"] == 6 and out[0]["num_nodes_generated"] > 10000


def test_trained_net_hard_instance_world4():
    """Trained tutorial network, fp32: the distributed search must solve the KAT instance with a valid path;
    its cost / nodes generated are reported next to the single-GPU KAT values."""
    from deepxube_b200.hda import HdaSearch
    from deepxube_b200.factories.domain_factory import get_domain_from_arg
    from deepxube_b200.factories.pathfind_fns_factory import get_path_fns_nnet_par_dict
    import deepxube_b200  # noqa: F401
    k = gu.load("kat_cube3")
    domain, name = get_domain_from_arg("cube3")
    pfns, _ = get_path_fns_nnet_par_dict(domain, name, ["heurv,resnet_fc.200H_2B_bn"])
    pfns.heurv.nnet.load_state_dict(gu.state_dict_np("cube3v"))
    dom = OracleDomain("cube3")
    h = HdaSearch("cube3", 0, 4, batch_size=100, weight=0.1, nnet=pfns.heurv.nnet, precision="fp32", max_nodes_per_rank=1 << 20)
    for i in (4, 8):
        r = h.solve(k["hard_states"][i], k["hard_goals"][i], max_itrs=2000)
        assert r["solved"] and _valid(dom, k["hard_states"][i], k["hard_goals"][i], r["actions"])
        assert r["path_cost"] == len(r["actions"])
        print(f"HDA* W=4 instance {i}: cost {r['path_cost']} (1-GPU {k['hard_path_costs'][i]}), nodes {r['num_nodes_generated']} "
              f"(1-GPU {k['hard_nodes'][i]}), itrs {r['iterations']} (1-GPU {k['hard_iterations'][i]})")
        assert abs(r["path_cost"] - k["hard_path_costs"][i]) <= 12
    h.close()
