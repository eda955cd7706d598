"""GPU parity tests added in round 2: h-injection of the in-graph device networks at BASELINE's full sizes, the
engine-lifecycle entry points (weight reload, instance removal / slot recycling), negative heuristic values, input
validation and HDA* error propagation.  Everything goes through the C ABI (ctypes) or the host mirror above it."""
import numpy as np
import pytest

import golden_utils as gu
from oracle.domains_np import OracleDomain
from oracle.nnet_torch import heurv_from_state_dict, random_state_dict
from oracle.pseudo_heur import make_pseudo_q, make_pseudo_q_fine, pseudo_v, pseudo_v_fine
from oracle.search import OracleSearch

pytestmark = pytest.mark.gpu


def _lib():
    from deepxube_b200 import _lib
    return _lib


def _scrambles(dom, n, k, seed):
    rng = np.random.default_rng(seed)
    goals = np.tile(dom.canonical_goal(), (n, 1))
    st = goals.copy()
    for _ in range(k):
        st = dom.next_state(st, rng.integers(0, dom.num_actions, size=n))
    return st, goals


def _spread_net(in_dim, h, nb, od, seed):
    """random-BN resnet_fc whose outputs are spread over positive values (a plain random init mostly lands below zero,
    where process_outputs clips everything to 0 and two different networks look identical)"""
    sd = {k: v.numpy().copy() for k, v in random_state_dict(in_dim, h, nb, od, True, seed=seed, randomize_bn=True).items()}
    sd["heur.2.weight"] *= 12.0
    sd["heur.2.bias"] = sd["heur.2.bias"] * 12.0 + 4.0
    return sd


def _compare_iteration(t, eng, oi):
    for j, (o, s) in enumerate(zip(oi, eng.status())):
        ids = o.nodes_curr
        cs, cg, _ = eng.curr_nodes(j, len(ids) + 8)
        if o.itr == t + 1:
            assert len(cs) == len(ids), (t, j, len(cs), len(ids))
            assert np.array_equal(cs, o.store.state[ids]) and np.array_equal(cg.astype(np.float64), o.store.g[ids]), (t, j)
        assert s.lb == o.lb and s.ub == o.ub and s.num_nodes_generated == o.num_nodes_generated, (t, j, s.lb, o.lb)
        assert s.itr == o.itr and bool(s.finished) == o.finished(), (t, j)
        if not s.finished:
            assert s.open_size == len(o.open), (t, j)


# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode_name", ["bf16", "fp32"])
@pytest.mark.parametrize("algo,n,iters", [("graph_v", 4, 6), ("graph_q", 2, 7)])
def test_in_graph_network_engine_vs_oracle_fed_the_engines_h(mode_name, algo, n, iters):
    """SURVEY 8c.1 (h-injection) at BASELINE's full size: cube3, B = 10 000, W = 0.6, resnet_fc.1000H_4B_bn evaluated INSIDE the
    captured device iteration (dx_run), against the CPU oracle whose heuristic callable returns the same device network's
    values (dx_heur_eval on a second engine with the same weights).  Every CLOSED decision, pop order, lb / ub / OPEN size
    and counter must be bit-equal in every iteration -- this pins the in-graph network path (fused one-hot input GEMM,
    absolute row indexing, kept-children-only evaluation) to the search logic, not just to sortedness properties."""
    L = _lib()
    is_q = algo == "graph_q"
    dom = OracleDomain("cube3")
    st, goals = _scrambles(dom, n, 150, 31)
    b, w = 10000, 0.6
    od = 12 if is_q else 1
    blob, _ = L.pack_state_dict(_spread_net(324, 1000, 4, od, 2))
    mode = L.DX_HEUR_NET_BF16 if mode_name == "bf16" else L.DX_HEUR_NET_FP32
    eng = L.Engine("cube3", 0, algo, mode, max_instances=n, max_pop_total=n * b, max_nodes=n * b * (1 if is_q else 12) * (iters + 1) + 1024)
    eng.load_weights(54, 6, 1000, 4, od, True, blob)
    ev = L.Engine("cube3", 0, algo, mode, max_instances=1, max_pop_total=8192, max_nodes=1024)
    ev.load_weights(54, 6, 1000, 4, od, True, blob)

    def heur(s, g):
        out = ev.heur_eval(s, None, od)
        return out.astype(np.float64) if is_q else out[:, 0].astype(np.float64)

    assert (heur(st, goals) > 0).mean() > 0.5, "network values must not be clipped away (the search would be breadth-first)"
    orc = OracleSearch(dom, heur, is_q, b, w)
    oi = orc.add_instances(st, goals)
    eng.add_instances(st, goals, b, w)
    for t in range(iters):
        orc.step()
        assert eng.run(1) == 1
        _compare_iteration(t, eng, oi)
    assert eng.counters().children_generated == sum(o.num_nodes_generated for o in oi)
    eng.close(); ev.close()


def test_config_c1_grid7_100_of_100_with_the_engines_h():
    """BASELINE config 1 (grid.7d, 100 instances, resnet_fc.100H_2B_bn, graph.1B_1.0W): with the oracle fed the device
    network's own fp32 values all 100 instances must match exactly (iterations, nodes generated, path cost, actions).
    Against the torch-evaluated network a handful of instances differ by fp32 re-association of exact ties; those are
    listed by tests/test_gpu_engine.py::test_config_c1_grid7_100_instances."""
    L = _lib()
    dom = OracleDomain("grid", 7)
    rng = np.random.default_rng(0)
    n = 100
    st = rng.integers(0, 7, size=(n, 2), dtype=np.uint8)
    goals = rng.integers(0, 7, size=(n, 2), dtype=np.uint8)
    sd = random_state_dict(28, 100, 2, 1, batch_norm=True, seed=3, randomize_bn=True)
    blob, _ = L.pack_state_dict({k: v.numpy() for k, v in sd.items()})
    eng = L.Engine("grid", 7, "graph_v", L.DX_HEUR_NET_FP32, max_instances=n, max_pop_total=n, max_nodes=100000)
    eng.load_weights(4, 7, 100, 2, 1, True, blob)
    ev = L.Engine("grid", 7, "graph_v", L.DX_HEUR_NET_FP32, max_instances=1, max_pop_total=256, max_nodes=1024)
    ev.load_weights(4, 7, 100, 2, 1, True, blob)
    orc = OracleSearch(dom, lambda s, g: ev.heur_eval(s, g, 1)[:, 0].astype(np.float64), False, 1, 1.0)
    oi = orc.add_instances(st, goals)
    while not orc.all_finished():
        orc.step()
    eng.add_instances(st, goals, 1, 1.0)
    eng.run(10000)
    assert eng.all_finished()
    for j, (o, s) in enumerate(zip(oi, eng.status())):
        acts, _ = eng.get_path(j)
        assert s.itr == o.itr and s.num_nodes_generated == o.num_nodes_generated and s.ub == o.ub, j
        assert acts == o.get_path()[1], j
    eng.close(); ev.close()


# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode_name", ["fp32", "bf16"])
def test_weights_can_be_reloaded_in_place(mode_name):
    """dx_load_weights twice on one engine (an RL loop's updated network): the second set of parameters is the one
    that is evaluated afterwards, and a search after the reload equals a fresh engine's (captured graphs survive)."""
    L = _lib()
    dom = OracleDomain("cube3")
    mode = L.DX_HEUR_NET_BF16 if mode_name == "bf16" else L.DX_HEUR_NET_FP32
    st, goals = _scrambles(dom, 3, 40, 5)
    blobs = [L.pack_state_dict(_spread_net(324, 256, 2, 1, s))[0] for s in (1, 2)]

    def fresh(blob):
        e = L.Engine("cube3", 0, "graph_v", mode, max_instances=3, max_pop_total=600, max_nodes=200000)
        e.load_weights(54, 6, 256, 2, 1, True, blob)
        return e

    def search(e):
        e.reset()
        e.add_instances(st, goals, 200, 0.6)
        e.run(6)
        return [(s.num_nodes_generated, s.lb, s.open_size) for s in e.status()], [e.curr_nodes(j, 300)[2].copy() for j in range(3)]

    e = fresh(blobs[0])
    h_a = e.heur_eval(st, None, 1)
    r_a = search(e)
    e.load_weights(54, 6, 256, 2, 1, True, blobs[1])            # in place
    h_b = e.heur_eval(st, None, 1)
    r_b = search(e)
    f = fresh(blobs[1])
    assert np.array_equal(h_b, f.heur_eval(st, None, 1)) and not np.array_equal(h_a, h_b)
    r_f = search(f)
    assert r_b[0] == r_f[0] and all(np.array_equal(x, y) for x, y in zip(r_b[1], r_f[1]))
    assert r_a[0] != r_b[0]
    with pytest.raises(L.DxError, match="shape"):
        e.load_weights(54, 6, 128, 2, 1, True, L.pack_state_dict({k: v.numpy() for k, v in random_state_dict(324, 128, 2, 1, True, seed=1).items()})[0])
    e.close(); f.close()


def test_host_mirror_reuploads_changed_weights():
    """nnet.load_state_dict() between two batches of a reused PathFind: the second batch searches with the new network
    (the reference builds each PathFind around the current function, updaters/updater_rl.py:137-165)."""
    import deepxube_b200  # noqa: F401
    from deepxube_b200.factories.domain_factory import get_domain_from_arg
    from deepxube_b200.factories.pathfind_fns_factory import get_path_fns_nnet_par_dict
    from deepxube_b200.factories.pathfinding_factory import get_pathfind_from_arg
    domain, dname = get_domain_from_arg("cube3")
    sds = [_spread_net(324, 128, 1, 1, s) for s in (3, 4)]
    states, goals = domain.sample_problem_instances([20, 25, 30])

    def run(pf):
        insts = pf.make_instances(states, goals, None, True)
        pf.add_instances(insts)
        for _ in range(5):
            pf.step()
        return [(i.num_nodes_generated, i.lb, i.frontier_size()) for i in pf.instances]

    pfns, _ = get_path_fns_nnet_par_dict(domain, dname, ["heurv,resnet_fc.128H_1B_bn"])
    pfns.heurv.nnet.load_state_dict(sds[0])
    pf = get_pathfind_from_arg(domain, pfns, "graph.50B_0.7W")[0]
    r0 = run(pf)
    h0 = pfns.heurv(states, goals, [None] * 3)
    pfns.heurv.nnet.load_state_dict(sds[1])
    pf.reset()
    r1 = run(pf)
    h1 = pfns.heurv(states, goals, [None] * 3)
    pfns2, _ = get_path_fns_nnet_par_dict(domain, dname, ["heurv,resnet_fc.128H_1B_bn"])
    pfns2.heurv.nnet.load_state_dict(sds[1])
    pf2 = get_pathfind_from_arg(domain, pfns2, "graph.50B_0.7W")[0]
    assert run(pf2) == r1 and r0 != r1
    assert h1 == pfns2.heurv(states, goals, [None] * 3) and h0 != h1
    pf.close(); pf2.close()


# ---------------------------------------------------------------------------------------------------
def test_remove_instances_stops_them_and_slots_recycle():
    """PathFind.remove_instances / remove_finished_instances on the device: a removed instance is not searched any more
    (its counters freeze, it stops counting against max_pop_total), the others are unaffected (bit-equal to the oracle),
    and once none is left the engine's slots are recycled by the next add_instances."""
    import deepxube_b200  # noqa: F401
    from deepxube_b200.base.pathfind_fns import PFNsHeurV
    from deepxube_b200.base.pathfinding import get_path
    from deepxube_b200.factories.domain_factory import get_domain_from_arg
    from deepxube_b200.factories.pathfinding_factory import get_pathfind_from_arg
    domain, _ = get_domain_from_arg("npuzzle.3")
    dom = OracleDomain("npuzzle", 3)
    st, goals = _scrambles(dom, 3, 30, 9)
    states, goal_objs = domain.rows_to_states(st), [type(domain.sample_problem_instances([0])[1][0])(g) for g in goals]
    pf = get_pathfind_from_arg(domain, PFNsHeurV(heurv=lambda s, g, c: pseudo_v(domain.states_to_rows(s), None).tolist()), "graph.20B_1.0W",
                               )[0]
    pf._max_instances = 3
    insts = pf.make_instances(states, goal_objs, None, True)
    pf.add_instances(insts)
    orc = OracleSearch(dom, pseudo_v, False, 20, 1.0)
    oi = orc.add_instances(st, goals)
    for _ in range(3):
        pf.step(); orc.step()
    victim = insts[1]
    removed = pf.remove_instances(lambda i: i is victim)
    assert removed == [victim] and len(pf.instances) == 2
    frozen = (victim.itr, victim.num_nodes_generated)
    oi_live = [oi[0], oi[2]]
    for t in range(60):
        if all(i.finished() for i in pf.instances):
            break
        pf.step()
        for o in oi_live:
            if not o.finished():
                orc_step_one(orc, o)
    for inst, o in zip(pf.instances, oi_live):
        assert inst.finished() and o.finished()
        assert inst.num_nodes_generated == o.num_nodes_generated and inst.itr == o.itr and inst.path_cost() == o.path_cost()
        assert [a.action for a in get_path(inst.goal_node)[1]] == o.get_path()[1]
    sts = pf._engine.status()
    assert (sts[1].itr, sts[1].num_nodes_generated) == frozen and sts[1].finished        # never stepped again
    # remove the rest, then add again: the engine is recycled (slots 0.. are reused) and the old solutions stay readable
    done = pf.remove_finished_instances(10 ** 9)
    assert len(done) == 2 and pf.instances == []
    again = pf.make_instances(states, goal_objs, None, True)
    pf.add_instances(again)
    assert pf._engine.num_instances() == 3 and [i._slot for i in again] == [0, 1, 2]
    while not all(i.finished() for i in pf.instances):
        pf.step()
    assert [i.num_nodes_generated for i in again][0::2] == [i.num_nodes_generated for i in done]
    assert [a.action for a in get_path(done[0].goal_node)[1]] == oi[0].get_path()[1]     # materialised before the recycle
    pf.close()


def orc_step_one(orc, inst):
    """advance ONE oracle instance by an iteration (the oracle steps all unfinished instances of its list)"""
    keep = orc.instances
    orc.instances = [inst]
    orc.step()
    orc.instances = keep


def test_remove_frees_pop_capacity():
    L = _lib()
    dom = OracleDomain("cube3")
    st, goals = _scrambles(dom, 4, 30, 2)
    eng = L.Engine("cube3", 0, "graph_v", L.DX_HEUR_ZERO, max_instances=4, max_pop_total=200, max_nodes=400000)
    eng.add_instances(st[:2], goals[:2], 100, 1.0)
    eng.run(3)
    with pytest.raises(L.DxError, match="max_pop_total"):
        eng.add_instances(st[2:3], goals[2:3], 100, 1.0)
    eng.remove_instances([0])
    eng.add_instances(st[2:3], goals[2:3], 100, 1.0)            # fits now
    gen0 = eng.status()[0].num_nodes_generated
    eng.run(3)
    s = eng.status()
    assert s[0].num_nodes_generated == gen0 and s[0].finished and s[1].itr == 6 and s[2].itr == 3
    with pytest.raises(L.DxError, match="slot"):
        eng.remove_instances([7])
    eng.close()


# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("q", [False, True])
def test_negative_heuristic_values_order_like_the_reference_heap(q):
    """A host callable may return negative values (the reference's heap orders them numerically, graph_search.py:48-51;
    Node.tree_backup clips with max(h, 0) precisely because they occur): OPEN keys must sort negative f below positive f."""
    L = _lib()
    dom = OracleDomain("cube3")
    st, goals = _scrambles(dom, 2, 25, 3)
    base = make_pseudo_q(12) if q else pseudo_v
    fn = (lambda s, g: base(s, g) - 32.0)                      # roughly half of the values below zero
    vals = fn(st, goals)
    assert (np.asarray(vals) < 0).any()
    orc = OracleSearch(dom, fn, q, 60, 0.3)
    oi = orc.add_instances(st, goals)
    eng = L.Engine("cube3", 0, "graph_q" if q else "graph_v", L.DX_HEUR_HOST, max_instances=2, max_pop_total=120, max_nodes=300000)
    eng.add_instances(st, goals, 60, 0.3, root_vals=vals)
    saw_negative_lb = False
    for t in range(8):
        orc.step()
        n = eng.step_begin()
        ps, pg = eng.get_pending(n)
        eng.step_end(fn(ps, pg) if n else np.zeros((0, 12) if q else (0,)))
        _compare_iteration(t, eng, oi)
        saw_negative_lb = saw_negative_lb or any(o.lb < 0 for o in oi) or any(s.lb < 0 for s in eng.status())
    eng.close()


def test_nan_heuristic_is_rejected_by_the_host_mirror():
    import deepxube_b200  # noqa: F401
    from deepxube_b200.base.pathfind_fns import PFNsHeurV
    from deepxube_b200.factories.domain_factory import get_domain_from_arg
    from deepxube_b200.factories.pathfinding_factory import get_pathfind_from_arg
    domain, _ = get_domain_from_arg("cube3")
    states, goals = domain.sample_problem_instances([5])
    pf = get_pathfind_from_arg(domain, PFNsHeurV(heurv=lambda s, g, c: [float("nan")] * len(s)), "graph.5B_1.0W")[0]
    with pytest.raises(ValueError, match="NaN"):
        pf.add_instances(pf.make_instances(states, goals, None, True))
    pf.close()


def test_one_hot_inputs_out_of_range_are_rejected():
    """F.one_hot raises in the reference when a value is >= the depth (pytorch_models.py:15-24); the C ABI refuses such rows
    instead of clamping them on the device."""
    L = _lib()
    eng = L.Engine("cube3", 0, "graph_v", L.DX_HEUR_NET_FP32, max_instances=2, max_pop_total=16, max_nodes=1024)
    blob, _ = L.pack_state_dict({k: v.numpy() for k, v in random_state_dict(324, 128, 1, 1, True, seed=1).items()})
    eng.load_weights(54, 6, 128, 1, 1, True, blob)
    bad = np.zeros((2, 54), np.uint8)
    bad[1, 17] = 6
    with pytest.raises(L.DxError, match="one-hot depth"):
        eng.heur_eval(bad, None, 1)
    with pytest.raises(L.DxError, match="one-hot depth"):
        eng.add_instances(bad, bad, 4, 1.0)
    eng.close()


# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("device_resident", [True, False])
def test_hda_capacity_error_reaches_every_rank(device_resident):
    """An exchange-segment / receive-capacity overflow on ONE rank must stop ALL ranks in the same iteration (the flag travels
    in the all-gathered scalars) -- otherwise the healthy ranks of a multi-process run wait forever in the next collective."""
    from deepxube_b200 import _lib as L
    from deepxube_b200.hda import HdaSearch
    dom = OracleDomain("cube3")
    st, goals = _scrambles(dom, 1, 60, 4)
    h = HdaSearch("cube3", 0, 4, batch_size=4000, weight=0.0, max_nodes_per_rank=1 << 20, capacity_slack=1.0,
                  device_resident=device_resident)
    if device_resident:
        h.seg_records = 40                      # far below the ~3000 records a rank sends to each owner
        h.seg_bytes = 16 + h.seg_records * h.rec
    else:
        h.cap_records = 500
    with pytest.raises(L.DxError):
        h.solve(st[0], goals[0], max_itrs=12)
    if device_resident:
        flags = [int(s.cpu()[6].item()) for s in h.summary]
        assert all(f != 0 for f in flags), flags          # every rank's summary carries the error, not only the failing one
    h.close()


# ---------------------------------------------------------------------------------------------------
# Young-run OPEN of small-batch A* engines (csrc/k_open.cu k_small_tail_v)
def _run_vs_oracle(L, dom, dname, dim, st, goals, batch, weight, iters, fn, add_later=None, max_pop=None, is_q=False, max_nodes=600000,
                   remove_at=None):
    """HOST-heuristic engine against the oracle, compared after every iteration (popped nodes, lb / ub / OPEN size / counters).
    remove_at = (iteration, [slots]): those instances are removed (dx_remove_instances) before that iteration."""
    orc = OracleSearch(dom, fn, is_q, batch, weight)
    oi = orc.add_instances(st, goals)
    adds = [] if not add_later else ([add_later] if not isinstance(add_later, list) else add_later)
    n_all = len(st) + sum(len(a[1]) for a in adds)
    eng = L.Engine(dname, dim, "graph_q" if is_q else "graph_v", L.DX_HEUR_HOST, max_instances=n_all, max_pop_total=max_pop or n_all * batch,
                   max_nodes=max_nodes)
    eng.add_instances(st, goals, batch, weight, root_vals=fn(st, goals))
    gone = set()
    for t in range(iters):
        for a in adds:
            if t == a[0]:
                st2, gl2, b2 = a[1], a[2], a[3]
                oi += orc.add_instances(st2, gl2, batch_size=b2)
                eng.add_instances(st2, gl2, b2, weight, root_vals=fn(st2, gl2))
        if remove_at and t == remove_at[0]:
            eng.remove_instances(list(remove_at[1]))
            gone.update(remove_at[1])
            orc.instances = [o for j, o in enumerate(oi) if j not in gone]     # (the oracle steps its `instances` list)
        orc.step()
        n = eng.step_begin()
        ps, pg = eng.get_pending(n)
        eng.step_end(fn(ps, pg) if n else np.zeros((0, dom.num_actions) if is_q else (0,)))
        for j, (o, s) in enumerate(zip(oi, eng.status())):
            if j in gone:
                continue
            assert s.lb == o.lb and s.ub == o.ub and s.num_nodes_generated == o.num_nodes_generated, (t, j, s.lb, o.lb)
            assert bool(s.finished) == o.finished(), (t, j)
            if not s.finished:
                assert s.open_size == len(o.open), (t, j, s.open_size, len(o.open))
                cs, cg, _ = eng.curr_nodes(j, len(o.nodes_curr) + 8)
                assert np.array_equal(cs, o.store.state[o.nodes_curr]), (t, j)
    return eng, oi


@pytest.mark.parametrize("dname,dim,n,batch,weight", [("cube3", 0, 1, 100, 0.2), ("cube3", 0, 7, 20, 0.6), ("npuzzle", 4, 24, 10, 0.8),
                                                      ("grid", 7, 60, 3, 1.0)])
def test_young_run_open_matches_the_oracle_over_many_compactions(dname, dim, n, batch, weight):
    """40 iterations of small-batch searches (several compactions of the young runs into the main runs, instances finishing at
    different iterations, 1 to 60 instances in one block): every iteration's popped nodes, OPEN sizes, lb / ub and counters are
    the oracle's.  The pseudo-heuristic has many exact ties, so the (f, push order) tie rule is exercised across the run boundaries."""
    L = _lib()
    dom = OracleDomain(dname, dim)
    if dname == "grid":
        rng = np.random.default_rng(4)
        st = rng.integers(0, dim, size=(n, 2)).astype(np.uint8)
        goals = rng.integers(0, dim, size=(n, 2)).astype(np.uint8)
    else:
        st, goals = _scrambles(dom, n, 40, 9)
    eng, oi = _run_vs_oracle(L, dom, dname, dim, st, goals, batch, weight, 40, pseudo_v)
    assert eng.counters().kernel_launches > 0
    eng.close()


@pytest.mark.parametrize("every", ["1", "3", "0"])
@pytest.mark.parametrize("dname,dim,is_q,n,batch,weight,iters", [
    ("cube3", 0, False, 3, 300, 0.6, 24),      # 3 x 300 x 12 = 10 800 pushed entries per iteration: the grid-wide kernels
    ("cube3", 0, True, 4, 700, 0.5, 22),       # Q*: edges in OPEN, 12 pushed per kept node
    ("npuzzle", 4, False, 40, 30, 0.8, 30),    # many instances, ragged runs, some finish early
    ("grid", 7, True, 200, 5, 1.0, 14),        # tiny state space: runs empty out, instances finish at different iterations
])
def test_two_run_open_of_big_batches_matches_the_oracle(monkeypatch, every, dname, dim, is_q, n, batch, weight, iters):
    """Big batches keep OPEN as two runs as well (dx_launch_young_iter: young <- merge(young, batch), pop split, popped list <- merge
    of the two prefixes; compaction every DXB200_YOUNG_EVERY iterations, or by the cost model when 0): every iteration's popped
    nodes, OPEN sizes, lb / ub and counters are the oracle's, with the tie-heavy pseudo-heuristic."""
    L = _lib()
    monkeypatch.setenv("DXB200_YOUNG_EVERY", every)
    monkeypatch.setenv("DXB200_YOUNG_BIG", "2")             # (graph_v engines use it only on request: see engine.cu)
    dom = OracleDomain(dname, dim)
    if dname == "grid":
        rng = np.random.default_rng(4)
        st = rng.integers(0, dim, size=(n, 2)).astype(np.uint8)
        goals = rng.integers(0, dim, size=(n, 2)).astype(np.uint8)
    else:
        st, goals = _scrambles(dom, n, 40, 9)
    fn = make_pseudo_q(dom.num_actions) if is_q else pseudo_v
    eng, oi = _run_vs_oracle(L, dom, dname, dim, st, goals, batch, weight, iters, fn, is_q=is_q, max_nodes=3000000)
    eng.close()


@pytest.mark.parametrize("levels,scale_log2,is_q", [
    (1 << 24, -42, False),     # ~3600 distinct keys per iteration inside a few 2^-30 windows: small unsorted groups everywhere
    (4096, -40, False),        # every pushed key of an iteration in ONE window, ~4096 distinct values: long unsorted groups (rank sort)
    (5, -40, True),            # five distinct values: long groups that are mostly exact ties, out of order across the values
    (1 << 20, -20, True),      # ordinary spread: single-entry groups
])
@pytest.mark.parametrize("sticky", ["0", "1"])
def test_sort_key_cut_keeps_the_full_key_order(monkeypatch, sticky, levels, scale_log2, is_q):
    """The digit passes of the batch sort leave out the low 23 bits of the f-key (k_open.cu digit_of / k_sort_fixup): heuristics
    whose values differ only far below that cut -- small unsorted groups, long unsorted groups, long groups of exact ties -- must
    still pop in exactly the oracle's (f, push order) order, every iteration.  A sort that meets more than 1024 runs inside one group
    makes the engine give up the cut (sticky = 1, the default); sticky = 0 keeps it, so the merge rounds run every iteration."""
    L = _lib()
    monkeypatch.setenv("DXB200_SORT_CUT_STICKY", sticky)
    dom = OracleDomain("cube3")
    st, goals = _scrambles(dom, 3, 40, 21)
    scale = 2.0 ** scale_log2

    def fn_v(states, goals_=None):
        return (np.floor(pseudo_v_fine(states) * 262144.0) % levels) * scale          # exact in float32: < 2^24 levels

    def fn_q(states, goals_=None):
        return (np.floor(make_pseudo_q_fine(12)(states) * 262144.0) % levels) * scale

    eng, oi = _run_vs_oracle(L, dom, "cube3", 0, st, goals, 300 if not is_q else 500, 0.6, 12, fn_q if is_q else fn_v, is_q=is_q,
                             max_nodes=3000000)
    eng.close()


def test_two_run_open_survives_regime_changes(monkeypatch):
    """small (one-block tail) -> big (grid-wide merges) when instances are added, big -> small when instances are removed, and a
    Q* engine that drops out of the big regime: the two runs carry over / are flushed and the searches go on unchanged."""
    L = _lib()
    monkeypatch.setenv("DXB200_YOUNG_EVERY", "2")
    monkeypatch.setenv("DXB200_YOUNG_BIG", "2")
    dom = OracleDomain("cube3")
    st, goals = _scrambles(dom, 5, 35, 12)
    # A*: 2 x 50 (small) -> + 2 x 400 (big) at iteration 6 -> the two big ones removed at iteration 13 -> + 1 x 20 at iteration 16:
    # the regime is decided when instances are added, so that is where the engine goes back to the one-block tail
    st6, goals6 = _scrambles(dom, 6, 35, 12)
    eng, oi = _run_vs_oracle(L, dom, "cube3", 0, st6[:2], goals6[:2], 50, 0.5, 26, pseudo_v,
                             add_later=[(6, st6[2:4], goals6[2:4], 400), (16, st6[4:5], goals6[4:5], 20)],
                             max_pop=2 * 50 + 2 * 400 + 20, remove_at=(13, [2, 3]), max_nodes=3000000)
    eng.close()
    # Q*: 3 x 600 (big) -> two removed at iteration 8: 600 x 12 pushes still big; then the engine is reused after a reset
    eng, oi = _run_vs_oracle(L, dom, "cube3", 0, st[:3], goals[:3], 600, 0.5, 16, make_pseudo_q(12), is_q=True, remove_at=(8, [0, 2]),
                             max_nodes=3000000)
    eng.close()
    # Q*: 2 x 150 = 3600 pushes (big) -> one removed: 1800 pushes, the small Q* kernels have one run: flush
    eng, oi = _run_vs_oracle(L, dom, "cube3", 0, st[:2], goals[:2], 150, 0.5, 16, make_pseudo_q(12), is_q=True, remove_at=(7, [1]),
                             max_nodes=3000000)
    eng.close()


def test_young_run_is_flushed_when_the_batch_outgrows_the_small_regime():
    """An engine that starts in the small regime (young runs active) and then receives instances whose batches exceed it: the
    young runs are merged back into one run and the search continues on the big-batch kernels with the same results."""
    L = _lib()
    dom = OracleDomain("cube3")
    st, goals = _scrambles(dom, 3, 35, 12)
    eng, oi = _run_vs_oracle(L, dom, "cube3", 0, st[:2], goals[:2], 50, 0.5, 14, pseudo_v, add_later=(7, st[2:], goals[2:], 400),
                             max_pop=2 * 50 + 400)
    eng.close()


def test_bursts_stop_in_the_exact_iteration_and_the_engine_continues_after_a_finished_search(monkeypatch):
    """dx_run launches bursts of iterations per host poll in the small-batch regime; the device raises a DONE bit in the iteration
    that finishes the last instance, so the rest of the burst must be a no-op: same iteration counts / nodes / costs as polling every
    iteration, max_iters honoured when it is not a multiple of the burst, and an engine whose search ended mid-burst (young runs,
    compaction schedule and buffer parities included) keeps working when another instance is added without a reset."""
    L = _lib()
    k = gu.load("kat_cube3")
    sd = gu.state_dict_np("cube3v")
    blob, info = L.pack_state_dict(sd)
    results = {}
    for poll in ("1", "4", "7"):
        monkeypatch.setenv("DXB200_POLL_EVERY", poll)
        eng = L.Engine("cube3", 0, "graph_v", L.DX_HEUR_NET_FP32, max_instances=3, max_pop_total=300, max_nodes=900000)
        eng.load_weights(54, 6, info["res_dim"], info["num_blocks"], 1, True, blob)
        eng.add_instances(k["hard_states"][:1], k["hard_goals"][:1], 100, 0.1)
        assert eng.run(10) == 10 and eng.status()[0].itr == 10              # (10 is neither a multiple of 4 nor of 7)
        done = 10 + eng.run(100000)
        s0 = eng.status()[0]
        assert s0.finished and s0.itr == done
        eng.add_instances(k["hard_states"][1:3], k["hard_goals"][1:3], 100, 0.1)      # no reset: slots 1 and 2 join a finished slot 0
        more = eng.run(100000)
        st = eng.status()
        results[poll] = (done, more, [(s.itr, s.num_nodes_generated, s.ub, bool(s.finished)) for s in st],
                         [eng.get_path(j)[0] for j in range(3)])
        eng.close()
    assert results["1"] == results["4"] == results["7"]
    done, more, stat, paths = results["4"]
    # instance 0 alone is the KAT search: the reference's iteration count / node count / cost (tests/golden/kat_cube3.npz)
    assert stat[0][1] == int(k["hard_nodes"][0]) and stat[0][2] == float(k["hard_path_costs"][0])
    assert all(s[3] for s in stat) and all(p is not None for p in paths)


@pytest.mark.parametrize("is_q", [False, True])
def test_big_batch_bursts_stop_in_the_exact_iteration(monkeypatch, is_q):
    """dx_run polls the host every DXB200_POLL_EVERY_BIG iterations for big batches too: identical iteration counts, node counts, costs
    and paths for 1 / 3 / 4 iterations per poll, max_iters honoured mid-burst, and the engine goes on after a search that ended inside
    a burst (two-run OPEN compaction schedule and buffer parities included)."""
    L = _lib()
    k = gu.load("kat_cube3")
    sd = gu.state_dict_np("cube3q" if is_q else "cube3v")
    blob, info = L.pack_state_dict(sd)
    results = {}
    if is_q:                   # (the golden Q network is barely trained, outputs 0 .. 0.1: breadth-first in effect, so 4-move scrambles)
        k = dict(zip(("hard_states", "hard_goals"), _scrambles(OracleDomain("cube3"), 4, 4, 33)))
    for poll in ("1", "3", "4"):
        monkeypatch.setenv("DXB200_POLL_EVERY_BIG", poll)
        eng = L.Engine("cube3", 0, "graph_q" if is_q else "graph_v", L.DX_HEUR_NET_FP32, max_instances=4, max_pop_total=1600, max_nodes=6000000)
        eng.load_weights(54, 6, info["res_dim"], info["num_blocks"], 12 if is_q else 1, True, blob)
        w = 0.05 if is_q else 0.2
        eng.add_instances(k["hard_states"][:2], k["hard_goals"][:2], 400, w)            # 2 x 400 x 12 pushes: the grid-wide kernels
        assert eng.run(5) == 5 and all(s.itr == 5 for s in eng.status())                  # (5 is not a multiple of 3 or 4)
        done = 5 + eng.run(100000)
        assert all(s.finished for s in eng.status()) and max(s.itr for s in eng.status()) == done
        eng.add_instances(k["hard_states"][2:4], k["hard_goals"][2:4], 400, w)           # no reset: two more join the finished ones
        more = eng.run(100000)
        st = eng.status()
        results[poll] = (done, more, [(s.itr, s.num_nodes_generated, s.ub, s.lb, bool(s.finished)) for s in st],
                         [eng.get_path(j)[0] for j in range(4)])
        eng.close()
    assert results["1"] == results["3"] == results["4"]
    assert all(s[4] for s in results["1"][2]) and all(p is not None for p in results["1"][3])


def test_instances_added_right_after_another_finished_keep_the_big_kernels():
    """Found by tests/fuzz_vs_oracle.py: an instance that has just finished still has its last pops in the current popped list (they
    stay readable); instances added at that moment made the live batches fit the small-batch regime (303 x 4 pushes) while the next
    iteration's CLOSED block still saw 603 x 4 children -- "max_pop_total" from its own sanity check.  The regime bound now covers the
    current popped list."""
    L = _lib()
    dom = OracleDomain("npuzzle", 4)
    easy, goals = _scrambles(dom, 1, 3, 5)
    hard, _ = _scrambles(dom, 4, 60, 6)
    st = np.concatenate([easy, hard[:1]])
    gl = np.tile(goals[0], (2, 1))
    orc = OracleSearch(dom, pseudo_v_fine, False, 300, 1.0)
    oi = orc.add_instances(st, gl)
    eng = L.Engine("npuzzle", 4, "graph_v", L.DX_HEUR_HOST, max_instances=5, max_pop_total=603, max_nodes=2000000)
    eng.add_instances(st, gl, 300, 1.0, root_vals=pseudo_v_fine(st, gl))
    added, after = False, 0
    for t in range(40):
        if not added and oi[0].finished() and not oi[1].finished():
            oi += orc.add_instances(hard[1:4], np.tile(goals[0], (3, 1)), batch_size=1)
            eng.add_instances(hard[1:4], np.tile(goals[0], (3, 1)), 1, 1.0, root_vals=pseudo_v_fine(hard[1:4]))
            added = True
        orc.step()
        k = eng.step_begin()
        ps, pg = eng.get_pending(k)
        eng.step_end(pseudo_v_fine(ps, pg) if k else np.zeros((0,)))
        for j, (o, s) in enumerate(zip(oi, eng.status())):
            assert s.lb == o.lb and s.ub == o.ub and s.num_nodes_generated == o.num_nodes_generated and bool(s.finished) == o.finished(), (t, j)
            if not s.finished:
                assert s.open_size == len(o.open), (t, j)
        after += added
        if after >= 6:
            break
    assert added and after >= 6
    eng.close()
