"""Randomised engine-vs-oracle runs (GPU): HOST-heuristic engines against oracle/search.py, compared after EVERY iteration (popped
nodes, lb / ub / OPEN size / counters), over random domains, A* / Q*, instance counts, batch sizes, weights, tie-heavy / fine / sub-cut
heuristics, instances added and removed mid-search, and the OPEN / sort switches (DXB200_YOUNG_BIG, DXB200_YOUNG_EVERY, DXB200_SORT_CUT).
usage: fuzz_vs_oracle.py [seconds=120] [seed=0] [first_case=1]     (test infrastructure: the oracle is the checker; cases before
first_case are drawn but not run, to replay a failure)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepxube_b200 import _lib as L
from oracle.domains_np import OracleDomain
from oracle.pseudo_heur import make_pseudo_q, make_pseudo_q_fine, pseudo_v, pseudo_v_fine
from oracle.search import OracleSearch

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
first_case = int(sys.argv[3]) if len(sys.argv) > 3 else 1


def instances(dom, n):
    if dom.name == "grid":
        return (rng.integers(0, dom.dim, size=(n, 2)).astype(np.uint8), rng.integers(0, dom.dim, size=(n, 2)).astype(np.uint8))
    goals = np.tile(dom.canonical_goal(), (n, 1))
    st = goals.copy()
    depth = rng.integers(0, 40, size=n)
    for t in range(40):
        mv = dom.next_state(st, rng.integers(0, dom.num_actions, size=n))
        st = np.where((depth > t)[:, None], mv, st)
    return st, goals


def heuristic(kind, is_q, A):
    if kind == "ties":
        return make_pseudo_q(A) if is_q else pseudo_v
    fine = make_pseudo_q_fine(A) if is_q else pseudo_v_fine
    if kind == "fine":
        return fine
    levels, scale = (4096, 2.0 ** -40) if kind == "subcut" else (1 << 22, 2.0 ** -36)
    return lambda s, g=None: (np.floor(fine(s) * 262144.0) % levels) * scale


t_end, case, checked = time.time() + budget, 0, 0
while time.time() < t_end:
    case += 1
    name, dim = [("cube3", 0), ("npuzzle", 3), ("npuzzle", 4), ("grid", 5), ("grid", 7), ("lightsout", 3)][rng.integers(0, 6)]
    dom = OracleDomain(name, dim)
    A = dom.num_actions
    is_q = bool(rng.integers(0, 2))
    n = int(rng.choice([1, 2, 5, 17, 60]))
    batch = int(rng.choice([1, 7, 60, 300, 800]))
    if n * batch * A > 40000:
        batch = max(1, 40000 // (n * A))
    weight = float(rng.choice([0.0, 0.2, 0.6, 1.0]))
    iters = int(rng.integers(6, 22))
    kind = str(rng.choice(["ties", "fine", "subcut", "subcut2", "zero-run"]))      # zero-run: device zero heuristic through dx_run (graphs, bursts)
    env = {"DXB200_YOUNG_BIG": str(rng.integers(0, 3)), "DXB200_YOUNG_EVERY": str(rng.choice([0, 1, 2, 5])),
           "DXB200_SORT_CUT": str(rng.choice([0, 23, 30])), "DXB200_SORT_CUT_STICKY": str(rng.integers(0, 2)),
           "DXB200_POLL_EVERY": str(rng.choice([1, 3, 4])), "DXB200_POLL_EVERY_BIG": str(rng.choice([1, 2, 4])),
           "DXB200_NO_GRAPHS": str(rng.choice([0, 0, 0, 1]))}
    for kv in os.environ.get("FUZZ_OVERRIDE", "").split():         # (replaying a failure with one switch changed)
        env[kv.split("=")[0]] = kv.split("=")[1]
    os.environ.update(env)
    run_mode = kind == "zero-run"
    backups = (not run_mode) and bool(rng.integers(0, 4) == 0)      # evaluate-all engine with the training-label log, drained at the end
    label_mode = str(rng.choice(["bellman", "ub_solns", "lhbl"]))
    chunks = [int(x) for x in rng.integers(1, 6, size=32)]          # iterations per dx_run call
    fn = None if run_mode else heuristic(kind, is_q, A)
    st, goals = instances(dom, n)
    n_late = int(rng.integers(0, 4))
    t_add, t_rem = int(rng.integers(1, iters)), int(rng.integers(1, iters))
    st2, goals2 = instances(dom, max(n_late, 1))
    b2 = int(rng.choice([1, batch, min(4 * batch, 1000)]))
    rem = sorted(set(rng.integers(0, n, size=int(rng.integers(0, 3))).tolist())) if n > 1 else []
    if backups:
        rem = []                                                     # (the oracle's log indexes instances by position)
    desc = (f"case {case}: {name}.{dim} {'Q*' if is_q else 'A*'} n={n} B={batch} w={weight} iters={iters} heur={kind} +{n_late}x{b2}@{t_add} -{rem}@{t_rem} "
            f"{'backups:' + label_mode + ' ' if backups else ''}{env}")
    if case < first_case:
        continue
    orc = OracleSearch(dom, fn, is_q, batch, weight)
    oi = orc.add_instances(st, goals)
    eng = L.Engine(name, dim, "graph_q" if is_q else "graph_v", L.DX_HEUR_ZERO if run_mode else L.DX_HEUR_HOST, max_instances=n + n_late,
                   max_pop_total=n * batch + n_late * b2, max_nodes=4000000,
                   max_backups=2 * (n * batch + n_late * b2) * (iters + 2) if backups else 0)
    eng.add_instances(st, goals, batch, weight, root_vals=None if run_mode else fn(st, goals))
    gone = set()
    t = -1
    try:
        t = 0
        while t < iters:
            if n_late and t >= t_add and len(oi) == n:
                oi += orc.add_instances(st2[:n_late], goals2[:n_late], batch_size=b2)
                eng.add_instances(st2[:n_late], goals2[:n_late], b2, weight, root_vals=None if run_mode else fn(st2[:n_late], goals2[:n_late]))
            if rem and t >= t_rem and not gone:
                eng.remove_instances(rem)
                gone.update(rem)
                orc.instances = [o for j, o in enumerate(oi) if j not in gone]
            if run_mode:                             # the engine stops in the iteration that finishes the last instance, like the loop below
                want = chunks[t % 32]
                did = 0
                while did < want and not all(o.finished() for o in orc.instances):
                    orc.step()
                    did += 1
                got = eng.run(want)
                assert got == did, (t, "iterations run", got, did)
                t += max(did, 1)
            else:
                orc.step()
                k = eng.step_begin()
                ps, pg = eng.get_pending(k)
                eng.step_end(fn(ps, pg) if k else np.zeros((0, A) if is_q else (0,)))
                t += 1
            for j, (o, s) in enumerate(zip(oi, eng.status())):
                if j in gone:
                    continue
                assert s.lb == o.lb and s.ub == o.ub and s.num_nodes_generated == o.num_nodes_generated, (t, j, "lb/ub/gen", s.lb, o.lb, s.ub, o.ub)
                assert bool(s.finished) == o.finished(), (t, j, "finished")
                if not s.finished:
                    assert s.open_size == len(o.open), (t, j, "open", s.open_size, len(o.open))
                    cs, _, _ = eng.curr_nodes(j, len(o.nodes_curr) + 8)
                    assert np.array_equal(cs, o.store.state[o.nodes_curr]), (t, j, "popped")
                checked += 1
        if backups:
            deep, depths = eng.backup_deepest()
            want = [o.deepest_with_children() for o in oi]
            assert np.array_equal(depths, np.array([d for _, d in want], np.int32)), "deepest depths"
            assert np.array_equal(deep, np.stack([s_ for s_, _ in want])), "deepest states"
            got = eng.drain_backups(label_mode)
            w_st = np.concatenate([x[0] for x in orc.backup_log]); w_idx = np.concatenate([x[2] for x in orc.backup_log])
            assert np.array_equal(got[0], w_st) and np.array_equal(got[-1], w_idx), "backup records"
            if is_q:
                assert np.array_equal(got[1], np.concatenate([x[3] for x in orc.backup_log])), "backup actions"
            assert np.array_equal(got[-2], (orc.labels_q if is_q else orc.labels_v)(label_mode)), "labels " + label_mode
    except Exception as ex:
        print("FAIL", desc, "at iteration", t, "->", repr(ex)[:300], flush=True)
        sys.exit(1)
    finally:
        eng.close()
    print("ok  ", desc, flush=True)
print(f"{case} cases, {checked} (instance, iteration) comparisons: all equal")
