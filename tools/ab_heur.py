"""GPU A/B of the bf16 network under two settings of one environment switch, interleaved in ONE process
(the B200s of this pool run power-capped: separate runs drift by several percent).
usage: ab_heur.py VAR valA valB [n_rows] [rounds]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepxube_b200 import _lib as L
from oracle.nnet_torch import random_state_dict
from oracle.domains_np import OracleDomain

var, va, vb = sys.argv[1], sys.argv[2], sys.argv[3]
N = int(sys.argv[4]) if len(sys.argv) > 4 else 600000
ROUNDS = int(sys.argv[5]) if len(sys.argv) > 5 else 12
sd = random_state_dict(324, 1000, 4, 1, True, seed=0)
blob, info = L.pack_state_dict({k: v.numpy() for k, v in sd.items()})
dom = OracleDomain("cube3")
rng = np.random.default_rng(0)
goals = np.tile(dom.canonical_goal(), (8, 1))
st = goals.copy()
for _ in range(100):
    st = dom.next_state(st, rng.integers(0, 12, size=8))


def mk(val):
    os.environ[var] = val
    e = L.Engine("cube3", 0, "graph_v", L.DX_HEUR_NET_BF16, max_instances=8, max_pop_total=80000, max_nodes=2500000)
    e.load_weights(54, 6, 1000, 4, 1, True, blob)
    e.add_instances(st, goals, 10000, 0.6)     # real states in the node store (bench_heur evaluates stored rows)
    e.run(5)
    return e


ea, eb = mk(va), mk(vb)
n = min(N, ea.counters().nodes_stored, eb.counters().nodes_stored)
for _ in range(3):
    ea.bench_heur(n, 3); eb.bench_heur(n, 3)
ta, tb = [], []
for r in range(ROUNDS):
    ta.append(ea.bench_heur(n, 3) / 3)
    tb.append(eb.bench_heur(n, 3) / 3)
ta, tb = np.array(ta), np.array(tb)
print(f"{var}={va}: median {np.median(ta):.3f} ms  min {ta.min():.3f}   |   {var}={vb}: median {np.median(tb):.3f} ms  min {tb.min():.3f}"
      f"   |   B/A median ratio {np.median(tb) / np.median(ta):.4f}  (n={n} rows, {ROUNDS} interleaved rounds)")
