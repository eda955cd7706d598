"""GPU experiment: K engines on K host threads (own streams), I/K instances each, vs one engine with I instances."""
import sys, os, time, threading
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepxube_b200 import _lib as L
from oracle.nnet_torch import random_state_dict
from oracle.domains_np import OracleDomain

I = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ITERS = int(sys.argv[2]) if len(sys.argv) > 2 else 20
B = 10000
dom = OracleDomain("cube3")
rng = np.random.default_rng(0)
goals = np.tile(dom.canonical_goal(), (I, 1))
st = goals.copy()
for _ in range(100):
    st = dom.next_state(st, rng.integers(0, 12, size=I))
sd = random_state_dict(324, 1000, 4, 1, True, seed=0)
blob, info = L.pack_state_dict({k: v.numpy() for k, v in sd.items()})

def mk(n):
    e = L.Engine("cube3", 0, "graph_v", L.DX_HEUR_NET_BF16, max_instances=n, max_pop_total=n * B, max_nodes=n * B * 12 * (ITERS + 2))
    e.load_weights(54, 6, 1000, 4, 1, True, blob)
    return e

for K in (1, 2, 4):
    if I % K: continue
    n = I // K
    engs = [mk(n) for _ in range(K)]
    gens = [0] * K
    def work(k):
        e = engs[k]
        e.reset()
        e.add_instances(st[k * n:(k + 1) * n], goals[k * n:(k + 1) * n], B, 0.6)
        e.run(ITERS)
        gens[k] = sum(s.num_nodes_generated for s in e.status())
    for rep in range(4):
        ths = [threading.Thread(target=work, args=(k,)) for k in range(K)]
        t0 = time.perf_counter()
        for t in ths: t.start()
        for t in ths: t.join()
        dt = time.perf_counter() - t0
        print(f"K={K} rep{rep}: gen={sum(gens)} time={dt*1e3:.1f} ms nodes/s={sum(gens)/dt:.3e}", flush=True)
    for e in engs: e.close()
