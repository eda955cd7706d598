"""Do two engines on two streams overlap usefully?  One engine x I instances vs G engines x I/G instances driven from G host
threads (ctypes releases the GIL in dx_run), same total work: cube3 batch weighted A*, B = 10000, bf16 network.
usage: overlap_probe.py [I=8] [ITERS=20] [G=2]"""
import sys, os, time, threading
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepxube_b200 import _lib as L
from oracle.nnet_torch import random_state_dict
from oracle.domains_np import OracleDomain

I = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ITERS = int(sys.argv[2]) if len(sys.argv) > 2 else 20
G = int(sys.argv[3]) if len(sys.argv) > 3 else 2
B = 10000
dom = OracleDomain("cube3")
rng = np.random.default_rng(0)
goals = np.tile(dom.canonical_goal(), (I, 1))
st = goals.copy()
for _ in range(100):
    st = dom.next_state(st, rng.integers(0, 12, size=I))
sd = random_state_dict(324, 1000, 4, 1, True, seed=0)
blob, info = L.pack_state_dict({k: v.numpy() for k, v in sd.items()})


def make(n):
    e = L.Engine("cube3", 0, "graph_v", L.DX_HEUR_NET_BF16, max_instances=n, max_pop_total=n * B, max_nodes=n * B * 12 * (ITERS + 2))
    e.load_weights(54, 6, 1000, 4, 1, True, blob)
    return e


def run_group(engs, parts):
    for e, (a, b) in zip(engs, parts):
        e.reset()
        e.add_instances(st[a:b], goals[a:b], B, 0.6)
    t0 = time.time()
    ths = [threading.Thread(target=e.run, args=(ITERS,)) for e in engs]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.time() - t0
    gen = sum(s.num_nodes_generated for e in engs for s in e.status())
    return gen, dt


one = [make(I)]
many = [make(I // G) for _ in range(G)]
parts1 = [(0, I)]
partsG = [(k * (I // G), (k + 1) * (I // G)) for k in range(G)]
for rep in range(4):
    g1, t1 = run_group(one, parts1)
    gG, tG = run_group(many, partsG)
    print(f"rep{rep}: 1 engine x {I}: {g1 / t1:.3e} nodes/s ({t1 * 1e3:.1f} ms)   {G} engines x {I // G}: {gG / tG:.3e} nodes/s ({tG * 1e3:.1f} ms)   "
          f"ratio {(gG / tG) / (g1 / t1):.3f}  (gen {g1} vs {gG})", flush=True)
