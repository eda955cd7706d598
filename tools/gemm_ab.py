"""Repeatable network-throughput measurement (min / median over several bench_heur calls)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepxube_b200 import _lib as L
from oracle.nnet_torch import random_state_dict
sd = random_state_dict(324, 1000, 4, 1, True, seed=0)
blob, info = L.pack_state_dict({k: v.numpy() for k, v in sd.items()})
eng = L.Engine("cube3", 0, "graph_v", L.DX_HEUR_NET_BF16, max_instances=8, max_pop_total=80000, max_nodes=1_000_000)
eng.load_weights(54, 6, 1000, 4, 1, True, blob)
for n in (100000, 617000, 960000):
    eng.bench_heur(n, 3)
    ms = sorted(eng.bench_heur(n, 4) / 4 for _ in range(7))
    fl = 2.0 * (324 * 1000 + 8 * 1000 * 1000 + 1000) * n
    print(f"n={n}: min {ms[0]:.3f} ms ({fl/ms[0]/1e9:.0f} TFLOP/s)  median {ms[3]:.3f} ms ({fl/ms[3]/1e9:.0f} TFLOP/s)", flush=True)
