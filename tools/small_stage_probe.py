"""GPU probe: per-stage device time of small-batch iterations (tutorial KAT shape: cube3, graph.100B_0.1W, zero heuristic);
also the command the ncu list profiles/r1m_small_iteration_ncu.csv was taken from."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepxube_b200 import _lib as L
g = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
k = np.load(os.path.join(g, "kat_cube3.npz"))
eng = L.Engine("cube3", 0, "graph_v", L.DX_HEUR_ZERO, max_instances=1, max_pop_total=100, max_nodes=600000)
eng.add_instances(k["hard_states"][:1], k["hard_goals"][:1], 100, 0.1)
eng.run(50)
eng.reset()
eng.set_profiling(1)
eng.add_instances(k["hard_states"][:1], k["hard_goals"][:1], 100, 0.1)
it = eng.run(200)
ms = eng.stage_ms()
print("iters", it, {kk: round(v * 1000 / it, 1) for kk, v in ms.items()}, "us per iteration per stage")
