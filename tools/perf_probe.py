"""GPU perf probe: stage breakdown of cube3 batch weighted A* (C2 shape) and raw network throughput."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepxube_b200 import _lib as L
from oracle.nnet_torch import random_state_dict
from oracle.domains_np import OracleDomain

I = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ITERS = int(sys.argv[2]) if len(sys.argv) > 2 else 20
MODE = {"bf16": L.DX_HEUR_NET_BF16, "fp32": L.DX_HEUR_NET_FP32, "zero": L.DX_HEUR_ZERO}[sys.argv[3] if len(sys.argv) > 3 else "bf16"]
B = 10000
dom = OracleDomain("cube3")
rng = np.random.default_rng(0)
goals = np.tile(dom.canonical_goal(), (I, 1))
st = goals.copy()
for _ in range(100):
    st = dom.next_state(st, rng.integers(0, 12, size=I))
sd = random_state_dict(324, 1000, 4, 1, True, seed=0)
blob, info = L.pack_state_dict({k: v.numpy() for k, v in sd.items()})
t0 = time.time()
eng = L.Engine("cube3", 0, "graph_v", MODE, max_instances=I, max_pop_total=I * B, max_nodes=I * B * 12 * (ITERS + 2))
if MODE != L.DX_HEUR_ZERO:
    eng.load_weights(54, 6, 1000, 4, 1, True, blob)
print(f"engine create+load {time.time()-t0:.2f}s", flush=True)
for rep in range(2):
    eng.reset()
    eng.set_profiling(rep == 1)
    t0 = time.time()
    eng.add_instances(st, goals, B, 0.6)
    t1 = time.time()
    done = eng.run(ITERS)
    t2 = time.time()
    c = eng.counters()
    gen = sum(s.num_nodes_generated for s in eng.status())
    print(f"rep{rep}: I={I} iters={done} gen={gen} evals={c.heur_evals} stored={c.nodes_stored} add={t1-t0:.3f}s run={t2-t1:.3f}s "
          f"nodes/s={gen/(t2-t0):.3e} launches={c.kernel_launches}", flush=True)
    if rep == 1:
        ms = eng.stage_ms()
        tot = sum(ms.values())
        print("stage ms:", {k: round(v, 2) for k, v in ms.items()}, "sum", round(tot, 2), flush=True)
if MODE != L.DX_HEUR_ZERO:
    for n in (8192, 65536, min(120000 * I, eng.counters().nodes_stored)):
        ms = eng.bench_heur(n, 5)
        flop = 2.0 * (324 * 1000 + 8 * 1000 * 1000 + 1000) * n * 5
        print(f"heur n={n}: {ms/5:.3f} ms/eval  {flop/ms/1e9:.1f} TFLOP/s (dense-equivalent, unpadded)  {n*5/ms*1e3:.3e} states/s", flush=True)
