"""GPU probe for the other BASELINE configs (C3 npuzzle A*, C4 cube3 Q*): nodes/s + stage breakdown.
usage: config_probe.py <domain> <dim> <algo v|q|b (beam_v)> <B> <I> <iters> [bf16|fp32|zero]"""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepxube_b200 import _lib as L
from oracle.nnet_torch import random_state_dict
from oracle.domains_np import OracleDomain

dname, dim, algo, B, I, ITERS = sys.argv[1], int(sys.argv[2]), sys.argv[3], int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
mode_s = sys.argv[7] if len(sys.argv) > 7 else "bf16"
MODE = {"bf16": L.DX_HEUR_NET_BF16, "fp32": L.DX_HEUR_NET_FP32, "zero": L.DX_HEUR_ZERO}[mode_s]
is_q = algo == "q"
dom = OracleDomain(dname, dim)
A = dom.num_actions
rng = np.random.default_rng(0)
goals = np.tile(dom.canonical_goal(), (I, 1))
st = goals.copy()
for _ in range(200):
    st = dom.next_state(st, rng.integers(0, A, size=I))
in_count, depth = dom.input_info()
out_dim = A if is_q else 1
sd = random_state_dict(in_count * depth, 1000, 4, out_dim, True, seed=0)
blob, info = L.pack_state_dict({k: v.numpy() for k, v in sd.items()})
per_iter_nodes = B if is_q else B * A
eng = L.Engine(dname, dim, "graph_q" if is_q else ("beam_v" if algo == "b" else "graph_v"), MODE, max_instances=I, max_pop_total=I * B,
               max_nodes=int(I * per_iter_nodes * (ITERS + 2)) + 4096)
if MODE != L.DX_HEUR_ZERO:
    eng.load_weights(in_count, depth, 1000, 4, out_dim, True, blob)
for rep in range(2):
    eng.reset()
    eng.set_profiling(rep == 1)
    t0 = time.time()
    eng.add_instances(st, goals, B, 0.6)
    done = eng.run(ITERS)
    dt = time.time() - t0
    c = eng.counters()
    gen = sum(s.num_nodes_generated for s in eng.status())
    print(f"{dname}.{dim} {'Q*' if is_q else ('beam' if algo == 'b' else 'A*')} B={B} I={I} {mode_s} rep{rep}: iters={done} gen={gen} evals={c.heur_evals} "
          f"open={sum(s.open_size for s in eng.status())} time={dt:.3f}s nodes/s={gen/dt:.3e} dev_ms={eng.last_run_ms():.1f}", flush=True)
    if rep == 1:
        print("  stage ms:", {k: round(v, 2) for k, v in eng.stage_ms().items()}, flush=True)
