"""GPU probe for the training-data slice: many tiny instances (batch size 1, like the reference's updater,
base/updater.py:899-900), every generated child evaluated by the network, Bellman backups logged on the device.
Reports labelled states/s and children evaluated/s next to the default (kept-children-only) search on the same instances,
and the CPU oracle on a bounded sample.
usage: backup_probe.py <domain> <dim> <I> <search_itrs> [bf16|fp32] [H] [NB] [v|q] [bellman|ub_solns|lhbl]"""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepxube_b200 import _lib as L
from oracle.nnet_torch import random_state_dict, heurv_from_state_dict
from oracle.domains_np import OracleDomain
from oracle.search import OracleSearch

dname, dim, I, ITERS = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
mode_s = sys.argv[5] if len(sys.argv) > 5 else "bf16"
H = int(sys.argv[6]) if len(sys.argv) > 6 else 1000
NB = int(sys.argv[7]) if len(sys.argv) > 7 else 4
ALGO = sys.argv[8] if len(sys.argv) > 8 else "v"
LABEL = sys.argv[9] if len(sys.argv) > 9 else "bellman"
is_q = ALGO == "q"
MODE = {"bf16": L.DX_HEUR_NET_BF16, "fp32": L.DX_HEUR_NET_FP32}[mode_s]
dom = OracleDomain(dname, dim)
A = dom.num_actions
rng = np.random.default_rng(0)
goals = np.tile(dom.canonical_goal(), (I, 1))
st = goals.copy()
nsteps = rng.integers(0, 31, size=I)                     # steps_gen ~ U[0, step_max], step_max 30
for t in range(30):
    mv = dom.next_state(st, rng.integers(0, A, size=I))
    st = np.where((nsteps > t)[:, None], mv, st)
in_count, depth = dom.input_info()
sd = random_state_dict(in_count * depth, H, NB, A if is_q else 1, True, seed=0)
blob, info = L.pack_state_dict({k: v.numpy() for k, v in sd.items()})
for backups in (0, 1):
    eng = L.Engine(dname, dim, "graph_q" if is_q else "graph_v", MODE, max_instances=I, max_pop_total=I,
                   max_nodes=I * (1 if is_q else A) * (ITERS + 1) + 4096, max_backups=I * ITERS * backups)
    eng.load_weights(in_count, depth, H, NB, A if is_q else 1, True, blob)
    for rep in range(3):
        eng.reset()
        t0 = time.time()
        eng.add_instances(st, goals, 1, 0.8)
        done = eng.run(ITERS)
        t1 = time.time()
        n_lab = len(eng.drain_backups(LABEL)[-1]) if backups else 0
        dt = time.time() - t0
        t_drain = time.time() - t1
        c = eng.counters()
        gen = sum(s.num_nodes_generated for s in eng.status())
    print(f"{dname}.{dim} {'Q*' if is_q else 'A*'} I={I} itrs={ITERS} {mode_s} {H}H_{NB}B backups={backups} labels={LABEL} drain={t_drain * 1e3:.1f}ms: iters={done} gen={gen} evals={c.heur_evals} "
          f"labelled={n_lab} wall={dt:.3f}s dev_ms={eng.last_run_ms():.1f} children/s={gen / dt:.3e} "
          f"labelled/s={n_lab / dt:.3e} (device-only {n_lab / max(eng.last_run_ms(), 1e-9) * 1e3:.3e})", flush=True)
    eng.close()
# CPU oracle on a bounded sample (same network in torch fp32, all host threads)
n_cpu = min(I, 2000)
from oracle.nnet_torch import heurq_from_state_dict
f = (heurq_from_state_dict if is_q else heurv_from_state_dict)(sd, depth)
orc = OracleSearch(dom, lambda s, g: f(dom.nnet_input(s, g)), is_q, 1, 0.8)
t0 = time.time()
orc.add_instances(st[:n_cpu], goals[:n_cpu])
for _ in range(ITERS):
    orc.step()
lab = (orc.labels_q if is_q else orc.labels_v)(LABEL)
dt = time.time() - t0
n_lab = len(lab)
print(f"cpu oracle sample I={n_cpu}: labelled={n_lab} wall={dt:.2f}s labelled/s={n_lab / dt:.3e}", flush=True)
# replay-buffer labels (dx_vi_targets): n sampled states -> expand + network + min, HOST buffers in and out
if not is_q:
    from oracle.search import vi_targets_v
    n_vi = min(I, 100000)
    eng = L.Engine(dname, dim, "graph_v", MODE, max_instances=1, max_pop_total=n_vi, max_nodes=4096)
    eng.load_weights(in_count, depth, H, NB, 1, True, blob)
    for rep in range(3):
        t0 = time.time()
        lab = eng.vi_targets(st[:n_vi], goals[:n_vi])
        dt = time.time() - t0
    print(f"vi_targets {dname}.{dim} n={n_vi} ({n_vi * A} evaluations) {mode_s} {H}H_{NB}B: wall={dt * 1e3:.1f}ms labels/s={n_vi / dt:.3e} evals/s={n_vi * A / dt:.3e}", flush=True)
    n_cpu = min(n_vi, 4000)
    t0 = time.time()
    ref = vi_targets_v(dom, lambda s, g: f(dom.nnet_input(s, g)), st[:n_cpu], goals[:n_cpu])
    dt = time.time() - t0
    print(f"cpu oracle vi_targets n={n_cpu}: wall={dt:.2f}s labels/s={n_cpu / dt:.3e}  max |gpu - cpu| = {np.abs(lab[:n_cpu] - ref).max():.4f}", flush=True)
    eng.close()
