"""Barrier-wait accounting of the tcgen05 GEMM (library must be built with DXB200_NVCC_DEFS=-DTC_PROFILE)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepxube_b200 import _lib as L
from oracle.nnet_torch import random_state_dict
n = int(sys.argv[1]) if len(sys.argv) > 1 else 600000
NB = int(sys.argv[2]) if len(sys.argv) > 2 else 4      # 0 blocks: only the input-layer GEMM runs
sd = random_state_dict(324, 1000, NB, 1, True, seed=0)
blob, info = L.pack_state_dict({k: v.numpy() for k, v in sd.items()})
eng = L.Engine("cube3", 0, "graph_v", L.DX_HEUR_NET_BF16, max_instances=8, max_pop_total=80000, max_nodes=2500000)
eng.load_weights(54, 6, 1000, NB, 1, True, blob)
# fill the node store with real states (bench_heur evaluates the stored rows as they are)
from oracle.domains_np import OracleDomain
dom = OracleDomain("cube3")
rng = np.random.default_rng(0)
goals = np.tile(dom.canonical_goal(), (8, 1))
st = goals.copy()
for _ in range(100):
    st = dom.next_state(st, rng.integers(0, 12, size=8))
eng.add_instances(st, goals, 10000, 0.6)
eng.run(5)
n = min(n, eng.counters().nodes_stored)
eng.bench_heur(n, 2)
eng.tc_profile()
ms = eng.bench_heur(n, 3)
p = eng.tc_profile()
names = ["mma<-operands", "mma<-tmem_free", "producer<-slot", "epi<-tmem_full", "epi<-residual", "epi<-store_drain", "epi_total", "mma_total", "epi<-tmem_ld", "epi math+stage", "epi fence+sync", "epi tma issue", "gen<-slot", "gen masks", "gen expand", "gen_total"]
print(f"n={n} blocks={NB} {ms/3:.3f} ms per eval (profiling build)")
for k, v in zip(names, p):
    base = p[7] if k.startswith(("mma", "producer")) else (p[15] if k.startswith("gen") else p[6])
    print(f"  {k:18s} {v/1e6:10.1f} Mcycles  {100.0*v/max(base,1):5.1f}% of role total")
