"""GPU probe: raw network throughput (dx_bench_heur) for cube3 resnet_fc.1000H_4B_bn in bf16."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepxube_b200 import _lib as L
from oracle.nnet_torch import random_state_dict
sd = random_state_dict(324, 1000, 4, 1, True, seed=0)
blob, info = L.pack_state_dict({k: v.numpy() for k, v in sd.items()})
N = int(sys.argv[1]) if len(sys.argv) > 1 else 800000
eng = L.Engine("cube3", 0, "graph_v", L.DX_HEUR_NET_BF16, max_instances=1, max_pop_total=N // 12 + 1, max_nodes=N + 1024)
eng.load_weights(54, 6, 1000, 4, 1, True, blob)
for rep in range(3):
    ms = eng.bench_heur(N, 5)
    flop = 2.0 * (324 * 1000 + 8 * 1000 * 1000 + 1000) * N * 5
    print(f"heur n={N}: {ms/5:.3f} ms/eval  {flop/ms/1e9:.1f} TFLOP/s dense-equiv  {N*5/ms*1e3:.3e} states/s", flush=True)
