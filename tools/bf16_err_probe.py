"""GPU probe: error of the bf16 tensor-core network against the fp32-accurate device network on the trained tutorial net
(cube3, resnet_fc.200H_2B_bn) over states around the KAT instances, for the kernel variants selected by environment
switches.  usage: bf16_err_probe.py"""
import json, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r + "/tests")
from deepxube_b200 import _lib as L
import golden_utils as gu
k = gu.load("kat_cube3")
rng = np.random.default_rng(0)
st = np.repeat(k["hard_states"], 600, axis=0)
steps = rng.integers(0, 25, size=len(st))
for t in range(25):
    live = np.nonzero(steps > t)[0]
    st[live] = L.next_state("cube3", 0, st[live], rng.integers(0, 12, size=live.size))
blob, info = L.pack_state_dict(gu.state_dict_np("cube3v"))
res = {}
outs = {}
for name, mode in (("fp32", L.DX_HEUR_NET_FP32), ("bf16", L.DX_HEUR_NET_BF16)):
    eng = L.Engine("cube3", 0, "graph_v", mode, max_instances=1, max_pop_total=1000, max_nodes=4096)
    eng.load_weights(54, 6, info["res_dim"], info["num_blocks"], 1, True, blob)
    outs[name] = eng.heur_eval(st, None, 1)[:, 0].astype(np.float64)
    eng.close()
e = np.abs(outs["bf16"] - outs["fp32"])
print("RESULT " + json.dumps({"n": len(st), "h_mean": outs["fp32"].mean(), "h_max": outs["fp32"].max(), "err_max": e.max(), "err_mean": e.mean(),
                              "err_rms": float(np.sqrt((e ** 2).mean())), "err_p99": float(np.percentile(e, 99))}))
'''
for label, env in (("default (fused first block)", {}), ("DXB200_FUSE_L1=0", {"DXB200_FUSE_L1": "0"}),
                   ("DXB200_FUSE_L1=0 DXB200_FUSE_IN=0", {"DXB200_FUSE_L1": "0", "DXB200_FUSE_IN": "0"})):
    r = subprocess.run([sys.executable, "-c", CHILD % (ROOT, ROOT)], env={**os.environ, **env}, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
    print(label, line[0][7:] if line else ("FAILED\n" + r.stdout[-1500:] + r.stderr[-2500:]), flush=True)
