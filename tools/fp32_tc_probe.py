"""GPU probe: accuracy and speed of the fp32-accurate network paths (split-fp16 tensor-core vs CUDA-core SGEMM) against a
float64 torch evaluation of the same network.   usage: fp32_tc_probe.py [rows]"""
import os, subprocess, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

CHILD = r'''
import sys, json, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r + "/tests")
from deepxube_b200 import _lib as L
import golden_utils as gu
import torch
from oracle.nnet_torch import random_state_dict, resnet_fc_forward
out = {}
rng = np.random.default_rng(0)
def f64(sd, x, depth):
    sd64 = {k: v.double() for k, v in sd.items()}
    return np.asarray(resnet_fc_forward(sd64, x.astype(np.int64), depth), np.float64)
for tag, dname, dim, depth, in_count, sd, q in (
        ("rand1000H_4B", "cube3", 0, 6, 54, random_state_dict(324, 1000, 4, 1, True, seed=3, randomize_bn=True), False),
        ("rand1000H_4B_q", "cube3", 0, 6, 54, random_state_dict(324, 1000, 4, 12, True, seed=4, randomize_bn=True), True),
        ("cube3v_trained", "cube3", 0, 6, 54, gu.state_dict_torch("cube3v"), False),
        ("np5v", "npuzzle", 5, 25, 25, gu.state_dict_torch("np5v"), False)):
    n = int(sys.argv[1])
    if tag == "cube3v_trained":
        st = np.concatenate([gu.load("nnet")["cube3v_states"], gu.load("kat_cube3")["hard_states"]])
    elif dname == "cube3":
        st = rng.integers(0, 6, size=(n, 54), dtype=np.uint8)
    else:
        st = np.stack([rng.permutation(25) for _ in range(min(n, 4096))]).astype(np.uint8)
    od = 12 if q else 1
    ref = np.maximum(f64(sd, st, depth).reshape(len(st), od), 0)
    blob, info = L.pack_state_dict({k: v.numpy() for k, v in sd.items()})
    eng = L.Engine(dname, dim, "graph_q" if q else "graph_v", L.DX_HEUR_NET_FP32, max_instances=1, max_pop_total=100000 if q else 9000, max_nodes=300000)
    eng.load_weights(in_count, depth, info["res_dim"], info["num_blocks"], od, info["batch_norm"], blob)
    got = eng.heur_eval(st, None, od)
    ms = eng.bench_heur(min(len(st), 100000), 3) / 3 if len(st) >= 1000 else None
    out[tag] = {"max_abs_err": float(np.abs(got - ref).max()), "max_ref": float(ref.max()), "ms_per_eval": ms, "rows": int(min(len(st), 100000))}
    eng.close()
print("RESULT " + json.dumps(out))
'''
rows = sys.argv[1] if len(sys.argv) > 1 else "100000"
for label, env in (("tensor-core split-fp16", {}), ("CUDA-core SGEMM", {"DXB200_FP32_TC": "0"})):
    r = subprocess.run([sys.executable, "-c", CHILD % (ROOT, ROOT), rows], env={**os.environ, **env}, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
    print(label, json.dumps(json.loads(line[0][7:]), indent=1) if line else ("FAILED\n" + r.stdout[-2000:] + r.stderr[-3000:]), flush=True)
