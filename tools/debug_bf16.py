"""GPU debug helper: device bf16 (tcgen05) network vs device fp32 network vs torch reference."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from deepxube_b200 import _lib as L
from oracle.nnet_torch import random_state_dict, resnet_fc_forward
import golden_utils as gu


def run(tag, sd_np, dname, dim, states, goals, out_dim, q):
    res = {}
    for mode, name in ((L.DX_HEUR_NET_FP32, "fp32"), (L.DX_HEUR_NET_BF16, "bf16")):
        eng = L.Engine(dname, dim, "graph_q" if q else "graph_v", mode, max_instances=4, max_pop_total=max(4096 // (12 if dname == "cube3" else 4), 64), max_nodes=1 << 16)
        blob, info = L.pack_state_dict(sd_np)
        eng.load_weights(eng.in_count, eng.depth, info["res_dim"], info["num_blocks"], info["out_dim"], info["batch_norm"], blob)
        t = time.time()
        res[name] = eng.heur_eval(states, goals, out_dim)
        print(f"  {tag} {name}: {time.time() - t:.3f}s  first={res[name][:3].ravel()[:6]}", flush=True)
        eng.close()
    err = np.abs(res["bf16"] - res["fp32"])
    print(f"{tag}: n={states.shape[0]} max|bf16-fp32|={err.max():.4f} mean={err.mean():.5f}  |fp32| mean={np.abs(res['fp32']).mean():.3f} max={np.abs(res['fp32']).max():.3f}", flush=True)
    return res


if __name__ == "__main__":
    z = gu.load("nnet")
    run("gridv(H100)", gu.state_dict_np("gridv"), "grid", 7, z["gridv_states"], z["gridv_goals"], 1, False)
    run("cube3v(H200)", gu.state_dict_np("cube3v"), "cube3", 0, z["cube3v_states"], None, 1, False)
    run("cube3q(H128)", gu.state_dict_np("cube3q"), "cube3", 0, z["cube3q_states"], None, 12, True)
    rng = np.random.default_rng(0)
    st = rng.integers(0, 6, size=(3000, 54), dtype=np.uint8)
    sd = random_state_dict(324, 1000, 4, 1, True, seed=0, randomize_bn=True)
    r = run("cube3 1000H_4B", {k: v.numpy() for k, v in sd.items()}, "cube3", 0, st, None, 1, False)
    ref = resnet_fc_forward(sd, st.astype(np.int64), 6)
    print("vs torch: fp32 err", np.abs(np.maximum(ref, 0) - r["fp32"]).max(), "bf16 err", np.abs(np.maximum(ref, 0) - r["bf16"]).max())
