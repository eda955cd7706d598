"""GPU debug: bf16 network with / without bias folding and output-layer fusion on the trained cube3v golden net."""
import sys, os, subprocess
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
if len(sys.argv) > 1:
    from deepxube_b200 import _lib as L
    import golden_utils as gu
    z = gu.load("nnet")
    sd = gu.state_dict_np("cube3v")
    st = z["cube3v_states"]
    blob, info = L.pack_state_dict(sd)
    eng = L.Engine("cube3", 0, "graph_v", L.DX_HEUR_NET_BF16, max_instances=4, max_pop_total=512, max_nodes=1 << 16)
    eng.load_weights(eng.in_count, eng.depth, info["res_dim"], info["num_blocks"], info["out_dim"], info["batch_norm"], blob)
    out = eng.heur_eval(st, None, 1).ravel()
    out2 = eng.heur_eval(st, None, 1).ravel()
    np.save(sys.argv[1], out)
    print(sys.argv[1], "repeatable", np.array_equal(out, out2), out[:8])
else:
    import torch
    import golden_utils as gu
    from oracle.nnet_torch import resnet_fc_forward
    sd = {k: torch.tensor(v) for k, v in gu.state_dict_np("cube3v").items()}
    ref = np.maximum(np.asarray(resnet_fc_forward(sd, gu.load("nnet")["cube3v_states"].astype(np.int64), 6)).ravel(), 0)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if os.path.exists(os.path.join(root, "ab_head.so")):
        for rep in range(2):
            subprocess.run([sys.executable, __file__, "/tmp/o_head.npy"], env=dict(os.environ, DXB200_LIB=os.path.join(root, "ab_head.so")), check=True)
            print("HEAD lib: max err vs fp32 =", np.abs(np.load("/tmp/o_head.npy") - ref).max())
    for fill in ("0", "255", "127", "60"):
        subprocess.run([sys.executable, __file__, "/tmp/o_fill.npy"], env=dict(os.environ, DXB200_DEBUG_FILL=fill, DXB200_FOLD_BIAS="0", DXB200_FUSE_OUT="0"), check=True)
        print("fill", fill, ": max err vs fp32 =", np.abs(np.load("/tmp/o_fill.npy") - ref).max())
    outs = {}
    for fold in "01":
        for fuse in "01":
            env = dict(os.environ, DXB200_FOLD_BIAS=fold, DXB200_FUSE_OUT=fuse)
            f = f"/tmp/o_{fold}{fuse}.npy"
            subprocess.run([sys.executable, __file__, f], env=env, check=True)
            outs[fold + fuse] = np.load(f)
    base = outs["00"]
    for k, v in outs.items():
        d = np.abs(v - base)
        print(f"fold={k[0]} fuse={k[1]}: max diff vs 00 = {d.max():.4f} at row {d.argmax()}  n_bad={(d > 0.1).sum()}  max err vs fp32 = {np.abs(v - ref).max():.4f}")
