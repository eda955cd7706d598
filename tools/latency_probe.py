"""GPU probe: per-iteration latency of small-batch searches (tutorial cube3 KAT, graph.100B_0.1W, fp32 / bf16)."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepxube_b200 import _lib as L
g = os.path.join(ROOT, "tests", "golden")
k = np.load(os.path.join(g, "kat_cube3.npz"))
z = np.load(os.path.join(g, "nnet.npz"))
sd = {key[len("cube3v/"):]: z[key] for key in z.files if key.startswith("cube3v/")}
want = sys.argv[1].split(",") if len(sys.argv) > 1 else ["fp32", "bf16", "zero"]
for prec, mode in (("fp32", L.DX_HEUR_NET_FP32), ("bf16", L.DX_HEUR_NET_BF16), ("zero", L.DX_HEUR_ZERO)):
    if prec not in want:
        continue
    eng = L.Engine("cube3", 0, "graph_v", mode, max_instances=1, max_pop_total=100, max_nodes=600000)
    if mode != L.DX_HEUR_ZERO:
        blob, info = L.pack_state_dict(sd)
        eng.load_weights(54, 6, info["res_dim"], info["num_blocks"], 1, True, blob)
    eng.add_instances(k["hard_states"][:1], k["hard_goals"][:1], 100, 0.1)
    eng.run(50)
    eng.tc_profile()
    tot_it, tot_s, dev_ms = 0, 0.0, 0.0
    l0 = eng.counters().kernel_launches
    for i in range(10):
        eng.reset()
        eng.add_instances(k["hard_states"][i:i + 1], k["hard_goals"][i:i + 1], 100, 0.1)
        t0 = time.perf_counter()
        it = eng.run(60 if mode == L.DX_HEUR_ZERO else int(os.environ.get("DXB200_PROBE_ITERS", "100000")))
        tot_s += time.perf_counter() - t0
        tot_it += it
        dev_ms += eng.last_run_ms()
    launches = eng.counters().kernel_launches - l0
    print(f"{prec}: {tot_it} iterations, wall {1e6 * tot_s / tot_it:.1f} us/iteration, device {1e3 * dev_ms / tot_it:.1f} us/iteration, "
          f"{launches / tot_it:.1f} launches/iteration", flush=True)
    if os.environ.get("DXB200_PROF_TAIL"):
        pr = eng.tc_profile()
        names = ["launches", "plan", "staging+keys", "sort", "merge by ranking", "pop split", "pop ranking", "write-back", "bookkeeping"]
        print("   k_small_tail_v timeline (cycles per launch):", {k: round(v / max(pr[0], 1)) for k, v in zip(names[1:], pr[1:9]) if v})
        fn = ["launches", "goal resolve", "link (list push + barrier)", "decide", "scan", "scatter", "link: loads + hash", "link: probes"]
        print("   k_small_filter_v timeline (cycles per launch):", {k: round(v / max(pr[0], 1)) for k, v in zip(fn[1:], pr[9:16]) if v})
    elif mode != L.DX_HEUR_ZERO and os.environ.get("DXB200_NVCC_DEFS", "").find("SM_PROFILE") >= 0:
        pr = eng.tc_profile()
        names = ["launches", "rows staged", "one-hot built"] + ["layer prologue", "first k step"] + [f"layer{l}" for l in range(2, 8)] + ["k loops", "epilogues", "layer barriers", "output", "-"]
        print("   k_mlp_small CTA-0 timeline (cycles per launch):", {k: round(v / max(pr[0], 1)) for k, v in zip(names[1:], pr[1:]) if v})
    eng.close()
