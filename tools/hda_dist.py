"""HDA* over NCCL (one process per GPU):  torchrun --nproc-per-node W tools/hda_dist.py [--check] [--bench]
  --check  solve a KAT instance with the trained net through the NCCL transport and (rank 0) compare with the
           same W emulated in-process on one GPU: the two transports must give identical results
  --bench  BASELINE config 5: one cube3 search, B = 10000 per GPU, W = 0.6, random-init resnet_fc.1000H_4B_bn,
           bf16, fixed number of iterations; prints nodes generated / s (device work + exchange, wall clock)"""
import argparse, json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import deepxube_b200  # noqa: F401
from deepxube_b200.hda import HdaSearch
from deepxube_b200.factories.domain_factory import get_domain_from_arg
from deepxube_b200.factories.pathfind_fns_factory import get_path_fns_nnet_par_dict

ap = argparse.ArgumentParser()
ap.add_argument("--check", action="store_true"); ap.add_argument("--bench", action="store_true")
ap.add_argument("--iters", type=int, default=20); ap.add_argument("--batch-per-gpu", type=int, default=10000)
args = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
domain, name = get_domain_from_arg("cube3")

if args.check:
    import golden_utils as gu
    k = gu.load("kat_cube3")
    pfns, _ = get_path_fns_nnet_par_dict(domain, name, ["heurv,resnet_fc.200H_2B_bn"])
    pfns.heurv.nnet.load_state_dict(gu.state_dict_np("cube3v"))
    h = HdaSearch("cube3", 0, world, 100, 0.1, nnet=pfns.heurv.nnet, precision="fp32", max_nodes_per_rank=1 << 20, dist=dist, device=local)
    res = [h.solve(k["hard_states"][i], k["hard_goals"][i], max_itrs=3000) for i in (4, 8)]
    h.close()
    if rank == 0:
        e = HdaSearch("cube3", 0, world, 100, 0.1, nnet=pfns.heurv.nnet, precision="fp32", max_nodes_per_rank=1 << 20, device=local)
        emu = [e.solve(k["hard_states"][i], k["hard_goals"][i], max_itrs=3000) for i in (4, 8)]
        e.close()
        print("W ranks :", [(r["path_cost"], r["num_nodes_generated"], r["iterations"]) for r in res])
        print("emulated:", [(r["path_cost"], r["num_nodes_generated"], r["iterations"]) for r in emu])
        assert res == emu, "distributed run and in-process emulation disagree"
        print(f"HDA* check OK: {world} processes ({h.exchange} exchange) == emulation")

if args.bench:
    pfns, _ = get_path_fns_nnet_par_dict(domain, name, ["heurv,resnet_fc.1000H_4B_bn"])
    pfns.heurv.nnet.init_random(0)
    rng = np.random.default_rng(0)
    goal = (np.arange(54) // 9).astype(np.uint8)
    from deepxube_b200.domains.tables import cube3_perm_table
    st = goal.copy()
    for a in rng.integers(0, 12, size=2000):
        st = st[cube3_perm_table()[a]]
    B = args.batch_per_gpu * world
    for dev_res, exch in ((True, "p2p"), (True, "nccl"), (False, "nccl")):
        h = HdaSearch("cube3", 0, world, B, 0.6, nnet=pfns.heurv.nnet, precision="bf16",
                      max_nodes_per_rank=int(args.batch_per_gpu * 12 * (args.iters + 2) * 1.5), dist=dist, device=local,
                      device_resident=dev_res, exchange=exch)
        for rep in range(3):
            dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = h.solve(st, goal, max_itrs=args.iters)
            torch.cuda.synchronize(); dist.barrier()
            dt = time.perf_counter() - t0
            if rank == 0:
                print(json.dumps({"config": "C5 HDA* cube3", "driver": (f"device-resident/{h.exchange}" if dev_res else "host-driven"), "n_gpus": world,
                                  "global_batch": B, "iters": r["iterations"], "nodes_generated": r["num_nodes_generated"], "seconds": dt,
                                  "nodes_per_sec": r["num_nodes_generated"] / dt, "rep": rep}))
        h.close()
dist.barrier()
dist.destroy_process_group()
