"""bench.py -- nodes generated / second of batch weighted A* on cube3 (BASELINE.json configs[1], "C2").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (config.workload): cube3, `graph.10000B_0.6W` (batch weighted A*, B=10000, W=0.6), heuristic
`heurv,resnet_fc.1000H_4B_bn` with random-init weights (seed 0), 1000 synthetic scrambles (random walks of
1000..10000 moves from the solved cube, seed 0).  A random-init heuristic never terminates a search, so every
instance runs a fixed number of iterations (--max-itrs, default 20; SURVEY.md 8d "C2 inputs").
One STEP = one PathFind holding G instances (--insts-per-step, default 8) advanced in lock-step for
max_itrs iterations: expand -> CLOSED filter -> heuristic network -> sort/merge/pop, all on the device.
  value   nodes generated / s with the start states already resident in HBM (timed region = the search
          iterations, CUDA events on the engine's stream, summed over the K steps, max over ranks)
  e2e     the same metric through the reference-facing API with HOST buffers: make_instances (H2D of the
          start states) + step loop + reading back every instance's status (D2H), wall clock, max over ranks
Multi-GPU: instances are independent, so ranks take disjoint scrambles (weak scaling, no data-path collective).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, NB, IN_COUNT, DEPTH, ACTS = 1000, 4, 54, 6, 12
FLOP_PER_STATE = 2.0 * (IN_COUNT * DEPTH * H + 2 * NB * H * H + H * 1)      # 16.65 MFLOP (V)
FLOP_PER_STATE_HIDDEN_LAYER = 2.0 * H * H                                    # one hidden-layer GEMM launch, per state


def make_scrambles(n: int, seed: int = 0):
    """n cube3 start states: solved cube + random walk of U{1000..10000} moves (host-side input generation)."""
    from deepxube_b200.domains.tables import cube3_perm_table
    perm = cube3_perm_table()
    rng = np.random.default_rng(seed)
    goal = (np.arange(54) // 9).astype(np.uint8)
    st = np.tile(goal, (n, 1))
    steps = rng.integers(1000, 10001, size=n)
    for t in range(int(steps.max())):
        live = np.nonzero(steps > t)[0]
        acts = rng.integers(0, 12, size=live.size)
        st[live] = np.take_along_axis(st[live], perm[acts], axis=1)
    return st, np.tile(goal, (n, 1))


def random_weights(seed: int = 0):
    """state_dict of resnet_fc.1000H_4B_bn for cube3 with nn.Linear's default init, eval-mode BN (mu 0, var 1)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def lin(o, i, p):
        b = 1.0 / np.sqrt(i)
        sd[p + ".weight"] = ((torch.rand(o, i, generator=g) * 2 - 1) * b).numpy()
        sd[p + ".bias"] = ((torch.rand(o, generator=g) * 2 - 1) * b).numpy()

    lin(H, IN_COUNT * DEPTH, "heur.0")
    for b in range(NB):
        for j in range(2):
            p = f"heur.1.blocks.{b}.0.layers.{j}"
            lin(H, H, p + ".0")
            sd[p + ".1.weight"] = np.ones(H, np.float32)
            sd[p + ".1.bias"] = np.zeros(H, np.float32)
            sd[p + ".1.running_mean"] = np.zeros(H, np.float32)
            sd[p + ".1.running_var"] = np.ones(H, np.float32)
    lin(1, H, "heur.2")
    return sd


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU with NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {pynvml.nvmlClocksEventReasonHwSlowdown: "hw_slowdown",
                     pynvml.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     pynvml.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     pynvml.nvmlClocksEventReasonSwPowerCap: "sw_power_cap",
                     pynvml.nvmlClocksEventReasonHwPowerBrakeSlowdown: "hw_power_brake"}
            self.ok = True
            while not self._stop_evt.is_set():
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
                time.sleep(0.05)
        except Exception as e:  # noqa: BLE001
            self.err = str(e)

    def stop(self) -> dict:
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def physical_gpu_index(local: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except Exception:  # noqa: BLE001
            return local
    return local


def load_peaks() -> dict:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        d["_source"] = "measured"
        return d
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "_source": "fallback"}


# ---------------------------------------------------------------------------------------------------
def cpu_reference_step(sd_torch, state, goal, iters: int, threads: int):
    """One bounded sample of the reference algorithm on the host (oracle port, torch-CPU network evaluating
    every generated child like the reference does)."""
    import torch
    from oracle.domains_np import OracleDomain
    from oracle.nnet_torch import heurv_from_state_dict
    from oracle.search import solve_one
    torch.set_num_threads(threads)
    dom = OracleDomain("cube3")
    f = heurv_from_state_dict(sd_torch, DEPTH)
    t0 = time.perf_counter()
    r = solve_one(dom, state, goal, lambda s, g: f(s.astype(np.int64)), False, 10000, 0.6, max_itrs=iters)
    return r["num_nodes_generated"], time.perf_counter() - t0


def run_reference(args, rank: int, world: int) -> None:
    """--impl reference: the reference's CPU algorithm (oracle port; the reference itself is Python and
    cannot travel to the GPU box) on rank 0's host cores; other ranks exit without work."""
    if rank != 0:
        return
    import torch
    threads = os.cpu_count() or 8
    sd = {k: torch.from_numpy(np.array(v)) for k, v in random_weights(0).items()}
    st, gl = make_scrambles(max(args.steps + args.warmup, 1), seed=0)
    iters = args.ref_itrs
    for i in range(args.warmup):
        cpu_reference_step(sd, st[i], gl[i], iters, threads)
    nodes, secs = 0, 0.0
    for i in range(args.steps):
        n, s = cpu_reference_step(sd, st[args.warmup + i], gl[args.warmup + i], iters, threads)
        nodes += n
        secs += s
    val = nodes / secs
    sample = f"1 instance x {iters} iterations ({nodes // max(args.steps, 1)} nodes) per step"
    line = {"impl": "reference", "metric": "nodes_generated_per_sec", "value": val, "unit": "nodes/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(args.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, 1),
            "cpu_baseline": {"value": val, "unit": "nodes/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "nodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def kat_solve_times(L, device: int) -> dict:
    """Solve time / instance (the second half of BASELINE's metric) on the reference's own known-answer set:
    tutorial/cube3 `hard` instances with the trained resnet_fc.200H_2B_bn heuristic, graph.100B_0.1W
    (fixtures under tests/golden/, generated from the unmodified reference by tools/gen_golden.py).  Timed like
    deepxube/_solve.py:150-161 (instance creation + every step), one instance at a time, fp32 network (bit-exact path)."""
    g = os.path.join(ROOT, "tests", "golden")
    k = np.load(os.path.join(g, "kat_cube3.npz"))
    z = np.load(os.path.join(g, "nnet.npz"))
    sd = {key[len("cube3v/"):]: z[key] for key in z.files if key.startswith("cube3v/")}
    out = {}
    for prec, mode in (("fp32", L.DX_HEUR_NET_FP32), ("bf16", L.DX_HEUR_NET_BF16)):
        eng = L.Engine("cube3", 0, "graph_v", mode, max_instances=1, max_pop_total=100, max_nodes=600000, device=device)
        blob, info = L.pack_state_dict(sd)
        eng.load_weights(IN_COUNT, DEPTH, info["res_dim"], info["num_blocks"], 1, True, blob)
        eng.add_instances(k["hard_states"][:1], k["hard_goals"][:1], 100, 0.1)      # warm-up: graph capture
        eng.run(100000)
        secs, nodes, costs = [], [], []
        for i in range(10):
            t0 = time.perf_counter()
            eng.reset()
            eng.add_instances(k["hard_states"][i:i + 1], k["hard_goals"][i:i + 1], 100, 0.1)
            eng.run(100000)
            stt = eng.status()[0]
            secs.append(time.perf_counter() - t0)
            nodes.append(int(stt.num_nodes_generated))
            costs.append(float(stt.ub))              # path cost of the goal node = ub
        eng.close()
        out[prec] = {"instances": 10, "mean_solve_time_s": float(np.mean(secs)), "max_solve_time_s": float(np.max(secs)),
                     "nodes_per_s": float(np.sum(nodes) / np.sum(secs)), "mean_path_cost": float(np.mean(costs)),
                     "path_costs_equal_reference": bool(np.array_equal(np.array(costs), k["hard_path_costs"])),
                     "nodes_generated_equal_reference": bool(np.array_equal(np.array(nodes), k["hard_nodes"]))}
    out["workload"] = "tutorial/cube3 hard set (10 instances), heurv resnet_fc.200H_2B_bn trained weights, graph.100B_0.1W"
    out["reference_mean_path_cost"] = float(np.mean(k["hard_path_costs"]))
    return out


def workload_config(args, world: int) -> dict:
    return {"workload": "cube3 batch weighted A* graph.10000B_0.6W, heurv resnet_fc.1000H_4B_bn random-init, "
                        "1000 synthetic scrambles (BASELINE.json configs[1])",
            "domain": "cube3", "pathfind": "graph.10000B_0.6W", "heuristic": "heurv,resnet_fc.1000H_4B_bn",
            "instances_per_step": args.insts_per_step, "max_itrs": args.max_itrs, "num_scrambles": 1000,
            "parallelism": f"instance-sharded x{world}", "l2_policy": "inputs larger than L2 (per-iteration working set "
            "~0.6 GB of activations per 120k-state batch >> 126 MB L2); no flush needed"}


# ---------------------------------------------------------------------------------------------------
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="b200", choices=["b200", "reference"])
    ap.add_argument("--insts-per-step", type=int, default=8)
    ap.add_argument("--max-itrs", type=int, default=20)
    ap.add_argument("--ref-itrs", type=int, default=5, help="iterations per bounded CPU sample")
    ap.add_argument("--precision", type=str, default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kat", action="store_true", help="skip the solve-time-per-instance report on the trained-net KAT set")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from deepxube_b200 import _lib as L
    G, ITERS, B = args.insts_per_step, args.max_itrs, 10000
    mode = L.DX_HEUR_NET_BF16 if args.precision == "bf16" else L.DX_HEUR_NET_FP32
    sd = random_weights(0)
    blob, info = L.pack_state_dict(sd)
    st_all, gl_all = make_scrambles(1000, seed=0)
    eng = L.Engine("cube3", 0, "graph_v", mode, max_instances=G, max_pop_total=G * B,
                   max_nodes=int(G * B * ACTS * (ITERS + 1) * 0.95) + 1024, device=local)
    eng.load_weights(IN_COUNT, DEPTH, H, NB, 1, True, blob)

    def pick(step_idx: int):
        base = ((step_idx * world + rank) * G) % 1000
        idx = (base + np.arange(G)) % 1000
        return st_all[idx], gl_all[idx]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- pass 1: device-resident metric ("value") + kernel timing for the roofline --------------------
    # Profiling level 2 brackets only the dominant kernel group (the 2*NB back-to-back hidden-layer GEMM launches of every
    # network evaluation) with CUDA events recorded inside the captured step graph; every other launch runs untimed.
    eng.set_profiling(2)
    for w in range(args.warmup):
        s, g = pick(w)
        eng.reset()
        eng.add_instances(s, g, B, 0.6)
        eng.run(ITERS)
    eng.set_profiling(2)                          # clears the timers; the captured graphs stay
    sampler = ClockSampler(physical_gpu_index(local))
    sampler.start()
    launches0 = eng.counters().kernel_launches
    barrier()
    dev_ms, nodes, evals = 0.0, 0, 0
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        s, g = pick(args.warmup + k)
        eng.reset()
        eng.add_instances(s, g, B, 0.6)          # start states -> HBM (outside the device-timed region)
        eng.run(ITERS)
        dev_ms += eng.last_run_ms()
        c = eng.counters()
        nodes += c.children_generated
        evals += c.heur_evals
    barrier()
    wall_value_pass = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    launches = eng.counters().kernel_launches - launches0
    stage = eng.stage_ms()
    calls = eng.stage_calls()
    # informational stage split: ONE extra step with every stage timed (outside both timed regions)
    eng.set_profiling(1)
    s, g = pick(args.warmup + args.steps)
    eng.reset()
    eng.add_instances(s, g, B, 0.6)
    eng.run(ITERS)
    stage_split = eng.stage_ms()
    eng.set_profiling(0)

    # ---- pass 2: end-to-end through the reference-facing API (registries -> PathFind -> C ABI), host objects ----
    # the `solve` loop of the reference (deepxube/_solve.py:146-160) with G instances per PathFind: make_instances
    # from host State/Goal objects (H2D inside), one Python-level step() per iteration (status D2H inside).
    eng.close()
    import deepxube_b200  # noqa: F401
    from deepxube_b200.domains.cube3 import Cube3Goal
    from deepxube_b200.factories.domain_factory import get_domain_from_arg
    from deepxube_b200.factories.pathfind_fns_factory import get_path_fns_nnet_par_dict
    from deepxube_b200.factories.pathfinding_factory import get_pathfind_from_arg
    domain, dname = get_domain_from_arg("cube3")
    pfns, _ = get_path_fns_nnet_par_dict(domain, dname, ["heurv,resnet_fc.1000H_4B_bn"], precision=args.precision)
    pfns.heurv.nnet.load_state_dict(sd)
    pathfind = get_pathfind_from_arg(domain, pfns, "graph.10000B_0.6W")[0]
    pathfind._max_nodes = int(G * B * ACTS * (ITERS + 1) * 0.95) + 1024

    def e2e_step(step_idx: int) -> int:
        s, g = pick(step_idx)
        pathfind.reset()
        insts = pathfind.make_instances(domain.rows_to_states(s), [Cube3Goal(x) for x in g], None, True)
        pathfind.add_instances(insts)
        itrs = 0
        while not min(x.finished() for x in pathfind.instances) and itrs < ITERS:
            pathfind.step()
            itrs += 1
        return int(sum(x.num_nodes_generated for x in pathfind.instances))

    e2e_step(0)                                   # engine creation + graph capture outside the timed region
    barrier()
    t0 = time.perf_counter()
    e2e_nodes = 0
    for k in range(args.steps):
        e2e_nodes += e2e_step(args.warmup + k)
    barrier()
    e2e_s = time.perf_counter() - t0
    h2d = int(G * 54 * 2 + G * 12)                # start + goal rows, batch sizes and weights
    d2h = int(ITERS * G * 56)                     # dx_inst_status of every instance after every step()
    pathfind.close()

    # ---- reduce over ranks: max time, sum of work -----------------------------------------------------------
    tot = torch.tensor([float(nodes), float(e2e_nodes), float(evals), float(launches)], device="cuda", dtype=torch.float64)
    tmax = torch.tensor([dev_ms, e2e_s, stage["gemm_hidden"]], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tot, tmax = tot.tolist(), tmax.tolist()

    if rank == 0:
        peaks = load_peaks()
        value = tot[0] / (tmax[0] / 1e3)
        e2e_val = tot[1] / tmax[1]
        # roofline of the dominant kernel (hidden-layer tcgen05 GEMM, 8 launches per network evaluation):
        # algorithmic FLOPs of this rank's launches / their summed CUDA-event durations
        hidden_flop = FLOP_PER_STATE_HIDDEN_LAYER * 2 * NB * evals
        achieved = hidden_flop / (stage["gemm_hidden"] / 1e3) / 1e12 if stage["gemm_hidden"] > 0 else None
        peak = peaks["bf16_tflops_sustained"] if args.precision == "bf16" else None
        traffic, traffic_note = None, None
        tf = os.path.join(ROOT, "profiles", "r1i_gemm_traffic.json")
        if args.precision == "bf16" and os.path.exists(tf):
            tj = json.load(open(tf))
            traffic = tj["traffic_per_launch_avg_bytes"]
            traffic_note = (f"dram__bytes_read.sum + dram__bytes_write.sum per launch from one ncu --set full capture of a "
                            f"{tj['rows_in_launch']}-state launch (profiles/r1i_gemm_traffic.json); algorithmic bytes of that launch: "
                            f"{tj['algorithmic_bytes']['hidden']} (plain) / {tj['algorithmic_bytes']['hidden_residual']} (with residual)")
        roof = {"kernel": "k_gemm_bf16_tc<256,2> (hidden layers, 1000x1000 per state)" if args.precision == "bf16" else "k_sgemm_bias_act",
                "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": (achieved / peak) if (achieved and peak) else None, "traffic": traffic, "traffic_note": traffic_note,
                "peak_source": peaks["_source"] + " (sustained bf16, kernel timed inside a long step)",
                # one CUDA-event interval brackets the 2*NB back-to-back hidden-layer launches of a network evaluation
                "launches_timed": calls["gemm_hidden"] * (2 * NB if args.precision == "bf16" else 1),
                "avg_launch_ms": stage["gemm_hidden"] / max(calls["gemm_hidden"] * (2 * NB if args.precision == "bf16" else 1), 1),
                "stage_ms_one_extra_step_rank0": {k: round(v, 3) for k, v in stage_split.items()}}
        line = {"metric": "nodes_generated_per_sec", "value": value, "unit": "nodes/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": tmax[0] / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": args.precision, "data": "synthetic", "config": workload_config(args, world),
                "clocks": clocks,
                "e2e": {"value": e2e_val, "unit": "nodes/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "api": "get_pathfind_from_arg('graph.10000B_0.6W') -> make_instances/add_instances/step() per iteration from host State objects"},
                "gpu_launches": int(tot[3]), "roofline": roof,
                "heur_evals": int(tot[2]), "nodes_generated": int(tot[0]), "wall_s_value_pass": wall_value_pass}
        if world == 1 and not args.no_kat:
            try:
                line["kat_solve"] = kat_solve_times(L, local)
            except Exception as e:  # noqa: BLE001  (extra information only; never fails the bench line)
                line["kat_solve"] = {"error": str(e)}
        if not args.no_cpu_baseline:
            import torch as _t
            threads = os.cpu_count() or 8
            sdt = {k: _t.from_numpy(np.array(v)) for k, v in sd.items()}
            n_cpu, s_cpu = cpu_reference_step(sdt, st_all[0], gl_all[0], args.ref_itrs, threads)
            line["cpu_baseline"] = {"value": n_cpu / s_cpu, "unit": "nodes/s", "cores": threads, "kind": "port",
                                    "sample": f"1 instance x {args.ref_itrs} iterations ({n_cpu} nodes, {s_cpu:.1f} s) of the same workload"}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
