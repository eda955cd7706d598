"""bench.py -- nodes generated / second of DeepXube's batch weighted A* / Q* hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

HEADLINE workload (config.workload; BASELINE.json configs[1], "C2"): cube3, `graph.10000B_0.6W` (batch weighted A*,
B=10000, W=0.6), heuristic `heurv,resnet_fc.1000H_4B_bn` with random-init weights (seed 0), 1000 synthetic scrambles
(random walks of 1000..10000 moves from the solved cube, seed 0).  A random-init heuristic never terminates a search, so
every instance runs a fixed number of iterations (--max-itrs, default 20; SURVEY.md 8d "C2 inputs").
One STEP = one PathFind holding G instances (--insts-per-step, default 8) advanced in lock-step for max_itrs
iterations: expand -> CLOSED filter -> heuristic network -> sort/merge/pop, all on the device.
  value   nodes generated / s with the start states already resident in HBM (timed region = the search
          iterations, CUDA events on the engine's stream, summed over the K steps, max over ranks)
  e2e     the same metric through the reference-facing API with HOST buffers: make_instances (H2D of the
          start states) + step loop + reading back every instance's status (D2H), wall clock, max over ranks
Multi-GPU: instances are independent, so ranks take disjoint scrambles (weak scaling, no data-path collective).

The same JSON line also carries, under "configs", the other BASELINE configs measured the same way (device-timed
nodes/s, heur_evals, the roofline of their dominant kernel, a bounded CPU sample of the same workload):
  c2_fp32   C2 with the fp32-accurate network (bit-exact-search mode)
  c4_qstar  cube3 batch weighted Q*, heurq_fixout, B=10000               (BASELINE configs[3])
  c3_np4    npuzzle.4 batch weighted A*, B=10000                          (BASELINE configs[2])
  c3_np5    npuzzle.5 batch weighted A*, B=1000                           (BASELINE configs[2])
and under "hda" BASELINE configs[4]: ONE cube3 search hash-distributed over the N ranks (B = 10000 * N, W=0.6), with
the peer-memory (p2p) and the NCCL exchange, per-phase device time and the per-rank load imbalance.

--impl reference: the UNMODIFIED reference (installed into baseline/_ref by __graft_entry__.build()) through its own
public API (get_pathfind_from_arg -> make_instances / step, deepxube/_solve.py:141-161) on the host cores, on a bounded
sample of C2 per step; the oracle port is the fallback when baseline/_ref is absent.
"""
import argparse
import json
import os
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, NB, ACTS = 1000, 4, 12
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def net_flops(in_dim: int, out_dim: int) -> float:
    """2 * (in * H + 2 * B * H^2 + H * out) FLOP per evaluated state (dense-equivalent)."""
    return 2.0 * (in_dim * H + 2 * NB * H * H + H * out_dim)


FLOP_PER_STATE = net_flops(324, 1)                   # 16.65 MFLOP (cube3, V)
FLOP_PER_STATE_HIDDEN_LAYER = 2.0 * H * H            # one hidden-layer GEMM launch, per state


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


def npuzzle_scrambles(dim: int, n: int, seed: int = 0, walk: int = 400):
    """n npuzzle.<dim> start states: random walk of `walk` blank moves from the goal (reverse walk from the goal,
    deepxube/domains/npuzzle.py:233-262); host-side input generation in plain numpy."""
    rng = np.random.default_rng(seed)
    goal = np.concatenate([np.arange(1, dim * dim), [0]]).astype(np.uint8)
    st = np.tile(goal, (n, 1))
    z = np.full(n, dim * dim - 1)
    rows = np.arange(n)
    for _ in range(walk):
        a = rng.integers(0, 4, size=n)
        r, c = z // dim, z % dim
        nr = np.clip(r + (a == 0) - (a == 1), 0, dim - 1)
        nc = np.clip(c + (a == 2) - (a == 3), 0, dim - 1)
        nz = nr * dim + nc
        st[rows, z] = st[rows, nz]
        st[rows, nz] = 0
        z = nz
    return st, np.tile(goal, (n, 1))


def random_weights(seed: int = 0, in_dim: int = 324, out_dim: int = 1):
    """state_dict of resnet_fc.1000H_4B_bn with nn.Linear's default init, eval-mode BN (mu 0, var 1)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def lin(o, i, p):
        b = 1.0 / np.sqrt(i)
        sd[p + ".weight"] = ((torch.rand(o, i, generator=g) * 2 - 1) * b).numpy()
        sd[p + ".bias"] = ((torch.rand(o, generator=g) * 2 - 1) * b).numpy()

    lin(H, in_dim, "heur.0")
    for b in range(NB):
        for j in range(2):
            p = f"heur.1.blocks.{b}.0.layers.{j}"
            lin(H, H, p + ".0")
            sd[p + ".1.weight"] = np.ones(H, np.float32)
            sd[p + ".1.bias"] = np.zeros(H, np.float32)
            sd[p + ".1.running_mean"] = np.zeros(H, np.float32)
            sd[p + ".1.running_var"] = np.ones(H, np.float32)
    lin(out_dim, H, "heur.2")
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
# CPU arms.  Only these functions touch baseline/_ref (the unmodified reference) or oracle/ (the port).
def _import_reference():
    """The unmodified reference from baseline/_ref (pip-installed there by __graft_entry__.build()), with stubs for the
    GUI / ASP packages it imports at module top (oracle/ref_shim.py).  None when it is not installed."""
    if not os.path.isdir(os.path.join(REF_DIR, "deepxube")):
        return None
    from oracle.ref_shim import install_stubs
    install_stubs()
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import deepxube
    if not os.path.abspath(deepxube.__file__).startswith(os.path.abspath(REF_DIR)):
        return None
    return deepxube


class CpuArm:
    """One CPU implementation of a config: the reference itself (kind "reference") or the oracle port (kind "port")."""

    def __init__(self, domain_str: str, fn: str, pathfind: str, sd: dict, threads: int):
        import torch
        self.domain_str, self.fn, self.pathfind_str, self.threads = domain_str, fn, pathfind, threads
        self.is_q = fn.startswith("heurq")
        torch.set_num_threads(threads)
        self.kind = "port"
        self.sd_t = {k: torch.from_numpy(np.array(v)) for k, v in sd.items()}
        if _import_reference() is not None:
            from deepxube.factories.domain_factory import get_domain_from_arg
            from deepxube.factories.pathfind_fns_factory import get_path_fns_nnet_par_dict
            sd_file = dict(self.sd_t)
            for k in list(sd_file):
                if k.endswith("running_var"):
                    sd_file[k.replace("running_var", "num_batches_tracked")] = torch.tensor(0)
            self._tmp = tempfile.NamedTemporaryFile(suffix=".pt", delete=False)
            torch.save(sd_file, self._tmp.name)
            self.domain, dname = get_domain_from_arg(domain_str)
            self.pfns, _ = get_path_fns_nnet_par_dict(self.domain, dname, [fn], torch.device("cpu"), nnet_files=[self._tmp.name])
            torch.set_num_threads(threads)           # (the reference's own default is 8 on CPU, nnet_utils.py:36-38)
            self.kind = "reference"

    def _ref_objects(self, state: np.ndarray, goal: np.ndarray):
        if self.domain_str == "cube3":
            from deepxube.domains.cube3 import Cube3Goal, Cube3State
            return Cube3State(state.astype(np.uint8)), Cube3Goal(goal.astype(np.uint8))
        from deepxube.domains.npuzzle import NPGoal, NPState
        return NPState(state.astype(np.uint8)), NPGoal(goal.astype(np.uint8))

    def sample(self, state: np.ndarray, goal: np.ndarray, iters: int):
        """make_instances + `iters` step()s of ONE instance (deepxube/_solve.py:146-160) -> (nodes generated, seconds)"""
        t0 = time.perf_counter()
        if self.kind == "reference":
            from deepxube.factories.pathfinding_factory import get_pathfind_from_arg
            pathfind = get_pathfind_from_arg(self.domain, self.pfns, self.pathfind_str)[0]
            s, g = self._ref_objects(state, goal)
            insts = pathfind.make_instances([s], [g], None, True)
            pathfind.add_instances(insts)
            for _ in range(iters):
                if insts[0].finished():
                    break
                pathfind.step()
            return int(insts[0].num_nodes_generated), time.perf_counter() - t0
        # fallback: the oracle port (torch-CPU network evaluating every generated child, like the reference)
        from oracle.domains_np import OracleDomain
        from oracle.nnet_torch import heurq_from_state_dict, heurv_from_state_dict
        from oracle.search import solve_one
        name, _, dim = self.domain_str.partition(".")
        dom = OracleDomain(name, int(dim) if dim else 0)
        depth = dom.input_info()[1]
        b, w = self.pathfind_str.split(".", 1)[1].split("_")
        f = (heurq_from_state_dict if self.is_q else heurv_from_state_dict)(self.sd_t, depth)
        r = solve_one(dom, state, goal, lambda s_, g_: f(s_.astype(np.int64)), self.is_q, int(b[:-1]), float(w[:-1]), max_itrs=iters)
        return int(r["num_nodes_generated"]), time.perf_counter() - t0

    def close(self):
        t = getattr(self, "_tmp", None)
        if t is not None:
            try:
                os.unlink(t.name)
            except OSError:
                pass


def cpu_sample_block(arm: CpuArm, state, goal, iters: int, what: str) -> dict:
    n, s = arm.sample(state, goal, iters)
    return {"value": n / s, "unit": "nodes/s", "cores": arm.threads, "kind": arm.kind,
            "sample": f"1 instance x {iters} iterations of {what} ({n} nodes, {s:.1f} s)"}


def run_reference(args, rank: int, world: int) -> None:
    """--impl reference: rank 0 times the reference's own CPU implementation of C2; the other ranks exit without work."""
    if rank != 0:
        return
    threads = os.cpu_count() or 8
    arm = CpuArm("cube3", "heurv,resnet_fc.1000H_4B_bn", "graph.10000B_0.6W", random_weights(0), threads)
    st, gl = make_scrambles(max(args.steps + args.warmup, 1), seed=0)
    iters = args.ref_itrs
    for i in range(args.warmup):                      # warm-up (thread pools, allocator) on a shorter ramp of the same search
        arm.sample(st[i], gl[i], max(iters - 1, 1))
    nodes, secs = 0, 0.0
    for i in range(args.steps):
        n, s = arm.sample(st[args.warmup + i], gl[args.warmup + i], iters)
        nodes += n
        secs += s
    arm.close()
    val = nodes / secs
    per = nodes // max(args.steps, 1)
    sample = (f"1 instance x {iters} iterations ({per} nodes: the ramp 1 -> 12 -> 144 -> 1728 -> 10000 popped nodes) per step, "
              f"{'unmodified reference from baseline/_ref' if arm.kind == 'reference' else 'oracle port (baseline/_ref absent)'}, "
              f"torch-CPU network on {threads} threads")
    cfg = workload_config(args, 1)
    cfg.update({"instances_per_step": 1, "max_itrs": iters, "parallelism": f"host CPU, {threads} threads",
                "note": "bounded sample of the headline workload: same domain / pathfind / network / scrambles, one instance and "
                        f"{iters} iterations per step instead of {args.insts_per_step} x {args.max_itrs}"})
    line = {"impl": "reference", "metric": "nodes_generated_per_sec", "value": val, "unit": "nodes/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(args.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": val, "unit": "nodes/s", "cores": threads, "kind": arm.kind, "sample": sample},
            "e2e": {"value": val, "unit": "nodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
def kat_solve_times(L, device: int) -> dict:
    """Solve time / instance (the second half of BASELINE's metric) on the reference's own known-answer set:
    tutorial/cube3 `hard` instances with the trained resnet_fc.200H_2B_bn heuristic, graph.100B_0.1W
    (fixtures under tests/golden/, generated from the unmodified reference by tools/gen_golden.py).  Timed like
    deepxube/_solve.py:150-161 (instance creation + every step), one instance at a time."""
    g = os.path.join(ROOT, "tests", "golden")
    k = np.load(os.path.join(g, "kat_cube3.npz"))
    z = np.load(os.path.join(g, "nnet.npz"))
    sd = {key[len("cube3v/"):]: z[key] for key in z.files if key.startswith("cube3v/")}
    out = {}
    for prec, mode in (("fp32", L.DX_HEUR_NET_FP32), ("bf16", L.DX_HEUR_NET_BF16)):
        eng = L.Engine("cube3", 0, "graph_v", mode, max_instances=1, max_pop_total=100, max_nodes=600000, device=device)
        blob, info = L.pack_state_dict(sd)
        eng.load_weights(54, 6, info["res_dim"], info["num_blocks"], 1, True, blob)
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
        dev = (np.array(nodes) - k["hard_nodes"]) / k["hard_nodes"]
        out[prec] = {"instances": 10, "mean_solve_time_s": float(np.mean(secs)), "max_solve_time_s": float(np.max(secs)),
                     "nodes_per_s": float(np.sum(nodes) / np.sum(secs)), "mean_path_cost": float(np.mean(costs)),
                     "path_costs_equal_reference": bool(np.array_equal(np.array(costs), k["hard_path_costs"])),
                     "nodes_generated_equal_reference": bool(np.array_equal(np.array(nodes), k["hard_nodes"])),
                     "nodes_generated_deviation_min_max": [float(dev.min()), float(dev.max())]}
    out["workload"] = "tutorial/cube3 hard set (10 instances), heurv resnet_fc.200H_2B_bn trained weights, graph.100B_0.1W"
    out["reference_mean_path_cost"] = float(np.mean(k["hard_path_costs"]))
    return out


def bf16_deviation_c2(L, device: int, st_all, gl_all, sd, iters: int = 8, insts: int = 2) -> dict:
    """north_star: "bf16-mode solution costs and nodes generated are reported with their deviation".  A random-init network never
    terminates a C2 search, so the deviation is reported on what the search DOES with the values: the same instances are advanced in
    lock-step by an fp32-accurate engine and a bf16 engine; per iteration, the fraction of the bf16 engine's popped nodes (the nodes
    it expands next) that the fp32 engine pops too, and the heuristic error on the fp32 engine's popped states."""
    B = 10000
    blob, _ = L.pack_state_dict(sd)
    engs = {}
    for name, mode in (("fp32", L.DX_HEUR_NET_FP32), ("bf16", L.DX_HEUR_NET_BF16)):
        e = L.Engine("cube3", 0, "graph_v", mode, max_instances=insts, max_pop_total=insts * B,
                     max_nodes=int(insts * B * ACTS * (iters + 1)) + 1024, device=device)
        e.load_weights(54, 6, H, NB, 1, True, blob)
        e.add_instances(st_all[:insts], gl_all[:insts], B, 0.6)
        engs[name] = e
    overlap, herr_max, herr_mean, gen = [], 0.0, [], {}
    for it in range(iters):
        for e in engs.values():
            e.run(1)
        same, tot = 0, 0
        for j in range(insts):
            a, _, _ = engs["fp32"].curr_nodes(j, B + 8)
            b, _, _ = engs["bf16"].curr_nodes(j, B + 8)
            sa = {r.tobytes() for r in a}
            same += sum(1 for r in b if r.tobytes() in sa)
            tot += len(b)
            if j == 0 and len(a):
                hf = engs["fp32"].heur_eval(a[:4096], gl_all[:1].repeat(min(len(a), 4096), axis=0), 1)[:, 0]
                hb = engs["bf16"].heur_eval(a[:4096], gl_all[:1].repeat(min(len(a), 4096), axis=0), 1)[:, 0]
                d = np.abs(hf.astype(np.float64) - hb)
                herr_max = max(herr_max, float(d.max()))
                herr_mean.append(float(d.mean()))
        overlap.append(same / max(tot, 1))
    for name, e in engs.items():
        gen[name] = int(sum(s.num_nodes_generated for s in e.status()))
        e.close()
    return {"workload": f"{insts} C2 instances x {iters} iterations, fp32-accurate engine vs bf16 engine in lock-step",
            "popped_set_overlap_per_iteration": [round(x, 4) for x in overlap],
            "heuristic_abs_error_max": herr_max, "heuristic_abs_error_mean": float(np.mean(herr_mean)) if herr_mean else None,
            "nodes_generated": gen, "nodes_generated_deviation": (gen["bf16"] - gen["fp32"]) / max(gen["fp32"], 1),
            "note": "solution costs cannot deviate on this workload (random-init weights: no search terminates); on the trained-network KAT "
                    "set (kat_solve) bf16 costs equal the reference's 10/10"}


def workload_config(args, world: int) -> dict:
    return {"workload": "cube3 batch weighted A* graph.10000B_0.6W, heurv resnet_fc.1000H_4B_bn random-init, "
                        "1000 synthetic scrambles (BASELINE.json configs[1])",
            "domain": "cube3", "pathfind": "graph.10000B_0.6W", "heuristic": "heurv,resnet_fc.1000H_4B_bn",
            "instances_per_step": args.insts_per_step, "max_itrs": args.max_itrs, "num_scrambles": 1000,
            "parallelism": f"instance-sharded x{world}", "l2_policy": "inputs larger than L2 (per-iteration working set "
            "~0.6 GB of activations per 120k-state batch >> 126 MB L2); no flush needed"}


# ---------------------------------------------------------------------------------------------------
# The other BASELINE configs, each measured like the headline's device-timed `value`.
SIDE_CONFIGS = {
    # name: (domain, dim, algo, fn, B, instances per step, iterations, precision, CPU-sample iterations, description)
    "c2_fp32": ("cube3", 0, "graph_v", "heurv,resnet_fc.1000H_4B_bn", 10000, 8, 20, "fp32", 0,
                "C2 with the fp32-accurate network (the mode whose searches are bit-exact against the reference)"),
    "c4_qstar": ("cube3", 0, "graph_q", "heurq_fixout,resnet_fc.1000H_4B_bn", 10000, 8, 40, "bf16", 8,
                 "cube3 batch weighted Q* graph.10000B_0.6W, heurq_fixout (BASELINE.json configs[3])"),
    "c3_np4": ("npuzzle", 4, "graph_v", "heurv,resnet_fc.1000H_4B_bn", 10000, 16, 20, "bf16", 16,
               "npuzzle.4 (15-puzzle) batch weighted A* graph.10000B_0.6W (BASELINE.json configs[2])"),
    "c3_np5": ("npuzzle", 5, "graph_v", "heurv,resnet_fc.1000H_4B_bn", 1000, 64, 20, "bf16", 17,
               "npuzzle.5 (24-puzzle) batch weighted A* graph.1000B_0.6W (BASELINE.json configs[2])"),
}


def run_side_config(L, name: str, rank: int, world: int, local: int, steps: int, warmup: int, peaks: dict, cpu: bool) -> dict:
    dname, dim, algo, fn, B, G, iters, prec, cpu_iters, desc = SIDE_CONFIGS[name]
    is_q = algo == "graph_q"
    S, A, in_count, depth = L.domain_info(dname, dim)
    in_dim, out_dim = in_count * depth, (A if is_q else 1)
    sd = random_weights(0, in_dim, out_dim)
    blob, _ = L.pack_state_dict(sd)
    n_pool = max(256, G * (steps + warmup + 1) * world)
    st_all, gl_all = make_scrambles(n_pool, 0) if dname == "cube3" else npuzzle_scrambles(dim, n_pool, 0)
    per_iter = G * B * (1 if is_q else A)                         # nodes stored per iteration, at most
    mode = L.DX_HEUR_NET_BF16 if prec == "bf16" else L.DX_HEUR_NET_FP32
    eng = L.Engine(dname, dim, algo, mode, max_instances=G, max_pop_total=G * B, max_nodes=int(per_iter * (iters + 1)) + 4096, device=local)
    eng.load_weights(in_count, depth, H, NB, out_dim, True, blob)

    def pick(i):
        idx = (((i * world + rank) * G) + np.arange(G)) % n_pool
        return st_all[idx], gl_all[idx]

    def one(i):
        s, g = pick(i)
        eng.reset()
        eng.add_instances(s, g, B, 0.6)
        eng.run(iters)
        return eng.last_run_ms(), eng.counters()

    eng.set_profiling(2)
    for w in range(warmup):
        one(w)
    eng.set_profiling(2)
    ms, nodes, evals = 0.0, 0, 0
    for k in range(steps):
        m, c = one(warmup + k)
        ms += m
        nodes += c.children_generated
        evals += c.heur_evals
    stage, calls = eng.stage_ms(), eng.stage_calls()
    n_hidden = int(eng.counters().hidden_gemm_launches)        # H x H GEMM launches per evaluation inside the timed interval
    eng.set_profiling(1)
    one(warmup + steps)
    split = eng.stage_ms()
    open_total = int(sum(s.open_size for s in eng.status()))
    eng.set_profiling(0)
    eng.close()
    hidden_flop = FLOP_PER_STATE_HIDDEN_LAYER * n_hidden * evals
    gms = stage["gemm_hidden"]
    achieved = hidden_flop / (gms / 1e3) / 1e12 if gms > 0 else None
    peak = peaks["bf16_tflops_sustained"]
    tot = sum(split[k] for k in ("expand", "filt", "scatter", "heur", "sort", "pushpop", "book"))
    out = {"workload": desc, "domain": dname + (f".{dim}" if dim else ""), "pathfind": f"graph.{B}B_0.6W", "heuristic": fn,
           "precision": prec, "instances_per_step": G, "max_itrs": iters, "steps": steps, "warmup": warmup,
           "value": nodes / (ms / 1e3), "unit": "nodes/s (this rank, device-timed)", "ms_per_step": ms / max(steps, 1),
           "nodes_generated": int(nodes), "heur_evals": int(evals), "open_entries_at_end": open_total,
           "net_tflops_whole_step": net_flops(in_dim, out_dim) * evals / (ms / 1e3) / 1e12,
           "roofline": {"kernel": "hidden-layer GEMMs (2*H^2 FLOP per state and layer; fp32 mode: 3 fp16 MMAs per algorithmic MAC)"
                        if prec == "fp32" else "k_gemm_bf16_tc hidden layers (2*H^2 FLOP per state and layer)",
                        "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                        "frac": (achieved / peak) if achieved else None, "traffic": None,
                        "peak_source": peaks["_source"] + " (sustained bf16)",
                        "launches_timed": int(calls["gemm_hidden"] * n_hidden),
                        "avg_launch_ms": gms / max(calls["gemm_hidden"] * n_hidden, 1)},
           "stage_share_one_extra_step": {k: round(split[k] / tot, 4) for k in ("expand", "filt", "scatter", "heur", "sort", "pushpop", "book")} if tot > 0 else None,
           "stage_ms_one_extra_step": {k: round(v, 3) for k, v in split.items()}}
    if cpu and cpu_iters > 0:
        try:
            arm = CpuArm(dname + (f".{dim}" if dim else ""), fn, f"graph.{B}B_0.6W", sd, os.cpu_count() or 8)
            out["cpu_baseline"] = cpu_sample_block(arm, st_all[0], gl_all[0], cpu_iters, "the same workload")
            arm.close()
        except Exception as e:  # noqa: BLE001  (side information; never fails the bench line)
            out["cpu_baseline"] = {"error": repr(e)}
    return out


def run_hda(rank: int, world: int, local: int, dist, iters: int, reps: int) -> dict:
    """BASELINE configs[4]: ONE cube3 search hash-distributed over the ranks, B = 10000 per GPU, W = 0.6, bf16."""
    import torch
    import deepxube_b200  # noqa: F401
    from deepxube_b200.factories.domain_factory import get_domain_from_arg
    from deepxube_b200.factories.pathfind_fns_factory import get_path_fns_nnet_par_dict
    from deepxube_b200.hda import HdaSearch
    domain, dname = get_domain_from_arg("cube3")
    pfns, _ = get_path_fns_nnet_par_dict(domain, dname, ["heurv,resnet_fc.1000H_4B_bn"], precision="bf16")
    pfns.heurv.nnet.load_state_dict(random_weights(0))
    st, gl = make_scrambles(1, seed=1)
    per_gpu = 10000
    B = per_gpu * world
    out = {"workload": f"one cube3 search, HDA*-style (owner = hash(state) mod {world}), graph.{B}B_0.6W, heurv resnet_fc.1000H_4B_bn "
                       f"random-init bf16, {iters} iterations (BASELINE.json configs[4])", "n_gpus": world, "global_batch": B,
           "iterations": iters, "reps": reps, "exchanges": {}}
    arms = ("p2p", "nccl") if (world > 1 and dist is not None) else ("p2p",)
    for exch in arms:
        h = HdaSearch("cube3", 0, world, B, 0.6, nnet=pfns.heurv.nnet, precision="bf16",
                      max_nodes_per_rank=int(per_gpu * 12 * (iters + 2) * 1.5), dist=dist, device=local, exchange=exch)
        h.solve(st[0], gl[0], max_itrs=iters)                      # warm-up
        h.timing = True
        best, nodes = None, 0
        for _ in range(reps):
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = h.solve(st[0], gl[0], max_itrs=iters)
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            dt = time.perf_counter() - t0
            nodes = r["num_nodes_generated"]
            best = dt if best is None else min(best, dt)
        tmax = torch.tensor([best] + [h.phase_ms[k] for k in ("pack", "exchange", "receive", "all_gather", "reduce")],
                            device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tl = tmax.tolist()
        n_it = max(h.timed_iters, 1)
        out["exchanges"][h.exchange if exch == "p2p" else exch] = {
            "nodes_generated": int(nodes), "seconds_best_of_reps_max_over_ranks": tl[0], "nodes_per_s": nodes / tl[0],
            "ms_per_iteration": 1e3 * tl[0] / max(iters, 1),
            "phase_ms_per_iteration_max_over_ranks": {k: tl[1 + j] / n_it for j, k in enumerate(("pack", "exchange", "receive", "all_gather", "reduce"))},
            "phase_note": "CUDA events on the iteration's stream: pack = expand + bucket by owner + pack (+ NVLink stores into the "
                          "owners' buffers for p2p); exchange = symmetric-memory barrier (p2p) or all_to_all_single (nccl); receive = unpack + "
                          "CLOSED filter + network + sort/merge/pop; all_gather = 8 doubles per rank; the rest of ms_per_iteration is the "
                          "host's 64-byte summary read + launch overhead",
            "received_children_per_iteration": {"max_over_ranks_mean": h.recv_max_sum / n_it, "mean_over_ranks_mean": h.recv_mean_sum / n_it,
                                                "imbalance": (h.recv_max_sum / h.recv_mean_sum) if h.recv_mean_sum > 0 else None}}
        h.close()
    return out


# ---------------------------------------------------------------------------------------------------
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="b200", choices=["b200", "reference"])
    ap.add_argument("--insts-per-step", type=int, default=8)
    ap.add_argument("--max-itrs", type=int, default=20)
    ap.add_argument("--ref-itrs", type=int, default=5, help="--impl reference: iterations per step (one instance; 134 340 nodes)")
    ap.add_argument("--cpu-sample-itrs", type=int, default=12,
                    help="iterations of the bounded CPU sample inside the GPU arm's line (one instance; 12 -> 974 k nodes, ~8-10 s)")
    ap.add_argument("--precision", type=str, default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kat", action="store_true", help="skip the solve-time-per-instance report on the trained-net KAT set")
    ap.add_argument("--no-configs", action="store_true", help="skip the side configs (c2_fp32, c4_qstar, c3_np4, c3_np5)")
    ap.add_argument("--configs", type=str, default="c2_fp32,c4_qstar,c3_np4,c3_np5")
    ap.add_argument("--config-steps", type=int, default=3)
    ap.add_argument("--no-hda", action="store_true", help="skip the HDA* block (BASELINE configs[4])")
    ap.add_argument("--hda-itrs", type=int, default=20)
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
    eng.load_weights(54, 6, H, NB, 1, True, blob)

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
    n_hidden = int(eng.counters().hidden_gemm_launches)   # H x H GEMM launches per network evaluation (7 of the 8 hidden layers when
    #                                                       the first one is computed by the input-layer launch)
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

    # ---- the other BASELINE configs (every rank runs them on its own scrambles; rank 0's numbers are printed) ----
    peaks = load_peaks()
    side = {}
    if not args.no_configs:
        for name in [x for x in args.configs.split(",") if x]:
            try:
                side[name] = run_side_config(L, name, rank, world, local, args.config_steps, 1, peaks,
                                             cpu=(rank == 0 and world == 1 and not args.no_cpu_baseline))
            except Exception as e:  # noqa: BLE001  (side information; never fails the bench line)
                side[name] = {"error": repr(e)}
            if dist is not None:
                t = torch.tensor([side[name].get("value", 0.0)], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
                side[name]["value_all_ranks"] = float(t.item())
    hda = None
    if not args.no_hda:
        try:
            hda = run_hda(rank, world, local, dist, args.hda_itrs, 3)
        except Exception as e:  # noqa: BLE001
            hda = {"error": repr(e)}
            if dist is not None:
                raise                              # a rank that left the collectives cannot rejoin the final reduction

    # ---- reduce over ranks: max time, sum of work -----------------------------------------------------------
    tot = torch.tensor([float(nodes), float(e2e_nodes), float(evals), float(launches)], device="cuda", dtype=torch.float64)
    tmax = torch.tensor([dev_ms, e2e_s, stage["gemm_hidden"]], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tot, tmax = tot.tolist(), tmax.tolist()

    if rank == 0:
        value = tot[0] / (tmax[0] / 1e3)
        e2e_val = tot[1] / tmax[1]
        # roofline of the dominant kernel (hidden-layer tcgen05 GEMM, 8 launches per network evaluation):
        # algorithmic FLOPs of this rank's launches / their summed CUDA-event durations
        hidden_flop = FLOP_PER_STATE_HIDDEN_LAYER * n_hidden * evals
        achieved = hidden_flop / (stage["gemm_hidden"] / 1e3) / 1e12 if stage["gemm_hidden"] > 0 else None
        peak = peaks["bf16_tflops_sustained"]
        traffic, traffic_note = None, None
        tf = os.path.join(ROOT, "profiles", "r2k_gemm_traffic.json")          # (round-2 capture: the launches of the last iteration of a step)
        if args.precision == "bf16" and os.path.exists(tf):
            tj = json.load(open(tf))
            traffic = tj["traffic_per_launch_avg_bytes"]
            traffic_note = (f"dram__bytes_read.sum + dram__bytes_write.sum per launch from one ncu --set full capture of a "
                            f"{tj['rows_in_launch']}-state launch (profiles/r2k_gemm_traffic.json); algorithmic bytes of that launch: "
                            f"{tj['algorithmic_bytes']['hidden']} (plain) / {tj['algorithmic_bytes']['hidden_residual']} (with residual)")
        n_l = calls["gemm_hidden"] * n_hidden
        roof = {"kernel": "k_gemm_bf16_tc<256,2> (hidden layers, 1000x1000 per state)" if args.precision == "bf16"
                else "k_gemm_bf16_tc<256,2,*,split> (hidden layers, fp16 hi/lo split: 3 MMAs per algorithmic MAC)",
                "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": (achieved / peak) if (achieved and peak) else None, "traffic": traffic, "traffic_note": traffic_note,
                "peak_source": peaks["_source"] + " (sustained bf16, kernel timed inside a long step)",
                # one CUDA-event interval brackets the 2*NB back-to-back hidden-layer launches of a network evaluation
                "launches_timed": n_l, "avg_launch_ms": stage["gemm_hidden"] / max(n_l, 1),
                "whole_step_frac": FLOP_PER_STATE * evals / (dev_ms / 1e3) / 1e12 / peak,
                "stage_ms_one_extra_step_rank0": {k: round(v, 3) for k, v in stage_split.items()}}
        line = {"metric": "nodes_generated_per_sec", "value": value, "unit": "nodes/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": tmax[0] / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": args.precision, "data": "synthetic", "config": workload_config(args, world),
                "clocks": clocks,
                "e2e": {"value": e2e_val, "unit": "nodes/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "api": "get_pathfind_from_arg('graph.10000B_0.6W') -> make_instances/add_instances/step() per iteration from host State objects"},
                "gpu_launches": int(tot[3]), "roofline": roof,
                "heur_evals": int(tot[2]), "nodes_generated": int(tot[0]), "wall_s_value_pass": wall_value_pass}
        if side:
            line["configs"] = side
        if hda is not None:
            line["hda"] = hda
        if world == 1 and not args.no_kat:
            try:
                line["kat_solve"] = kat_solve_times(L, local)
            except Exception as e:  # noqa: BLE001  (extra information only; never fails the bench line)
                line["kat_solve"] = {"error": str(e)}
            try:
                line["bf16_deviation"] = bf16_deviation_c2(L, local, st_all, gl_all, sd)
            except Exception as e:  # noqa: BLE001
                line["bf16_deviation"] = {"error": repr(e)}
        if not args.no_cpu_baseline and world == 1:
            try:
                arm = CpuArm("cube3", "heurv,resnet_fc.1000H_4B_bn", "graph.10000B_0.6W", sd, os.cpu_count() or 8)
                line["cpu_baseline"] = cpu_sample_block(arm, st_all[0], gl_all[0], args.cpu_sample_itrs, "the headline workload")
                arm.close()
            except Exception as e:  # noqa: BLE001
                line["cpu_baseline"] = {"error": repr(e)}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
