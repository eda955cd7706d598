"""ctypes binding of libdxb200.so (the C ABI declared in include/dxb200.h).

There is NO CPU fallback: if the shared library is missing or no CUDA device is visible, every entry
point raises.  ``load()`` builds nothing -- run ``python -m deepxube_b200.build`` (or
``__graft_entry__.build()``) first.
"""
import ctypes as C
import os
from typing import List, Optional, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DXB200_LIB") or os.path.join(HERE, "libdxb200.so")   # DXB200_LIB: A/B builds

DX_DOMAIN = {"cube3": 0, "npuzzle": 1, "grid": 2, "lightsout": 3}
DX_ALGO_GRAPH_V, DX_ALGO_GRAPH_Q, DX_ALGO_BEAM_V, DX_ALGO_BEAM_Q = 0, 1, 2, 3
DX_BACKUP_MODE = {"bellman": 0, "ub_solns": 1, "lhbl": 2}
DX_ALGO = {"graph_v": DX_ALGO_GRAPH_V, "graph_q": DX_ALGO_GRAPH_Q, "beam_v": DX_ALGO_BEAM_V, "beam_q": DX_ALGO_BEAM_Q}
DX_HEUR_ZERO, DX_HEUR_HOST, DX_HEUR_NET_FP32, DX_HEUR_NET_BF16 = 0, 1, 2, 3
DX_HDA_LOCAL_DOUBLES = 8       # per-rank scalars of one HDA* iteration (include/dxb200.h)
STAGE_NAMES = ["expand", "filt", "scatter", "heur", "sort", "pushpop", "book", "gemm_hidden", "gemm_input"]


class DxError(RuntimeError):
    pass


class dx_config(C.Structure):
    _fields_ = [("device", C.c_int32), ("domain", C.c_int32), ("domain_dim", C.c_int32), ("algo", C.c_int32),
                ("heur_mode", C.c_int32), ("max_instances", C.c_int32), ("max_pop_total", C.c_int32),
                ("max_backups", C.c_int32), ("max_nodes", C.c_int64), ("max_open", C.c_int64), ("reserved", C.c_int32 * 8)]


class dx_net_desc(C.Structure):
    _fields_ = [("in_count", C.c_int32), ("one_hot_depth", C.c_int32), ("res_dim", C.c_int32), ("num_blocks", C.c_int32),
                ("out_dim", C.c_int32), ("batch_norm", C.c_int32), ("act_depth", C.c_int32)]


class dx_inst_status(C.Structure):
    _fields_ = [("lb", C.c_double), ("ub", C.c_double), ("num_nodes_generated", C.c_int64), ("open_size", C.c_int64),
                ("goal_node", C.c_int64), ("itr", C.c_int32), ("finished", C.c_int32), ("num_curr", C.c_int32),
                ("reserved", C.c_int32)]


class dx_hda_local(C.Structure):
    _fields_ = [("k_popped", C.c_int64), ("f_first", C.c_double), ("ub", C.c_double), ("goal_local", C.c_int64),
                ("n_kept", C.c_int64), ("open_size", C.c_int64), ("children_sent", C.c_int64)]


class dx_counters(C.Structure):
    _fields_ = [("nodes_stored", C.c_int64), ("heur_evals", C.c_int64), ("children_generated", C.c_int64),
                ("iterations", C.c_int64), ("kernel_launches", C.c_int64), ("backups_logged", C.c_int64),
                ("hidden_gemm_launches", C.c_int64), ("reserved", C.c_int64 * 1)]


_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int32)
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_engp = C.c_void_p

# name -> (restype, argtypes); every exported symbol of include/dxb200.h
SIGNATURES = {
    "dx_last_error": (C.c_char_p, []),
    "dx_version": (C.c_int, []),
    "dx_engine_create": (C.c_int, [C.POINTER(dx_config), C.POINTER(_engp)]),
    "dx_engine_destroy": (C.c_int, [_engp]),
    "dx_engine_reset": (C.c_int, [_engp]),
    "dx_load_weights": (C.c_int, [_engp, C.POINTER(dx_net_desc), _f32p, C.c_int64]),
    "dx_add_instances": (C.c_int, [_engp, _u8p, _u8p, C.c_int32, _i32p, _f64p, _f32p]),
    "dx_remove_instances": (C.c_int, [_engp, _i32p, C.c_int32]),
    "dx_run": (C.c_int, [_engp, C.c_int32, _i32p]),
    "dx_step_begin": (C.c_int, [_engp, C.POINTER(C.c_int64)]),
    "dx_get_pending": (C.c_int, [_engp, _u8p, _u8p, C.c_int64]),
    "dx_step_end": (C.c_int, [_engp, _f32p, C.c_int64]),
    "dx_num_instances": (C.c_int, [_engp]),
    "dx_all_finished": (C.c_int, [_engp, _i32p]),
    "dx_get_status": (C.c_int, [_engp, C.POINTER(dx_inst_status), C.c_int32]),
    "dx_get_counters": (C.c_int, [_engp, C.POINTER(dx_counters)]),
    "dx_drain_backups": (C.c_int, [_engp, C.c_int32, _u8p, _i32p, C.POINTER(C.c_double), _i32p, C.c_int64, C.POINTER(C.c_int64)]),
    "dx_get_path": (C.c_int, [_engp, C.c_int32, _i32p, _u8p, C.c_int32, _i32p]),
    "dx_get_curr_nodes": (C.c_int, [_engp, C.c_int32, _u8p, _f32p, _f32p, C.c_int32, _i32p]),
    "dx_domain_info": (C.c_int, [C.c_int32, C.c_int32, _i32p, _i32p, _i32p, _i32p]),
    "dx_expand": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _u8p, C.c_int64, _u8p]),
    "dx_next_state": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _u8p, _i32p, C.c_int64, _u8p]),
    "dx_is_solved": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _u8p, _u8p, C.c_int64, _u8p]),
    "dx_heur_eval": (C.c_int, [_engp, _u8p, _u8p, C.c_int64, _f32p]),
    "dx_vi_targets": (C.c_int, [_engp, _u8p, _u8p, _u8p, C.c_int64, C.POINTER(C.c_double)]),
    "dx_row_bytes": (C.c_int, [_engp, _i32p]),
    "dx_bench_heur": (C.c_int, [_engp, C.c_int64, C.c_int32, _f32p]),
    "dx_set_profiling": (C.c_int, [_engp, C.c_int32]),
    "dx_get_stage_ms": (C.c_int, [_engp, _f32p, C.c_int32]),
    "dx_get_stage_calls": (C.c_int, [_engp, C.POINTER(C.c_int64), C.c_int32]),
    "dx_last_run_ms": (C.c_int, [_engp, _f32p]),
    "dx_debug_tc_profile": (C.c_int, [_engp, C.POINTER(C.c_uint64)]),
    "dx_backup_deepest": (C.c_int, [_engp, _u8p, _i32p, C.c_int32]),
    "dx_hda_config": (C.c_int, [_engp, C.c_int32, C.c_int32]),
    "dx_hda_record_bytes": (C.c_int, [_engp, _i32p]),
    "dx_hda_owner": (C.c_int, [_engp, _u8p, _i32p]),
    "dx_hda_add_instance": (C.c_int, [_engp, _u8p, _u8p, C.c_int32, C.c_double]),
    "dx_hda_expand": (C.c_int, [_engp, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
    "dx_hda_absorb": (C.c_int, [_engp, C.c_void_p, C.c_int64, C.POINTER(dx_hda_local)]),
    "dx_hda_finish": (C.c_int, [_engp, C.c_int64, C.c_double, C.c_double, C.c_int64, _i32p]),
    "dx_hda_node": (C.c_int, [_engp, C.c_int64, _u8p, C.POINTER(C.c_uint32), _i32p, _f32p]),
    "dx_engine_set_stream": (C.c_int, [_engp, C.c_void_p]),
    "dx_hda_send": (C.c_int, [_engp, C.c_void_p, C.c_int64]),
    "dx_hda_send_peers": (C.c_int, [_engp, C.POINTER(C.c_uint64), C.c_int64]),
    "dx_hda_receive": (C.c_int, [_engp, C.c_void_p, C.c_int64, C.c_void_p]),
    "dx_hda_reduce": (C.c_int, [_engp, C.c_void_p, C.c_void_p]),
}

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DxError(f"{LIB_PATH} not found: build it with `python -m deepxube_b200.build` "
                          "(deepxube_b200 has no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().dx_last_error()
        raise DxError(f"libdxb200 error {rc}: {msg.decode() if msg else ''}")


def _p(arr: Optional[np.ndarray], typ):
    return None if arr is None else arr.ctypes.data_as(typ)


def _u8(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint8)


def current_device() -> int:
    """cuda:LOCAL_RANK under torchrun, else DXB200_DEVICE or 0."""
    return int(os.environ.get("DXB200_DEVICE", os.environ.get("LOCAL_RANK", "0")))


def domain_info(domain: str, dim: int) -> Tuple[int, int, int, int]:
    """(state_bytes, num_actions, nn_in_count, one_hot_depth)"""
    s, a, i, d = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
    check(load().dx_domain_info(DX_DOMAIN[domain], dim, C.byref(s), C.byref(a), C.byref(i), C.byref(d)))
    return s.value, a.value, i.value, d.value


def expand(domain: str, dim: int, states: np.ndarray, device: Optional[int] = None) -> np.ndarray:
    states = _u8(states)
    sb, na, _, _ = domain_info(domain, dim)
    assert states.ndim == 2 and states.shape[1] == sb, states.shape
    out = np.empty((states.shape[0] * na, sb), np.uint8)
    check(load().dx_expand(current_device() if device is None else device, DX_DOMAIN[domain], dim, _p(states, _u8p),
                           states.shape[0], _p(out, _u8p)))
    return out


def next_state(domain: str, dim: int, states: np.ndarray, actions, device: Optional[int] = None) -> np.ndarray:
    states = _u8(states)
    acts = np.ascontiguousarray(actions, dtype=np.int32)
    sb = domain_info(domain, dim)[0]
    assert states.ndim == 2 and states.shape[1] == sb and acts.shape == (states.shape[0],)
    out = np.empty_like(states)
    check(load().dx_next_state(current_device() if device is None else device, DX_DOMAIN[domain], dim, _p(states, _u8p),
                               _p(acts, _i32p), states.shape[0], _p(out, _u8p)))
    return out


def is_solved(domain: str, dim: int, states: np.ndarray, goals: np.ndarray, device: Optional[int] = None) -> np.ndarray:
    states, goals = _u8(states), _u8(goals)
    assert states.shape == goals.shape
    out = np.empty(states.shape[0], np.uint8)
    check(load().dx_is_solved(current_device() if device is None else device, DX_DOMAIN[domain], dim, _p(states, _u8p),
                              _p(goals, _u8p), states.shape[0], _p(out, _u8p)))
    return out.astype(bool)


class Engine:
    """Thin object wrapper over a dx_engine handle; numpy in, numpy out."""

    def __init__(self, domain: str, dim: int = 0, algo: str = "graph_v", heur_mode: int = DX_HEUR_ZERO,
                 max_instances: int = 1, max_pop_total: int = 1, max_nodes: int = 1 << 20, max_open: int = 0,
                 device: Optional[int] = None, max_backups: int = 0):
        """max_backups > 0 (graph_v): evaluate every generated child before the CLOSED filter and log the Bellman backup of
        every expanded node (see drain_backups)"""
        lib = load()
        self.domain, self.dim = domain, dim
        self.max_backups = max_backups
        self.state_bytes, self.num_actions, self.in_count, self.depth = domain_info(domain, dim)
        self.is_q = algo in ("graph_q", "beam_q")
        self.algo = algo
        self.heur_mode = heur_mode
        cfg = dx_config(device=current_device() if device is None else device, domain=DX_DOMAIN[domain], domain_dim=dim,
                        algo=DX_ALGO[algo], heur_mode=heur_mode,
                        max_instances=max_instances, max_pop_total=max_pop_total, max_nodes=max_nodes, max_open=max_open,
                        max_backups=max_backups)
        self._h = _engp()
        check(lib.dx_engine_create(C.byref(cfg), C.byref(self._h)))
        self._lib = lib

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.dx_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self) -> None:
        check(self._lib.dx_engine_reset(self._h))

    def load_weights(self, in_count: int, depth: int, res_dim: int, num_blocks: int, out_dim: int, batch_norm: bool,
                     blob: np.ndarray, act_depth: int = 0) -> None:
        """act_depth > 0: action-as-input Q network (heurq_in): the action index is one more input of that one-hot depth"""
        blob = np.ascontiguousarray(blob, dtype=np.float32)
        desc = dx_net_desc(in_count, depth, res_dim, num_blocks, out_dim, 1 if batch_norm else 0, act_depth)
        check(self._lib.dx_load_weights(self._h, C.byref(desc), _p(blob, _f32p), blob.size))

    def add_instances(self, states: np.ndarray, goals: np.ndarray, batch_sizes=None, weights=None, root_vals=None) -> None:
        states, goals = _u8(states), _u8(goals)
        n = states.shape[0]
        assert states.shape == (n, self.state_bytes) and goals.shape == (n, self.state_bytes), (states.shape, goals.shape)
        b = None if batch_sizes is None else np.ascontiguousarray(np.broadcast_to(batch_sizes, (n,)), dtype=np.int32)
        w = None if weights is None else np.ascontiguousarray(np.broadcast_to(weights, (n,)), dtype=np.float64)
        r = None if root_vals is None else np.ascontiguousarray(root_vals, dtype=np.float32)
        check(self._lib.dx_add_instances(self._h, _p(states, _u8p), _p(goals, _u8p), n, _p(b, _i32p), _p(w, _f64p), _p(r, _f32p)))

    def remove_instances(self, slots) -> None:
        """The listed instance slots stop searching (PathFind.remove_instances); their nodes stay readable until reset()."""
        sl = np.ascontiguousarray(slots, dtype=np.int32)
        if sl.size:
            check(self._lib.dx_remove_instances(self._h, _p(sl, _i32p), int(sl.size)))

    def run(self, max_iters: int) -> int:
        done = C.c_int32()
        check(self._lib.dx_run(self._h, max_iters, C.byref(done)))
        return done.value

    def step_begin(self) -> int:
        n = C.c_int64()
        check(self._lib.dx_step_begin(self._h, C.byref(n)))
        return n.value

    def get_pending(self, n: int) -> Tuple[np.ndarray, np.ndarray]:
        st = np.empty((n, self.state_bytes), np.uint8)
        gl = np.empty((n, self.state_bytes), np.uint8)
        check(self._lib.dx_get_pending(self._h, _p(st, _u8p), _p(gl, _u8p), n))
        return st, gl

    def step_end(self, vals: np.ndarray) -> None:
        vals = np.ascontiguousarray(vals, dtype=np.float32)
        n = vals.shape[0] if vals.ndim else 0
        check(self._lib.dx_step_end(self._h, _p(vals, _f32p), n))

    def num_instances(self) -> int:
        return self._lib.dx_num_instances(self._h)

    def all_finished(self) -> bool:
        out = C.c_int32()
        check(self._lib.dx_all_finished(self._h, C.byref(out)))
        return bool(out.value)

    def status(self) -> List[dx_inst_status]:
        n = self.num_instances()
        arr = (dx_inst_status * max(n, 1))()
        check(self._lib.dx_get_status(self._h, arr, n))
        return [arr[i] for i in range(n)]

    def counters(self) -> dx_counters:
        c = dx_counters()
        check(self._lib.dx_get_counters(self._h, C.byref(c)))
        return c

    def vi_targets(self, states: np.ndarray, goals: np.ndarray, solved: Optional[np.ndarray] = None) -> np.ndarray:
        """Replay-buffer labels with the loaded network (updaters/updater_rl.py:41-69 / :102-120) -> float64 [n].
        graph_v: min_a (1 + h(child_a)) * (not solved); graph_q: `states` = next states, (1 + min_a q) * (not solved)."""
        states, goals = _u8(states), _u8(goals)
        n = states.shape[0]
        out = np.empty(n, np.float64)
        sol = None if solved is None else np.ascontiguousarray(np.asarray(solved).astype(np.uint8))
        check(self._lib.dx_vi_targets(self._h, _p(states, _u8p), _p(goals, _u8p), _p(sol, _u8p), n,
                                      out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def drain_backups(self, mode: str = "bellman"):
        """graph_v -> (states [n, S] uint8, backups [n] float64, insts [n] int32): Node.bellman_backup() of every node expanded
        since the last drain, in the reference's `popped` order (updaters/updater_rl.py:143-158);
        graph_q -> (states, actions [n] int32, backups, insts): Node.backup_act(action) of every popped edge (:169-189).
        mode (graph_v): "bellman" | "ub_solns" (ub_heur_solns) | "lhbl" (Node.tree_backup) -- the options of UpRLArgs
        (base/updater.py:669-672), computed over the records in the log at drain time.  Empties the log."""
        cap = max(int(self.counters().backups_logged), 1)       # exact-size buffers: the copies land in them directly
        st = np.empty((cap, self.state_bytes), np.uint8)
        bk = np.empty(cap, np.float64)
        ins = np.empty(cap, np.int32)
        act = np.empty(cap, np.int32)
        n = C.c_int64()
        check(self._lib.dx_drain_backups(self._h, DX_BACKUP_MODE[mode], _p(st, _u8p), _p(act, _i32p), bk.ctypes.data_as(C.POINTER(C.c_double)),
                                         _p(ins, _i32p), cap, C.byref(n)))
        k = n.value
        if self.is_q:
            return st[:k], act[:k], bk[:k], ins[:k]
        return st[:k], bk[:k], ins[:k]

    def backup_deepest(self) -> Tuple[np.ndarray, np.ndarray]:
        """UpdateHER._get_her_goals (base/updater.py:485-530): per instance slot, (state of the deepest node with children among
        the logged records -- the root's when nothing deeper was expanded, depth).  Call before drain_backups."""
        n = max(self.num_instances(), 1)
        st = np.empty((n, self.state_bytes), np.uint8)
        dp = np.empty(n, np.int32)
        check(self._lib.dx_backup_deepest(self._h, _p(st, _u8p), _p(dp, _i32p), n))
        k = self.num_instances()
        return st[:k], dp[:k]

    def get_path(self, inst: int, cap: int = 4096, with_states: bool = False):
        """-> (actions list or None, states_on_path [len+1, S] or None)"""
        acts = np.empty(cap, np.int32)
        st = np.empty((cap + 1, self.state_bytes), np.uint8) if with_states else None
        ln = C.c_int32()
        check(self._lib.dx_get_path(self._h, inst, _p(acts, _i32p), _p(st, _u8p), cap, C.byref(ln)))
        if ln.value < 0:
            return None, None
        return acts[: ln.value].tolist(), (st[: ln.value + 1] if with_states else None)

    def curr_nodes(self, inst: int, cap: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        cap = max(cap, 1)
        st = np.empty((cap, self.state_bytes), np.uint8)
        g = np.empty(cap, np.float32)
        h = np.empty(cap, np.float32)
        n = C.c_int32()
        check(self._lib.dx_get_curr_nodes(self._h, inst, _p(st, _u8p), _p(g, _f32p), _p(h, _f32p), cap, C.byref(n)))
        return st[: n.value], g[: n.value], h[: n.value]

    def heur_eval(self, states: np.ndarray, goals: Optional[np.ndarray], out_dim: int) -> np.ndarray:
        states = _u8(states)
        goals = None if goals is None else _u8(goals)
        out = np.empty((states.shape[0], out_dim), np.float32)
        check(self._lib.dx_heur_eval(self._h, _p(states, _u8p), _p(goals, _u8p), states.shape[0], _p(out, _f32p)))
        return out

    # ---- HDA* (hash-distributed single search) ----------------------------------------------------------
    def hda_config(self, rank: int, world: int) -> None:
        check(self._lib.dx_hda_config(self._h, rank, world))

    def hda_record_bytes(self) -> int:
        n = C.c_int32()
        check(self._lib.dx_hda_record_bytes(self._h, C.byref(n)))
        return n.value

    def hda_owner(self, state: np.ndarray) -> int:
        st = _u8(state)
        o = C.c_int32()
        check(self._lib.dx_hda_owner(self._h, _p(st, _u8p), C.byref(o)))
        return o.value

    def hda_add_instance(self, state: np.ndarray, goal: np.ndarray, batch_local: int, weight: float) -> None:
        st, gl = _u8(state), _u8(goal)
        check(self._lib.dx_hda_add_instance(self._h, _p(st, _u8p), _p(gl, _u8p), batch_local, weight))

    def hda_expand(self, send_ptr: int, send_cap_bytes: int, world: int) -> np.ndarray:
        counts = (C.c_int64 * 16)()
        check(self._lib.dx_hda_expand(self._h, C.c_void_p(send_ptr), send_cap_bytes, counts))
        return np.array(counts[:world], dtype=np.int64)

    def hda_absorb(self, recv_ptr: int, n_records: int) -> dx_hda_local:
        out = dx_hda_local()
        check(self._lib.dx_hda_absorb(self._h, C.c_void_p(recv_ptr), n_records, C.byref(out)))
        return out

    def hda_finish(self, k_total: int, f_first_min: float, ub_min: float, children_total: int) -> bool:
        fin = C.c_int32()
        check(self._lib.dx_hda_finish(self._h, k_total, f_first_min, ub_min, children_total, C.byref(fin)))
        return bool(fin.value)

    def set_stream(self, cuda_stream: int) -> None:
        """Run every launch of this engine on the caller's CUDA stream (0 / None: back to the engine's own stream)."""
        check(self._lib.dx_engine_set_stream(self._h, C.c_void_p(cuda_stream or None)))

    def hda_send(self, send_ptr: int, seg_records: int) -> None:
        check(self._lib.dx_hda_send(self._h, C.c_void_p(send_ptr), seg_records))

    def hda_send_peers(self, peer_recv_ptrs, seg_records: int) -> None:
        arr = (C.c_uint64 * len(peer_recv_ptrs))(*[int(x) for x in peer_recv_ptrs])
        check(self._lib.dx_hda_send_peers(self._h, arr, seg_records))

    def hda_receive(self, recv_ptr: int, seg_records: int, local_ptr: int) -> None:
        check(self._lib.dx_hda_receive(self._h, C.c_void_p(recv_ptr), seg_records, C.c_void_p(local_ptr)))

    def hda_reduce(self, all_ptr: int, summary_ptr: int) -> None:
        check(self._lib.dx_hda_reduce(self._h, C.c_void_p(all_ptr), C.c_void_p(summary_ptr)))

    def hda_node(self, local_id: int):
        """-> (state [S], parent global id or None for the root, action, g)"""
        st = np.empty(self.state_bytes, np.uint8)
        par, act, g = C.c_uint32(), C.c_int32(), C.c_float()
        check(self._lib.dx_hda_node(self._h, local_id, _p(st, _u8p), C.byref(par), C.byref(act), C.byref(g)))
        return st, (None if par.value == 0xFFFFFFFF else par.value), act.value, g.value

    def tc_profile(self) -> list:
        arr = (C.c_uint64 * 16)()
        check(self._lib.dx_debug_tc_profile(self._h, arr))
        return [int(x) for x in arr]

    def bench_heur(self, n: int, reps: int) -> float:
        ms = C.c_float()
        check(self._lib.dx_bench_heur(self._h, n, reps, C.byref(ms)))
        return ms.value

    def set_profiling(self, level) -> None:
        """0 / False: off; 1 / True: every stage + GEMM groups; 2: only the hidden-layer GEMM group (lowest overhead)."""
        check(self._lib.dx_set_profiling(self._h, int(level)))

    def stage_ms(self) -> dict:
        arr = (C.c_float * 16)()
        check(self._lib.dx_get_stage_ms(self._h, arr, 16))
        return {k: arr[i] for i, k in enumerate(STAGE_NAMES)}

    def stage_calls(self) -> dict:
        arr = (C.c_int64 * 16)()
        check(self._lib.dx_get_stage_calls(self._h, arr, 16))
        return {k: arr[i] for i, k in enumerate(STAGE_NAMES)}

    def last_run_ms(self) -> float:
        ms = C.c_float()
        check(self._lib.dx_last_run_ms(self._h, C.byref(ms)))
        return ms.value


def pack_state_dict(sd: dict) -> Tuple[np.ndarray, dict]:
    """state_dict (name -> array-like, optional ``module.`` prefix; layout verified against
    tutorial/cube3/models/heur.pt, see deepxube/nnets/resnet_fc.py:41-47) -> (blob, shape info)
    in the order dx_load_weights expects."""
    sd = {k[len("module."):] if k.startswith("module.") else k: np.asarray(v, dtype=np.float32) for k, v in sd.items()
          if not k.endswith("num_batches_tracked")}
    w0 = sd["heur.0.weight"]
    res_dim, in_dim = w0.shape
    nb = 0
    while f"heur.1.blocks.{nb}.0.layers.0.0.weight" in sd:
        nb += 1
    bn = any(k.endswith("running_mean") for k in sd)
    parts = [w0.ravel(), sd["heur.0.bias"].ravel()]
    for b in range(nb):
        for j in range(2):
            p = f"heur.1.blocks.{b}.0.layers.{j}"
            parts += [sd[p + ".0.weight"].ravel(), sd[p + ".0.bias"].ravel()]
            if bn:
                parts += [sd[p + ".1.weight"].ravel(), sd[p + ".1.bias"].ravel(), sd[p + ".1.running_mean"].ravel(),
                          sd[p + ".1.running_var"].ravel()]
    wo = sd["heur.2.weight"]
    parts += [wo.ravel(), sd["heur.2.bias"].ravel()]
    return np.concatenate(parts).astype(np.float32), dict(res_dim=res_dim, in_dim=in_dim, num_blocks=nb, batch_norm=bn,
                                                          out_dim=wo.shape[0])
