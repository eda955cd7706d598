"""HDA*-style hash-distributed batch weighted A*: ONE search instance over W GPUs (BASELINE config 5).

Not in the reference (its search is single-process Python); semantics follow graph_search.py with the
distribution rules of SURVEY.md 8e: owner(state) = hash(state) mod W owns the state's CLOSED entry, node and
OPEN entry; per iteration every rank pops its local top-(B/W), expands, and routes each child record to its
owner with one all-to-all; lb / ub / finished() are evaluated on globally reduced quantities so every rank
takes the same decision in the same iteration.  Local top-(B/W) per shard is not the global top-B, so costs /
node counts are compared with the single-GPU engine statistically, not bit-for-bit (W = 1 IS bit-identical).

Two transports with identical engine calls:
  * `dist` given       one process per GPU, `torch.distributed` (NCCL): `all_to_all_single` moves the records,
                       `all_gather` the per-rank scalars
  * `dist` is None     W ranks emulated in one process on one GPU (the engines run one after the other, so
                       nothing ever waits on a concurrently running kernel); used by the single-GPU tests

Two drivers of the iteration:
  * device-resident (default): the engine kernels and the collectives are ordered on ONE CUDA stream and nothing inside
    an iteration synchronises with the host -- fixed-size per-destination segments carry their record count in a
    header, so the exchange needs no counts on the host; with `exchange="p2p"` (default) the pack kernel itself stores
    the segments into the owners' receive buffers over NVLink peer memory (symmetric memory) and one device-side
    barrier follows, with `exchange="nccl"` an equal-split all-to-all moves them; per-rank scalars are all-gathered as
    a device tensor and reduced by a kernel on every rank; the host reads one 64-byte summary per iteration
  * host-driven (`device_resident=False`): variable-size all-to-all with the counts exchanged through the host
    (the first implementation, kept for A/B measurements; ~8 device-to-host reads per iteration)
Both produce the same received-children order (source-major, sender's order), hence identical searches.
"""
from typing import Any, List, Optional

import numpy as np
import torch

from deepxube_b200 import _lib
from deepxube_b200.base.pathfind_fns import load_nnet_into_engine


class HdaSearch:
    def __init__(self, domain_name: str, dim: int, world: int, batch_size: int, weight: float, nnet: Any = None,
                 precision: str = "fp32", max_nodes_per_rank: int = 1 << 22, dist: Optional[Any] = None,
                 device: Optional[int] = None, capacity_slack: float = 2.0, device_resident: bool = True,
                 exchange: str = "p2p"):
        self.dist = dist
        self.device_resident = device_resident
        self.world = world
        self.weight = weight
        self.batch_local = max(1, -(-batch_size // world))
        if dist is not None:
            assert dist.get_world_size() == world
            self.local_ranks = [dist.get_rank()]
        else:
            self.local_ranks = list(range(world))
        dev = _lib.current_device() if device is None else device
        self.tdev = torch.device("cuda", dev)
        mode = _lib.DX_HEUR_ZERO if nnet is None else (_lib.DX_HEUR_NET_BF16 if precision == "bf16" else _lib.DX_HEUR_NET_FP32)
        self.engines: List[_lib.Engine] = []
        for r in self.local_ranks:
            eng = _lib.Engine(domain_name, dim, "graph_v", mode, max_instances=1,
                              max_pop_total=int(np.ceil(self.batch_local * capacity_slack)) + 64,
                              max_nodes=min(max_nodes_per_rank, (1 << 28) - 1), device=dev)
            if nnet is not None:
                load_nnet_into_engine(eng, nnet)
            eng.hda_config(r, world)
            self.engines.append(eng)
        e0 = self.engines[0]
        self.rec = e0.hda_record_bytes()
        self.num_actions = e0.num_actions
        self.cap_records = (int(np.ceil(self.batch_local * capacity_slack)) + 64) * self.num_actions
        self.send = [torch.empty(self.cap_records * self.rec, dtype=torch.uint8, device=self.tdev) for _ in self.local_ranks]
        self.recv = [torch.empty(self.cap_records * self.rec, dtype=torch.uint8, device=self.tdev) for _ in self.local_ranks]
        self.itr = 0
        self.num_nodes_generated = 0
        self.goal = None          # (ub, rank, local id)
        self.finished = False
        # per-iteration phase timing (CUDA events on the iteration's stream) and load statistics, for bench.py's `hda` block
        self.timing = False
        self.phase_ms = {"pack": 0.0, "exchange": 0.0, "receive": 0.0, "all_gather": 0.0, "reduce": 0.0}
        self.recv_max_sum = 0.0       # sum over iterations of max over ranks of the children received
        self.recv_mean_sum = 0.0      # ... of the mean over ranks (= children generated / W)
        self.timed_iters = 0
        self._ev = None
        if device_resident:
            # expected records per destination = batch_local * actions / world; the slack covers the hash imbalance
            self.seg_records = int(np.ceil(self.batch_local * self.num_actions / world * max(capacity_slack, 1.0) * 1.25)) + 64
            self.seg_bytes = 16 + self.seg_records * self.rec
            self.stream = torch.cuda.Stream(device=self.tdev)
            with torch.cuda.stream(self.stream):
                self.send = [torch.zeros(world * self.seg_bytes, dtype=torch.uint8, device=self.tdev) for _ in self.local_ranks]
                self.recv = [torch.zeros(world * self.seg_bytes, dtype=torch.uint8, device=self.tdev) for _ in self.local_ranks]
                nl = _lib.DX_HDA_LOCAL_DOUBLES
                self.loc = [torch.zeros(nl, dtype=torch.float64, device=self.tdev) for _ in self.local_ranks]
                self.loc_all = torch.zeros(world * nl, dtype=torch.float64, device=self.tdev)
                self.summary = [torch.zeros(8, dtype=torch.float64, device=self.tdev) for _ in self.local_ranks]
            self.stream.synchronize()
            for eng in self.engines:
                eng.set_stream(self.stream.cuda_stream)
            # Record exchange.  "p2p": the pack kernel stores every segment straight into the owner's receive buffer over
            # peer memory (torch symmetric memory maps the W receive buffers into every process), followed by one
            # symmetric-memory barrier -- no collective moves the records.  "nccl": pack locally, all_to_all_single.
            self.exchange = exchange
            self.symm = None
            if exchange == "p2p":
                if dist is None:
                    self.peer_ptrs = [[self.recv[d].data_ptr() for d in range(world)] for _ in self.local_ranks]
                else:
                    try:
                        import torch.distributed._symmetric_memory as symm_mem
                        buf = symm_mem.empty(world * self.seg_bytes, dtype=torch.uint8, device=self.tdev)
                        self.symm = symm_mem.rendezvous(buf, dist.group.WORLD)
                        buf.zero_()
                        self.recv = [buf]
                        self.peer_ptrs = [[int(p) for p in self.symm.buffer_ptrs]]
                        torch.cuda.synchronize(self.tdev)
                        dist.barrier()
                    except Exception as ex:  # noqa: BLE001  (no peer access on this machine: keep the NCCL exchange)
                        print(f"HDA*: peer-memory exchange unavailable ({ex}); using the NCCL all-to-all", flush=True)
                        self.exchange, self.symm = "nccl", None

    # ---- transports --------------------------------------------------------------------------------------
    def _exchange(self, counts: List[np.ndarray]) -> List[int]:
        """counts[i][dst] = records local rank i sends to dst.  Fills self.recv, returns records received."""
        rec = self.rec
        if self.dist is None:
            n_recv = []
            for di, dst in enumerate(self.local_ranks):
                off = 0
                for si, _ in enumerate(self.local_ranks):
                    c = int(counts[si][dst])
                    s0 = int(counts[si][:dst].sum())
                    if c:
                        self.recv[di][off * rec:(off + c) * rec].copy_(self.send[si][s0 * rec:(s0 + c) * rec])
                    off += c
                if off > self.cap_records:
                    raise _lib.DxError("HDA*: receive buffer too small (raise capacity_slack)")
                n_recv.append(off)
            torch.cuda.synchronize(self.tdev)
            return n_recv
        dist = self.dist
        mine = torch.tensor(counts[0], dtype=torch.int64, device=self.tdev)
        allc = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(allc, mine)
        allc = torch.stack(allc).cpu().numpy()                     # allc[src][dst]
        r = self.local_ranks[0]
        in_split = [int(c) * rec for c in counts[0]]
        out_split = [int(allc[src][r]) * rec for src in range(self.world)]
        n = sum(out_split) // rec
        if n > self.cap_records:
            raise _lib.DxError("HDA*: receive buffer too small (raise capacity_slack)")
        dist.all_to_all_single(self.recv[0][: n * rec], self.send[0][: int(counts[0].sum()) * rec], out_split, in_split)
        torch.cuda.synchronize(self.tdev)
        return [n]

    def _gather_locals(self, locs: List[Any], failed: bool = False) -> List[Any]:
        """Per-rank scalars of every rank, in rank order: rows of (k, f_first, ub, goal_local, children_sent, error flag)."""
        rows = [[float(l.k_popped), l.f_first, l.ub, float(l.goal_local), float(l.children_sent), 1.0 if failed else 0.0] for l in locs]
        if self.dist is None:
            return rows
        t = torch.tensor(rows[0], dtype=torch.float64, device=self.tdev)
        out = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [o.cpu().tolist() for o in out]

    # ---- search ---------------------------------------------------------------------------------------------
    def start(self, state: np.ndarray, goal: np.ndarray) -> None:
        for eng in self.engines:
            eng.reset()
            eng.hda_add_instance(state, goal, self.batch_local, self.weight)
        self.itr, self.num_nodes_generated, self.goal, self.finished = 0, 0, None, False

    def _step_device(self) -> bool:
        """One iteration with every launch and collective on self.stream; the only host read is the summary."""
        W, seg = self.world, self.seg_bytes
        tm = self.timing
        if tm and self._ev is None:
            self._ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        ev = self._ev
        gen_before = self.num_nodes_generated
        with torch.cuda.stream(self.stream):
            if tm:
                ev[0].record(self.stream)
            if self.exchange == "p2p":
                for i, eng in enumerate(self.engines):
                    eng.hda_send_peers(self.peer_ptrs[i], self.seg_records)               # pack + store into the owners' buffers
                if tm:
                    ev[1].record(self.stream)
                if self.symm is not None:
                    self.symm.barrier(channel=0)                                          # every rank's segments have landed
            else:
                for i, eng in enumerate(self.engines):
                    eng.hda_send(self.send[i].data_ptr(), self.seg_records)
                if tm:
                    ev[1].record(self.stream)
                if self.dist is not None:
                    self.dist.all_to_all_single(self.recv[0], self.send[0])             # equal splits of seg_bytes
                else:
                    for di in range(W):
                        for si in range(W):
                            self.recv[di][si * seg:(si + 1) * seg].copy_(self.send[si][di * seg:(di + 1) * seg], non_blocking=True)
            if tm:
                ev[2].record(self.stream)
            for i, eng in enumerate(self.engines):
                eng.hda_receive(self.recv[i].data_ptr(), self.seg_records, self.loc[i].data_ptr())
            if tm:
                ev[3].record(self.stream)
            if self.dist is not None:
                self.dist.all_gather_into_tensor(self.loc_all, self.loc[0])
            else:
                torch.cat(self.loc, out=self.loc_all)
            if tm:
                ev[4].record(self.stream)
            for i, eng in enumerate(self.engines):
                eng.hda_reduce(self.loc_all.data_ptr(), self.summary[i].data_ptr())
            if tm:
                ev[5].record(self.stream)
            sm = self.summary[0].cpu().tolist()                                          # the iteration's one synchronisation
        if sm[6] != 0:
            # the flags are OR-ed over ALL ranks by dx_hda_reduce, so every rank raises here in the same iteration
            # (no rank is left waiting in a barrier / collective for a peer that has gone)
            raise _lib.DxError(f"HDA*: device capacity exceeded on some rank (flags {int(sm[6])}): raise capacity_slack / max_nodes_per_rank")
        if tm:
            for k, name in enumerate(("pack", "exchange", "receive", "all_gather", "reduce")):
                self.phase_ms[name] += ev[k].elapsed_time(ev[k + 1])
            self.recv_max_sum += sm[7]
            self.recv_mean_sum += (int(sm[2]) - gen_before) / W
            self.timed_iters += 1
        self.finished = sm[0] != 0
        self.itr = int(sm[1])
        self.num_nodes_generated = int(sm[2])
        self.goal = (sm[3], int(sm[4]), int(sm[5])) if sm[4] >= 0 else None
        return self.finished

    def step(self) -> bool:
        if self.device_resident:
            return self._step_device()
        # host-driven arm: a DxError on one rank must not leave the others waiting in a collective -- the failing rank keeps
        # taking part (empty sends) and the flag travels with the gathered scalars, so every rank raises together
        err: Optional[Exception] = None
        counts = []
        for i, eng in enumerate(self.engines):
            try:
                counts.append(eng.hda_expand(self.send[i].data_ptr(), self.send[i].numel(), self.world))
            except _lib.DxError as ex:
                err = err or ex
                counts.append(np.zeros(self.world, np.int64))
        try:
            n_recv = self._exchange(counts)
        except _lib.DxError as ex:                     # receive buffer too small: detected after the collective completed
            err = err or ex
            n_recv = [0 for _ in self.engines]
        locs = []
        for i, eng in enumerate(self.engines):
            try:
                if err is not None:
                    raise err
                locs.append(eng.hda_absorb(self.recv[i].data_ptr(), n_recv[i]))
            except _lib.DxError as ex:
                err = err or ex
                locs.append(_lib.dx_hda_local(0, float("inf"), float("inf"), -1, 0, 0, 0))
        rows = self._gather_locals(locs, failed=err is not None)
        if any(r[5] != 0 for r in rows):
            raise _lib.DxError(f"HDA*: a rank failed ({err if err is not None else 'reported by a peer'})")
        k_total = int(sum(r[0] for r in rows))
        f_first_min = min(r[1] for r in rows)
        children = int(sum(r[4] for r in rows))
        best = min(((r[2], rank, int(r[3])) for rank, r in enumerate(rows) if r[3] >= 0), default=None)
        ub_min = best[0] if best is not None else float("inf")
        self.goal = best
        fins = [eng.hda_finish(k_total, f_first_min, ub_min, children) for eng in self.engines]
        self.itr += 1
        self.num_nodes_generated += children
        self.finished = fins[0]
        return self.finished

    def solve(self, state: np.ndarray, goal: np.ndarray, max_itrs: Optional[int] = None) -> dict:
        self.start(state, goal)
        while not self.finished and (max_itrs is None or self.itr < max_itrs):
            self.step()
        out = dict(iterations=self.itr, num_nodes_generated=self.num_nodes_generated, solved=self.goal is not None,
                   path_cost=float("inf"), actions=None)
        if self.goal is not None:
            out["path_cost"] = float(self.goal[0])
            out["actions"] = self.get_path()
        return out

    def _node(self, rank: int, local_id: int):
        """(parent global id or None, action) of node `local_id` on `rank`, known to every rank."""
        if self.dist is None:
            _, par, act, _ = self.engines[self.local_ranks.index(rank)].hda_node(local_id)
            return par, act
        obj = [None]
        if self.local_ranks[0] == rank:
            _, par, act, _ = self.engines[0].hda_node(local_id)
            obj = [(par, act)]
        self.dist.broadcast_object_list(obj, src=rank)
        return obj[0]

    def get_path(self) -> List[int]:
        """Parent walk across shards (deepxube/base/pathfinding.py:93-120 with global node ids)."""
        assert self.goal is not None
        _, rank, nid = self.goal
        acts: List[int] = []
        while True:
            par, act = self._node(rank, nid)
            if par is None:
                break
            acts.append(act)
            rank, nid = par >> 28, par & ((1 << 28) - 1)
        acts.reverse()
        return acts

    def close(self) -> None:
        for eng in self.engines:
            eng.close()
        self.engines = []
