"""GPU parity tests of the one-launch small-evaluation network kernel (csrc/k_mlp_small.cu): engines whose evaluations never
exceed DX_SMALL_NET_ROWS rows run the whole resnet_fc in ONE launch (solve-time regime, deepxube/_solve.py:150-161).
Compared with the torch-CPU restatement of the reference network (oracle/nnet_torch.py: resnet_fc.py:19-52,
pytorch_models.py:15-24, 149-209) in float64; tolerances: fp32-accurate mode 5e-5 absolute, bf16 mode 0.05 absolute on
outputs spread over 0..20.  Shapes: every domain's input encoding (one-hot, goal inputs, depth-1 raw values), V and Q
outputs, widths that do not fill the 16-column warp tiles, zero residual blocks, tiles with ragged row counts."""
import numpy as np
import pytest

from oracle.domains_np import OracleDomain
from oracle.nnet_torch import random_state_dict, resnet_fc_forward

pytestmark = pytest.mark.gpu


def _lib():
    from deepxube_b200 import _lib
    return _lib


def _net(in_dim, h, nb, od, seed):
    sd = {k: v.numpy().copy() for k, v in random_state_dict(in_dim, h, nb, od, True, seed=seed, randomize_bn=True).items()}
    sd["heur.2.weight"] *= 12.0
    sd["heur.2.bias"] = sd["heur.2.bias"] * 12.0 + 4.0
    return sd


def _states(dom, n, seed):
    rng = np.random.default_rng(seed)
    if dom.name == "grid":                      # (state = robot position, goal = goal position: both random cells)
        return (rng.integers(0, dom.dim, size=(n, dom.state_len)).astype(np.uint8),
                rng.integers(0, dom.dim, size=(n, dom.state_len)).astype(np.uint8))
    goals = np.tile(dom.canonical_goal(), (n, 1))
    st = goals.copy()
    for _ in range(30):
        st = dom.next_state(st, rng.integers(0, dom.num_actions, size=n))
    return st, goals


def _ref(sd, dom, st, goals, depth):
    import torch
    sd64 = {k: torch.from_numpy(np.asarray(v)).double() if np.asarray(v).dtype.kind == "f" else torch.from_numpy(np.asarray(v))
            for k, v in sd.items()}
    x = dom.nnet_input(st, goals).astype(np.int64)
    return np.maximum(np.asarray(resnet_fc_forward(sd64, x, depth)), 0.0)


CASES = [  # domain, dim, H, blocks, Q?, rows
    ("cube3", 0, 200, 2, False, 1203),      # the tutorial KAT shape (graph.100B: ~1200 children), ragged last tile
    ("cube3", 0, 200, 2, True, 333),        # 12 outputs, h = min_a q
    ("cube3", 0, 100, 1, False, 17),        # 100 = 6 full warps + 4 columns
    ("cube3", 0, 256, 1, False, 64),        # every consumer warp active
    ("cube3", 0, 72, 0, False, 50),         # no residual block: first layer -> output layer
    ("npuzzle", 4, 128, 2, False, 700),     # 16 x 16 one-hot
    ("npuzzle", 5, 200, 2, False, 100),     # 625 one-hot inputs: 40 k-steps in the first layer
    ("grid", 7, 100, 2, False, 300),        # goal position is a network input (second staged row)
    ("lightsout", 5, 64, 1, True, 40),      # depth-1 inputs (raw 0 / 1 values), 25 outputs
]


@pytest.mark.parametrize("mode_name", ["fp32", "bf16"])
@pytest.mark.parametrize("dname,dim,H,nb,is_q,rows", CASES)
def test_small_network_kernel_matches_the_reference_network(mode_name, dname, dim, H, nb, is_q, rows):
    L = _lib()
    dom = OracleDomain(dname, dim)
    in_count, depth = dom.input_info()
    od = dom.num_actions if is_q else 1
    sd = _net(in_count * max(depth, 1), H, nb, od, seed=H + nb)
    blob, info = L.pack_state_dict(sd)
    mode = L.DX_HEUR_NET_FP32 if mode_name == "fp32" else L.DX_HEUR_NET_BF16
    # capacity 400 popped nodes x actions <= 8192 rows: the engine takes the one-launch kernel
    pops = max(8, min(400, 8192 // dom.num_actions - 1))
    eng = L.Engine(dname, dim, "graph_q" if is_q else "graph_v", mode, max_instances=2, max_pop_total=pops, max_nodes=50000)
    eng.load_weights(in_count, depth, H, nb, od, True, blob)
    st, goals = _states(dom, rows, seed=3)
    l0 = eng.counters().kernel_launches
    got = eng.heur_eval(st, goals, od)
    launches = eng.counters().kernel_launches - l0
    ref = _ref(sd, dom, st, goals, depth)
    tol = 5e-5 if mode_name == "fp32" else 0.05
    err = float(np.abs(got.astype(np.float64) - ref).max())
    assert err <= tol, (mode_name, dname, H, nb, err)
    assert ref.max() > 1.0                                      # the outputs are not all clipped to zero
    # one launch per chunk of the engine's capacity (plus nothing else): the fused kernel really ran
    cap = pops * (1 if is_q else dom.num_actions)
    assert launches <= -(-rows // cap) * 2 + 1, launches       # (job set + ONE network launch per chunk, job restore)
    # the same rows in another order / batch composition give bit-identical values (row results do not depend on the tile)
    perm = np.random.default_rng(0).permutation(rows)
    again = eng.heur_eval(st[perm], goals[perm], od)
    assert np.array_equal(again, got[perm])
    eng.close()


def test_small_and_large_engines_agree_within_tolerance():
    """The same weights on an engine sized for big batches (per-layer tcgen05 path) and on a small engine (one-launch kernel):
    fp32-accurate modes agree to 1e-4, and each engine is self-consistent (the choice of kernel is per engine, never per call)."""
    L = _lib()
    dom = OracleDomain("cube3")
    sd = _net(324, 200, 2, 1, seed=11)
    blob, _ = L.pack_state_dict(sd)
    st, goals = _states(dom, 500, seed=5)
    outs = []
    for pops in (100, 4000):
        eng = L.Engine("cube3", 0, "graph_v", L.DX_HEUR_NET_FP32, max_instances=1, max_pop_total=pops, max_nodes=200000)
        eng.load_weights(54, 6, 200, 2, 1, True, blob)
        a = eng.heur_eval(st, goals, 1)
        b = eng.heur_eval(st[:7], goals[:7], 1)                 # a tiny call on the same engine: same kernel, same bits
        assert np.array_equal(a[:7], b)
        outs.append(a)
        eng.close()
    assert float(np.abs(outs[0] - outs[1]).max()) <= 1e-4
