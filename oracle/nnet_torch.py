"""Oracle restatement of the ``resnet_fc.<H>H_<B>B[_bn]`` cost-to-go network (torch CPU, fp32).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows, op for op:
  OneHot.forward                 deepxube/pytorch/pytorch_models.py:15-24   (one_hot(x.long(), depth).float().reshape(N, d*c))
  ResnetFCHeur._forward          deepxube/nnets/resnet_fc.py:49-52          (Linear(in,H) with NO activation / BN after it)
  FullyConnectedModel.forward    deepxube/pytorch/pytorch_models.py:198-209 ([Linear, BN, ReLU], [Linear, BN, Linear-act])
  ResnetModel.forward            deepxube/pytorch/pytorch_models.py:149-160 (x = relu(block(x) + x))
  final Linear(H, out_dim)       deepxube/nnets/resnet_fc.py:46
  HeurVNNetPar.process_outputs   deepxube/base/pathfind_fns.py:218-221      (h = max(out[:,0], 0) -> float64)
  HeurQNNetParFixOut.process_outputs  deepxube/pathfind_fns/fns_core.py:78-84  (q = max(out, 0) -> float64)
The q_fix gather (base/nnet.py:129-135) is the identity for heurq_fixout because action_idxs is
[0..A-1] for every row (nnet_input.py:270-276), so it is omitted.

state_dict key layout (verified on tutorial/cube3/models/heur.pt, optional ``module.`` prefix):
  heur.0.{weight,bias}
  heur.1.blocks.<i>.0.layers.<j>.0.{weight,bias}
  heur.1.blocks.<i>.0.layers.<j>.1.{weight,bias,running_mean,running_var,num_batches_tracked}
  heur.2.{weight,bias}
"""
import re
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # nn.BatchNorm1d default, used unchanged by the reference


def strip_module_prefix(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """(ref: deepxube/pytorch/nnet_utils.py:52-56)"""
    return {re.sub(r"^module\.", "", k): v for k, v in sd.items()}


def num_blocks_of(sd: Dict[str, torch.Tensor]) -> int:
    idx = [int(m.group(1)) for k in sd for m in [re.match(r"heur\.1\.blocks\.(\d+)\.", k)] if m]
    return (max(idx) + 1) if idx else 0


def has_bn(sd: Dict[str, torch.Tensor]) -> bool:
    return any(k.endswith("running_mean") for k in sd)


def random_state_dict(in_dim: int, res_dim: int, num_blocks: int, out_dim: int, batch_norm: bool = True,
                      seed: int = 0, randomize_bn: bool = False) -> Dict[str, torch.Tensor]:
    """Random-init weights with nn.Linear's default distribution (kaiming_uniform(a=sqrt(5)) ==
    U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias), eval-mode BN mu=0, var=1, gamma=1, beta=0
    unless ``randomize_bn`` (used by tests to exercise BN folding)."""
    g = torch.Generator().manual_seed(seed)

    def lin(o: int, i: int, prefix: str, sd: dict) -> None:
        bound = 1.0 / np.sqrt(i)
        sd[prefix + ".weight"] = (torch.rand(o, i, generator=g) * 2 - 1) * bound
        sd[prefix + ".bias"] = (torch.rand(o, generator=g) * 2 - 1) * bound

    sd: Dict[str, torch.Tensor] = {}
    lin(res_dim, in_dim, "heur.0", sd)
    for b in range(num_blocks):
        for j in range(2):
            p = f"heur.1.blocks.{b}.0.layers.{j}"
            lin(res_dim, res_dim, p + ".0", sd)
            if batch_norm:
                if randomize_bn:
                    sd[p + ".1.weight"] = torch.rand(res_dim, generator=g) + 0.5
                    sd[p + ".1.bias"] = torch.rand(res_dim, generator=g) - 0.5
                    sd[p + ".1.running_mean"] = (torch.rand(res_dim, generator=g) - 0.5) * 0.2
                    sd[p + ".1.running_var"] = torch.rand(res_dim, generator=g) + 0.5
                else:
                    sd[p + ".1.weight"] = torch.ones(res_dim)
                    sd[p + ".1.bias"] = torch.zeros(res_dim)
                    sd[p + ".1.running_mean"] = torch.zeros(res_dim)
                    sd[p + ".1.running_var"] = torch.ones(res_dim)
                sd[p + ".1.num_batches_tracked"] = torch.tensor(0)
    lin(out_dim, res_dim, "heur.2", sd)
    return sd


@torch.no_grad()
def resnet_fc_forward(sd: Dict[str, torch.Tensor], x_int: np.ndarray, one_hot_depth: int,
                      batch_size: Optional[int] = None, act_depth: int = 0) -> np.ndarray:
    """x_int [N, D] integer inputs -> raw network output [N, out_dim] float32.
    Chunked like nnet_batched (ref: deepxube/pytorch/nnet_utils.py:73-111).
    act_depth > 0: the LAST column of x_int is an action index with its own one-hot depth, encoded separately and
    concatenated (ref: deepxube/nnets/resnet_fc.py:23-30, :49-51; action-as-input Q networks)."""
    sd = strip_module_prefix(sd)
    nb, bn = num_blocks_of(sd), has_bn(sd)
    n = x_int.shape[0]
    outs = []
    dt = sd["heur.0.weight"].dtype           # float32 like the reference; a float64 state_dict gives a float64 yardstick
    step = n if not batch_size else batch_size
    for s in range(0, max(n, 1), max(step, 1)):
        xi = torch.as_tensor(np.ascontiguousarray(x_int[s:s + step]))
        xa = None
        if act_depth > 0:
            xa = F.one_hot(xi[:, -1].long(), act_depth).to(dt)
            xi = xi[:, :-1]
        if one_hot_depth > 1:
            x = F.one_hot(xi.long(), one_hot_depth).to(dt)
            x = x.reshape(x.shape[0], -1)
        else:
            x = xi.to(dt)
        if xa is not None:
            x = torch.cat([x, xa], dim=1)
        x = F.linear(x, sd["heur.0.weight"], sd["heur.0.bias"])
        for b in range(nb):
            res = x
            for j in range(2):
                p = f"heur.1.blocks.{b}.0.layers.{j}"
                x = F.linear(x, sd[p + ".0.weight"], sd[p + ".0.bias"])
                if bn:
                    x = F.batch_norm(x, sd[p + ".1.running_mean"], sd[p + ".1.running_var"], sd[p + ".1.weight"],
                                     sd[p + ".1.bias"], training=False, eps=BN_EPS)
                if j == 0:
                    x = F.relu(x)
            x = F.relu(x + res)
        x = F.linear(x, sd["heur.2.weight"], sd["heur.2.bias"])
        outs.append(x.numpy())
    return np.concatenate(outs, axis=0) if outs else np.zeros((0, sd["heur.2.weight"].shape[0]), np.float32)


def heurv_from_state_dict(sd: Dict[str, torch.Tensor], one_hot_depth: int, batch_size: Optional[int] = None):
    """Returns fn(x_int [N, D]) -> float64 [N] (clipped at 0)."""
    def fn(x_int: np.ndarray) -> np.ndarray:
        out = resnet_fc_forward(sd, x_int, one_hot_depth, batch_size)
        return np.maximum(out[:, 0], 0).astype(np.float64)
    return fn


def heurq_from_state_dict(sd: Dict[str, torch.Tensor], one_hot_depth: int, batch_size: Optional[int] = None):
    """Returns fn(x_int [N, D]) -> float64 [N, A] (clipped at 0 per element)."""
    def fn(x_int: np.ndarray) -> np.ndarray:
        out = resnet_fc_forward(sd, x_int, one_hot_depth, batch_size)
        return np.maximum(out, 0).astype(np.float64)
    return fn


def heurq_in_from_state_dict(sd: Dict[str, torch.Tensor], one_hot_depth: int, num_actions: int, batch_size: Optional[int] = None):
    """Action-as-input Q function (ref: deepxube/pathfind_fns/fns_core.py:99-133): every state is repeated once per action,
    the action index is the extra input; returns fn(x_int [N, D]) -> float64 [N, A] (clipped at 0)."""
    def fn(x_int: np.ndarray) -> np.ndarray:
        n = x_int.shape[0]
        rep = np.repeat(np.asarray(x_int, np.int64), num_actions, axis=0)
        acts = np.tile(np.arange(num_actions, dtype=np.int64), n)[:, None]
        out = resnet_fc_forward(sd, np.concatenate([rep, acts], axis=1), one_hot_depth, batch_size, act_depth=num_actions)
        return np.maximum(out[:, 0], 0).astype(np.float64).reshape(n, num_actions)
    return fn
