"""`resnet_fc.<H>H_<B>B[_bn]` cost-to-go network: shape description + weights (host side).

Registered as "resnet_fc" with the reference's constructor signature and parser tokens
(deepxube/nnets/resnet_fc.py:13-52, 117-133).  The arithmetic runs on the GPU (csrc/k_mlp_f32.cu,
csrc/k_mlp_tc.cu); this class only owns the state_dict, in the reference's key layout
  heur.0.{weight,bias}; heur.1.blocks.<i>.0.layers.<j>.0.{weight,bias};
  heur.1.blocks.<i>.0.layers.<j>.1.{weight,bias,running_mean,running_var,num_batches_tracked}; heur.2.{weight,bias}
so that `.pt` files written by the reference's trainer load unchanged (optional `module.` prefix stripped,
deepxube/pytorch/nnet_utils.py:44-62).  weight_norm / layer_norm / non-ReLU activations are outside the
accelerated path and raise.
"""
import re
from typing import Any, Dict, Optional

import numpy as np

from deepxube_b200.base.factory import DelimParser
from deepxube_b200.factories.nnet_factory import deepxube_nnet_factory


@deepxube_nnet_factory.register_class("resnet_fc")
class ResnetFCHeur:
    def __init__(self, nnet_input: Any, out_dim: int, q_fix: bool, res_dim: int = 1000, num_blocks: int = 4,
                 batch_norm: bool = False, weight_norm: bool = False, layer_norm: bool = False, act_fn: str = "RELU"):
        if weight_norm or layer_norm or act_fn.upper() != "RELU":
            raise NotImplementedError("deepxube_b200 accelerates resnet_fc with ReLU and optional batch norm only "
                                      "(weight_norm / layer_norm / other activations are not on the B200 path)")
        self.nnet_input = nnet_input
        self.out_dim, self.q_fix = int(out_dim), bool(q_fix)
        self.res_dim, self.num_blocks, self.batch_norm = int(res_dim), int(num_blocks), bool(batch_norm)
        dims, depths = nnet_input.get_input_info()
        # one flat one-hot block, optionally followed by the action index of an action-as-input Q network
        # (deepxube/nnets/resnet_fc.py:23-30: every input is one-hot encoded with its own depth and concatenated)
        assert len(dims) == len(depths) and len(dims) in (1, 2), "flat one-hot input (+ one action input) expected"
        self.in_count, self.one_hot_depth = int(dims[0]), int(depths[0])
        self.act_depth = 0
        if len(dims) == 2:
            assert int(dims[1]) == 1 and not q_fix, "second input must be a single action index"
            self.act_depth = int(depths[1])
        self.in_dim = self.in_count * self.one_hot_depth + self.act_depth
        self._sd: Optional[Dict[str, np.ndarray]] = None
        self.version = 0          # bumped whenever the parameters change: engines holding a copy re-upload (dx_load_weights)

    # ---- weights -------------------------------------------------------------------------------------
    def expected_shapes(self) -> Dict[str, tuple]:
        h, shapes = self.res_dim, {}
        shapes["heur.0.weight"], shapes["heur.0.bias"] = (h, self.in_dim), (h,)
        for b in range(self.num_blocks):
            for j in range(2):
                p = f"heur.1.blocks.{b}.0.layers.{j}"
                shapes[p + ".0.weight"], shapes[p + ".0.bias"] = (h, h), (h,)
                if self.batch_norm:
                    for nm in ("weight", "bias", "running_mean", "running_var"):
                        shapes[f"{p}.1.{nm}"] = (h,)
        shapes["heur.2.weight"], shapes["heur.2.bias"] = (self.out_dim, h), (self.out_dim,)
        return shapes

    def load_state_dict(self, sd: Dict[str, Any]) -> None:
        clean = {}
        for k, v in sd.items():
            k = re.sub(r"^module\.", "", k)
            if k.endswith("num_batches_tracked"):
                continue
            clean[k] = np.asarray(v.detach().cpu().numpy() if hasattr(v, "detach") else v, dtype=np.float32)
        want = self.expected_shapes()
        missing, extra = sorted(set(want) - set(clean)), sorted(set(clean) - set(want))
        if missing or extra:
            raise RuntimeError(f"Error(s) in loading state_dict for ResnetFCHeur: missing keys {missing}, unexpected keys {extra}")
        for k, shp in want.items():
            if tuple(clean[k].shape) != shp:
                raise RuntimeError(f"size mismatch for {k}: got {tuple(clean[k].shape)}, expected {shp}")
        self._sd = clean
        self.version += 1

    def init_random(self, seed: Optional[int] = None) -> None:
        """nn.Linear default init in the reference's module construction order (so `torch.manual_seed(s)`
        gives the weights the reference's freshly constructed ResnetFCHeur would have); BN gamma 1, beta 0,
        running stats (0, 1)."""
        import torch
        from torch import nn
        if seed is not None:
            torch.manual_seed(seed)
        sd: Dict[str, np.ndarray] = {}

        def take(prefix: str, lin: nn.Linear) -> None:
            sd[prefix + ".weight"] = lin.weight.detach().numpy().copy()
            sd[prefix + ".bias"] = lin.bias.detach().numpy().copy()

        h = self.res_dim
        take("heur.0", nn.Linear(self.in_dim, h))
        for b in range(self.num_blocks):
            for j in range(2):
                p = f"heur.1.blocks.{b}.0.layers.{j}"
                take(p + ".0", nn.Linear(h, h))
                if self.batch_norm:
                    sd[p + ".1.weight"], sd[p + ".1.bias"] = np.ones(h, np.float32), np.zeros(h, np.float32)
                    sd[p + ".1.running_mean"], sd[p + ".1.running_var"] = np.zeros(h, np.float32), np.ones(h, np.float32)
        take("heur.2", nn.Linear(h, self.out_dim))
        self._sd = sd
        self.version += 1

    def state_dict(self) -> Dict[str, np.ndarray]:
        if self._sd is None:
            self.init_random()
        assert self._sd is not None
        return self._sd

    def save(self, file: str) -> None:
        import torch
        sd = {}
        for k, v in self.state_dict().items():
            sd[k] = torch.from_numpy(np.array(v))
            if k.endswith("running_var"):
                sd[k.replace("running_var", "num_batches_tracked")] = torch.tensor(0)
        torch.save(sd, file)

    def num_params(self) -> int:
        return int(sum(int(np.prod(s)) for k, s in self.expected_shapes().items() if "running" not in k))

    def __repr__(self) -> str:
        return (f"ResnetFCHeur(in={self.in_count}x{self.one_hot_depth}, res_dim={self.res_dim}, num_blocks={self.num_blocks}, "
                f"batch_norm={self.batch_norm}, out_dim={self.out_dim}, q_fix={self.q_fix})")


def load_nnet(model_file: str, nnet: ResnetFCHeur) -> ResnetFCHeur:
    """deepxube/pytorch/nnet_utils.py:44-62 (weights_only load on the host, `module.` prefix stripped)."""
    import torch
    nnet.load_state_dict(torch.load(model_file, map_location=torch.device("cpu"), weights_only=True))
    return nnet


@deepxube_nnet_factory.register_parser("resnet_fc")
class ResnetFCParser(DelimParser):
    def __init__(self) -> None:
        super().__init__()
        self.add_argument("H", "res_dim", int, "dimensionality of hidden layers in residual blocks")
        self.add_argument("B", "num_blocks", int, "number of residual blocks")
        self.add_argument("bn", "batch_norm", None, "Batch normalization")
        self.add_argument("wn", "weight_norm", None, "Weight normalization")
        self.add_argument("ln", "layer_norm", None, "Layer normalization")

    @property
    def delim(self) -> str:
        return "_"
