"""Function registries and `--fn` resolution (mirror of deepxube/factories/pathfind_fns_factory.py:15-116).

`--fn <fn>[,<nnet_input>],<nnet>[,<file>]`: <fn> in {heurv, heurq_fixout}; without a network the reference's
all-zero default function is used; with a network but no file the network is randomly initialised."""
from typing import Any, Dict, List, Optional, Tuple

from deepxube_b200.base.factory import Factory, FactoryAutoBuild
from deepxube_b200.base.nnet_input import get_domain_nnet_input_keys, get_nnet_input_t
from deepxube_b200.utils.command_line_utils import get_name_args

deepxube_nnet_par_factory: Factory = Factory("DeepXubeNNetPar")
pathfind_fns_factory: FactoryAutoBuild = FactoryAutoBuild("PathFindFNs")


def get_compat_nnet_input_name(domain: Any, domain_name: str, nnet_par_t: type, nnet_name_args: str, nnet_par_name_args: str) -> str:
    reasons: List[str] = []
    for key in get_domain_nnet_input_keys(domain_name):
        reason = nnet_par_t.get_incompat_reason(domain, get_nnet_input_t(domain_name, key), None)
        if reason is None:
            return key
        reasons.append(f"{reason} (NNetInput name: {key})")
    raise ValueError(f"Cannot build fn for domain: {domain_name}, nnet_args: {nnet_name_args}, nnet_par_args: {nnet_par_name_args}."
                     "\nIncompatibility reasons:\n" + "\n".join(reasons))


def get_dx_nnet_par(domain: Any, domain_name: str, nnet_input_name_args: Optional[str], nnet_par_name_args: str,
                    nnet_name_args: Optional[str]) -> Tuple[Any, str]:
    par_name, par_args = get_name_args(nnet_par_name_args)
    par_t = deepxube_nnet_par_factory.get_type(par_name)
    if nnet_name_args is not None and nnet_input_name_args is None:
        nnet_input_name_args = get_compat_nnet_input_name(domain, domain_name, par_t, nnet_name_args, nnet_par_name_args)
    kwargs = deepxube_nnet_par_factory.get_kwargs(par_name, par_args)
    kwargs.update(domain=domain, domain_name=domain_name, nnet_input_name_args=nnet_input_name_args, nnet_name_args=nnet_name_args)
    return deepxube_nnet_par_factory.build_class(par_name, kwargs), par_name


def get_path_fns_nnet_par_dict(domain: Any, domain_name: str, fn_name_args_l: List[str], device: Any = None,
                               nnet_files: Optional[List[Optional[str]]] = None, nnet_batch_size: Optional[int] = None,
                               precision: Optional[str] = None) -> Tuple[Any, Dict[str, Any]]:
    from deepxube_b200.nnets.resnet_fc import load_nnet
    if nnet_files is not None:
        assert len(nnet_files) == len(fn_name_args_l)
    fn_dict: Dict[str, Any] = {}
    par_dict: Dict[str, Any] = {}
    for idx, fn_arg in enumerate(fn_name_args_l):
        parts = fn_arg.split(",")
        nnet_input_name_args: Optional[str] = None
        nnet_name_args: Optional[str] = None
        if len(parts) == 1:
            par_name_args = parts[0]
        elif len(parts) == 2:
            par_name_args, nnet_name_args = parts
        elif len(parts) == 3:
            par_name_args, nnet_input_name_args, nnet_name_args = parts
        else:
            raise ValueError("Each element of fn_name_args_l must be either <fn>, <fn>,<nnet>, or <fn>,<nnet_input>,<nnet>")
        par, _ = get_dx_nnet_par(domain, domain_name, nnet_input_name_args, par_name_args, nnet_name_args)
        if nnet_name_args is not None:
            nnet = par.get_nnet()
            nnet_file = None if nnet_files is None else nnet_files[idx]
            if nnet_file is not None:
                par.set_nnet_file(nnet_file)
                load_nnet(nnet_file, nnet)
            fn = par.get_nnet_fn(nnet, nnet_batch_size, precision)
        else:
            fn = par.get_default_fn()
        fn_dict[par.get_field_name()] = fn
        par_dict[par.get_field_name()] = par
    return pathfind_fns_factory.build_class(fn_dict), par_dict
