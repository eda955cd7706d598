"""`solve` entry point: same flags, loop, result keys and output lines as deepxube/_solve.py:45-224
(one instance at a time, a fresh PathFind per instance, results.pkl rewritten after every instance so that a
rerun resumes at len(results["actions"]) unless --redo)."""
import argparse
import os
import pickle
import sys
import time
from argparse import ArgumentParser
from typing import Any, Dict, List, Optional

import numpy as np

from deepxube_b200 import compat_pickle
from deepxube_b200.base.pathfinding import get_path
from deepxube_b200.factories.domain_factory import get_domain_from_arg
from deepxube_b200.factories.pathfind_fns_factory import get_path_fns_nnet_par_dict
from deepxube_b200.factories.pathfinding_factory import get_pathfind_from_arg
from deepxube_b200.pathfinding.utils import is_valid_soln
from deepxube_b200.utils import data_utils
from deepxube_b200.utils.command_line_utils import print_command


def parse_solve(parser: ArgumentParser) -> None:
    parser.add_argument('--domain', type=str, required=True, help="Domain name and arguments.")
    parser.add_argument('--fn', type=str, nargs='*', help="Function, neural network arguments, and neural network file separated by a comma.")
    parser.add_argument('--nnet_batch_size', type=int, default=None, help="Maximum number of inputs to give to any nnet at a time "
                                                                          "through the host call path. None means no limit.")
    parser.add_argument('--pathfind', type=str, required=True, help="Pathfinding algorithm and arguments.")
    parser.add_argument('--time_limit', type=float, default=-1.0, help="A time limit (in seconds) for search. Default is -1, which means infinite.")
    parser.add_argument('--max_itrs', type=int, default=None, help="Maximum number of search iterations. None for infinite.")
    parser.add_argument('--file', type=str, required=True, help="File containing problem instances to solve")
    parser.add_argument('--results', type=str, required=True, help="Directory to save results. Saves results after every instance.")
    parser.add_argument('--start_idx', type=int, default=None, help="Index of instance at which to start. Useful for debugging.")
    parser.add_argument('--redo', action='store_true', default=False, help="Set to redo already completed instances")
    parser.add_argument('--verbose', action='store_true', default=False, help="Set for verbose")
    parser.add_argument('--debug', action='store_true', default=False, help="Set when debugging with breakpoints")
    parser.add_argument('--precision', type=str, default=None, choices=["fp32", "bf16"],
                        help="(B200) arithmetic of the device network: fp32 parity path (default) or bf16 tensor cores")
    parser.add_argument('--max_nodes', type=int, default=None, help="(B200) node-store capacity of the engine")
    parser.add_argument('--ref_pickle', action='store_true', default=False,
                        help="(B200) write results.pkl with the reference's class paths (deepxube.domains.*) so that the "
                             "reference's own tooling opens it")
    parser.set_defaults(func=solve_cli)


def split_fn_args(fn_args: List[str]):
    fns: List[str] = []
    files: List[Optional[str]] = []
    for fn in fn_args:
        parts = fn.split(",")
        if len(parts) == 1:
            fns.append(parts[0]); files.append(None)
        elif len(parts) == 3:
            fns.append(f"{parts[0]},{parts[1]}"); files.append(parts[2])
        elif len(parts) == 4:
            fns.append(f"{parts[0]},{parts[1]},{parts[2]}"); files.append(parts[3])
        else:
            raise ValueError("--fn must be either --fn <fn>, --fn <fn>,<nnet>,<nnet_file>, or --fn <fn>,<nnet_input>,<nnet>,<nnet_file>")
    return fns, files


def solve_cli(args: argparse.Namespace) -> None:
    os.makedirs(args.results, exist_ok=True)
    data: Dict = compat_pickle.load(args.file)
    states, goals = data['states'], data['goals']
    results_file, output_file = f"{args.results}/results.pkl", f"{args.results}/output.txt"
    results: Dict[str, Any]
    if os.path.isfile(results_file) and not args.redo:
        results = compat_pickle.load(results_file)
        if not args.debug:
            sys.stdout = data_utils.Logger(output_file, "a")
    else:
        results = {"states": states, "goals": goals, "actions": [], "states_on_path": [], "path_costs": [], "iterations": [],
                   "times": [], "itrs/sec": [], "num_nodes_generated": [], "solved": []}
        if not args.debug:
            sys.stdout = data_utils.Logger(output_file, "w")
    print_command()
    from deepxube_b200 import _lib
    print(f"device: cuda:{_lib.current_device()} (libdxb200, sm_100a), on_gpu: True")
    domain, domain_name = get_domain_from_arg(args.domain)
    fns, nnet_files = split_fn_args(args.fn or [])
    pathfind_fns, nnet_par_dict = get_path_fns_nnet_par_dict(domain, domain_name, fns, None, nnet_files=nnet_files,
                                                             nnet_batch_size=args.nnet_batch_size, precision=args.precision)
    for par in nnet_par_dict.values():
        print(par)
        print(f"(name: {par.get_field_name()}, nnet_input_name_args: {par.nnet_input_name_args})")
    print(pathfind_fns)
    pathfind, pathfind_name, _ = get_pathfind_from_arg(domain, pathfind_fns, args.pathfind)
    print(pathfind, f"(name: {pathfind_name})")
    pathfind.close()

    start_idx = args.start_idx if args.start_idx is not None else len(results["actions"])
    for state_idx in range(start_idx, len(states)):
        state, goal = states[state_idx], goals[state_idx]
        pathfind = get_pathfind_from_arg(domain, pathfind_fns, args.pathfind)[0]
        if args.max_nodes is not None:
            pathfind._max_nodes = args.max_nodes
        # the reference's "Times - heur: ..., expand: ..., filt: ..., cost: ..., pushpop: ..." line, from device stage timers
        pathfind.stage_times_on = os.environ.get("DXB200_STAGE_TIMES", "1") == "1"
        start_time = time.time()
        num_itrs = 0
        instance = pathfind.make_instances([state], [goal], None, True)[0]
        pathfind.add_instances([instance])
        while not min(x.finished() for x in pathfind.instances):
            pathfind.step(verbose=args.verbose)
            num_itrs += 1
            if (args.time_limit >= 0) and ((time.time() - start_time) > args.time_limit):
                break
            if (args.max_itrs is not None) and (num_itrs == args.max_itrs):
                break
        solve_time = time.time() - start_time

        solved = False
        path_states, path_actions, path_cost = None, None, np.inf
        itrs_per_sec = num_itrs / solve_time
        num_nodes_gen = pathfind.instances[0].num_nodes_generated
        goal_node = pathfind.instances[0].goal_node
        if goal_node is not None:
            path_states, path_actions, _, path_cost = get_path(goal_node)
            solved = True
            assert is_valid_soln(state, goal, path_actions, domain)
        results["actions"].append(path_actions)
        results["states_on_path"].append(path_states)
        results["path_costs"].append(path_cost)
        results["iterations"].append(num_itrs)
        results["itrs/sec"].append(itrs_per_sec)
        results["times"].append(solve_time)
        results["num_nodes_generated"].append(num_nodes_gen)
        results["solved"].append(solved)
        print(f"State: %i, SolnCost: %.2f, # Nodes Gen: %s, Itrs: %i, Itrs/sec: %.2f, Solved: {solved}, "
              f"Time: %.2f" % (state_idx, path_cost, format(num_nodes_gen, ","), num_itrs, itrs_per_sec, solve_time))
        print("Times - %s, num_itrs: %i" % (pathfind.times.get_time_str(), num_itrs))
        print("Means - SolnCost: %.2f, # Nodes Gen: %.2f, Itrs: %.2f, Itrs/sec: %.2f, Solved: %.2f%%, "
              "Time: %.2f" % (_get_mean(results, "path_costs"), _get_mean(results, "num_nodes_generated"),
                              _get_mean(results, "iterations"), _get_mean(results, "itrs/sec"),
                              float(100.0 * np.mean(results["solved"])), _get_mean(results, "times")))
        print("")
        pathfind.close()
        compat_pickle.dump(results, results_file, ref_classes=getattr(args, "ref_pickle", False))


def _get_mean(results: Dict[str, Any], key: str) -> float:
    vals = [x for x, solved in zip(results[key], results["solved"]) if solved]
    return float(np.mean(vals)) if vals else 0
