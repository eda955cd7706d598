"""Command line: `python -m deepxube_b200 {solve,time,problem_inst,pathfind_info,domain_info}` -- the hot-path
subset of deepxube/_cli.py:470-582 with the same flags."""
import argparse
import os
import pickle
import time
from typing import Dict, List

import numpy as np

from deepxube_b200._solve import parse_solve
from deepxube_b200.factories.domain_factory import domain_factory, get_domain_from_arg
from deepxube_b200.factories.pathfinding_factory import pathfinding_factory


def problem_inst_gen(args: argparse.Namespace) -> None:
    """deepxube/_cli.py:448-467"""
    if os.path.isfile(args.file) and (not args.redo):
        print(f"File {args.file} already exists and redo not set. Not generating data.")
        return
    domain, _ = get_domain_from_arg(args.domain)
    num_steps_l: List[int] = list(np.random.randint(args.step_min, args.step_max + 1, size=args.num))
    print(f"Generating {args.num} states")
    t0 = time.time()
    states, goals = domain.sample_problem_instances(num_steps_l)
    print(f"Time: {time.time() - t0}")
    print(f"Saving data to {args.file}")
    data: Dict = {"states": states, "goals": goals}
    from deepxube_b200 import compat_pickle
    compat_pickle.dump(data, args.file, ref_classes=getattr(args, "ref_pickle", False))


def time_test_args(args: argparse.Namespace) -> None:
    """Throughput printouts of the domain kernels and one heuristic call (deepxube/tests/time_tests.py:30-246)."""
    from deepxube_b200.factories.pathfind_fns_factory import get_path_fns_nnet_par_dict
    domain, domain_name = get_domain_from_arg(args.domain)
    n = args.num_insts
    steps = list(np.random.randint(args.step_min, args.step_max + 1, size=n))
    t0 = time.time()
    states, goals = domain.sample_problem_instances(steps)
    dt = time.time() - t0
    print(f"Generated {n} problem instances in {dt:.4f} seconds ({n / dt:.2f}/second)")
    for name, fn in (("next_state", lambda: domain.next_state(states, domain.sample_state_action(states))),
                     ("is_solved", lambda: domain.is_solved(states, goals)),
                     ("expand", lambda: domain.expand(states))):
        fn()
        t0 = time.time()
        fn()
        dt = time.time() - t0
        print(f"{name}: {n} states in {dt:.4f} seconds ({n / dt:.2f}/second)")
    if args.fn:
        pfns, _ = get_path_fns_nnet_par_dict(domain, domain_name, args.fn)
        for field in ("heurv", "heurq"):
            fn = getattr(pfns, field, None)
            if fn is None:
                continue
            call = (lambda: fn(states, goals, [None] * n)) if field == "heurv" else \
                (lambda: fn(states, goals, domain.get_state_actions(states), [None] * n))
            call()
            t0 = time.time()
            call()
            dt = time.time() - t0
            print(f"Computed heuristic ({field}) for {n} states in {dt:.4f} seconds ({n / dt:.2f}/second)")


def main() -> None:
    parser = argparse.ArgumentParser(prog="deepxube_b200")
    sub = parser.add_subparsers(dest="cmd", required=True)
    parse_solve(sub.add_parser("solve", help="Solve problem instances with batch weighted A* / Q* on the GPU"))
    p = sub.add_parser("time", help="Time basic domain functionality and heuristic functions")
    p.add_argument('--domain', type=str, required=True)
    p.add_argument('--fn', type=str, nargs='*')
    p.add_argument('--num_insts', type=int, default=10)
    p.add_argument('--step_min', type=int, default=0)
    p.add_argument('--step_max', type=int, default=100)
    p.set_defaults(func=time_test_args)
    p = sub.add_parser("problem_inst", help="Generate problem instances")
    p.add_argument('--domain', type=str, required=True)
    p.add_argument('--step_min', type=int, default=0)
    p.add_argument('--step_max', type=int, required=True)
    p.add_argument('--num', type=int, required=True)
    p.add_argument('--file', type=str, required=True)
    p.add_argument('--redo', action='store_true', default=False)
    p.add_argument('--ref_pickle', action='store_true', default=False,
                   help="(B200) write the file with the reference's class paths (deepxube.domains.*)")
    p.set_defaults(func=problem_inst_gen)
    p = sub.add_parser("domain_info", help="List registered domains")
    p.set_defaults(func=lambda a: print("\n".join(domain_factory.get_all_class_names())))
    p = sub.add_parser("pathfind_info", help="List registered pathfinding algorithms")
    p.set_defaults(func=lambda a: print("\n".join(f"{n}: {pathfinding_factory.get_type(n).description()}"
                                                  for n in pathfinding_factory.get_all_class_names())))
    args = parser.parse_args()
    args.func(args)


if __name__ == "__main__":
    main()
