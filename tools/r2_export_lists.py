"""Exports the tracked launch lists of the final round-2 kernels from the ncu CSVs of tools/r2_final_profile.sh:
profiles/r2_{c2,q}_last_iteration_ncu.csv (the launches between the last two k_end_iter) and profiles/r2f_launch_list_ncu.csv (one
whole C2 bench step).  usage: r2_export_lists.py [tag]"""
import csv, os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from launch_list_summary import load

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2f"


def last_iteration(seq):
    idx = [i for i, (n, _) in enumerate(seq) if "k_end_iter" in n]
    return seq[idx[-2] + 1: idx[-1] + 1]


for name, src in (("c2", f"{tag}_c2_launches.csv"), ("q", f"{tag}_q_launches.csv")):
    seq = load(os.path.join(ROOT, "gpurun_out", src))
    with open(os.path.join(ROOT, "profiles", f"r2_{name}_last_iteration_ncu.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["launch", "kernel", "duration_us"])
        for i, (n, t) in enumerate(last_iteration(seq)):
            w.writerow([i, n, f"{t:.2f}"])
    if name == "c2":                      # the last whole step: from the last k_closed_init on
        starts = [i for i, (n, _) in enumerate(seq) if n.startswith("k_closed_init")]
        with open(os.path.join(ROOT, "profiles", "r2f_launch_list_ncu.csv"), "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["launch", "kernel", "duration_us"])
            for i, (n, t) in enumerate(seq[starts[-1]:]):
                w.writerow([i, n, f"{t:.2f}"])
    print(name, len(seq), "launches;", len(last_iteration(seq)), "in the last iteration,", f"{sum(t for _, t in last_iteration(seq)):.1f} us")
