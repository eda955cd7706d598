"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel times of the last complete search
iteration (the launches between the last two k_end_iter) and totals per kernel family.  usage: launch_list_summary.py <csv> [marker]"""
import collections
import csv
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    seq = []
    for x in csv.DictReader(lines):
        t = float(x["Metric Value"].replace(",", ""))
        t = t / 1000 if x["Metric Unit"] == "ns" else (t * 1000 if x["Metric Unit"] == "ms" else t)
        seq.append((x["Kernel Name"].split("(")[0].replace("void ", ""), t))
    return seq


if __name__ == "__main__":
    seq = load(sys.argv[1])
    marker = sys.argv[2] if len(sys.argv) > 2 else "k_end_iter"
    idx = [i for i, (n, _) in enumerate(seq) if marker in n]
    a, b = (idx[-2] + 1, idx[-1] + 1) if len(idx) >= 2 else (0, len(seq))
    tot = sum(t for _, t in seq[a:b])
    print(f"{len(seq)} launches in the list; last iteration = launches {a}..{b - 1}, {tot:.1f} us (serialised, cold-cache ncu times)")
    for n, t in seq[a:b]:
        print(f"{t:9.1f} us {100 * t / tot:5.1f} %  {n[:100]}")
    fam = collections.OrderedDict()
    for n, t in seq[a:b]:
        key = n.split("<")[0]
        fam[key] = fam.get(key, 0.0) + t
    print("per kernel family:")
    for k, v in sorted(fam.items(), key=lambda kv: -kv[1]):
        print(f"{v:9.1f} us {100 * v / tot:5.1f} %  {k}")
