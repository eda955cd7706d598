"""Turns the two ncu outputs of a profiling gpurun into the tracked summaries under profiles/.

  launch list : ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file <launches.csv> python bench.py ...
  full capture: ncu --set full --clock-control none --import-source on -k regex:k_gemm_bf16_tc -s S -c C -o <rep> python bench.py ...
                ncu -i <rep>.ncu-rep --page raw --csv > <raw.csv>        (run here; the .ncu-rep stays in gpurun_out/)

usage: profile_summary.py <tag> <launches.csv> <raw.csv> <rows_in_captured_launch> "<command>"
writes profiles/<tag>_launch_list_ncu.csv (timed step only), profiles/<tag>_summary.md, profiles/<tag>_gemm_traffic.json"""
import csv, json, os, sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches_csv, raw_csv, rows, command = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), sys.argv[5]


def read_ncu_csv(path):
    lines = open(path, errors="replace").read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    return list(csv.DictReader(lines[start:]))


def short(name):
    name = name.replace("void ", "")
    i = name.find("(")
    return name[:i] if i > 0 else name


# ---- launch list: keep the SECOND search step (steps start with k_closed_init; the first is the warm-up) ----
recs = [r for r in read_ncu_csv(launches_csv) if r.get("Metric Name") == "gpu__time_duration.sum"]
starts = [i for i, r in enumerate(recs) if short(r["Kernel Name"]) == "k_closed_init"]
seg = recs[starts[1]:starts[2]] if len(starts) > 2 else recs[starts[-1]:]
out_list = os.path.join(ROOT, "profiles", f"{tag}_launch_list_ncu.csv")
with open(out_list, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["launch", "kernel", "grid", "block", "duration_ns"])
    for i, r in enumerate(seg):
        ns = float(r["Metric Value"].replace(",", "")) * (1e3 if r["Metric Unit"] in ("usecond", "us") else 1.0)
        w.writerow([i, short(r["Kernel Name"]), r.get("Grid Size", ""), r.get("Block Size", ""), f"{ns:.0f}"])
agg = OrderedDict()
for r in seg:
    ns = float(r["Metric Value"].replace(",", "")) * (1e3 if r["Metric Unit"] in ("usecond", "us") else 1.0)
    k = short(r["Kernel Name"])
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += ns
total = sum(v[1] for v in agg.values())

# ---- full capture of the GEMM launches ----
raw = read_ncu_csv(raw_csv)
wanted = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
          "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
          "launch__cluster_size", "launch__block_size", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
          "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max"]
units = raw[0] if raw and raw[0].get("ID", "") == "" else None      # the first data row of a raw page carries the units
rows_raw = raw[1:] if units else raw


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


md = [f"# {tag}: ncu profiles of the bench step (B200, sm_100a)", "",
      "Command (exited 0 without ncu first, in the same gpurun call):", f"`{command}`", "",
      "## 1. Launch list of one search step (`ncu --metrics gpu__time_duration.sum --clock-control none`)", "",
      f"Raw per-launch list: `profiles/{tag}_launch_list_ncu.csv` ({len(seg)} launches = one step; cold-cache, serialised:",
      "compare SHARES with the live CUDA-event numbers of bench.py, not absolute times).", "",
      f"Total {total / 1e3:.0f} us over {len(seg)} launches.", "", "| kernel | launches | us | share |", "|---|---:|---:|---:|"]
for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    md.append(f"| `{k}` | {n} | {ns / 1e3:.1f} | {100 * ns / total:.1f} % |")
md += ["", f"## 2. `ncu --set full` of the GEMM launches of the last search iteration (~{rows} rows)", "",
       "| launch | kernel | " + " | ".join(m.split(".")[0].replace("__", " ") for m in wanted) + " |",
       "|---|---|" + "---:|" * len(wanted)]
traffic = []
for i, r in enumerate(rows_raw):
    vals = []
    for m in wanted:
        v = r.get(m, "")
        u = units.get(m, "") if units else ""
        vals.append(f"{v} {u}".strip())
    md.append(f"| {i} | `{short(r.get('Kernel Name', ''))}` | " + " | ".join(vals) + " |")
    try:
        rd = to_bytes(r["dram__bytes_read.sum"], units["dram__bytes_read.sum"])
        wr = to_bytes(r["dram__bytes_write.sum"], units["dram__bytes_write.sum"])
        dur = r["gpu__time_duration.sum"] + " " + units["gpu__time_duration.sum"]
        traffic.append({"launch": i, "kernel": short(r.get("Kernel Name", "")), "dram_read_bytes": rd, "dram_write_bytes": wr,
                        "duration": dur, "tensor_pipe_active_pct": r.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")})
    except Exception:
        pass
open(os.path.join(ROOT, "profiles", f"{tag}_summary.md"), "w").write("\n".join(md) + "\n")
hidden = [t for t in traffic if t["kernel"].endswith("0, 0>")]      # plain hidden-layer launches
sel = hidden if hidden else traffic
if sel:
    Hp = 1024
    js = {"source": f"ncu --set full, {command}", "rows_in_launch": rows,
          "launches": traffic,
          "traffic_per_launch_avg_bytes": sum(t["dram_read_bytes"] + t["dram_write_bytes"] for t in sel) / len(sel),
          "algorithmic_bytes": {"hidden": rows * Hp * 2 * 2 + Hp * Hp * 2, "hidden_residual": rows * Hp * 2 * 3 + Hp * Hp * 2}}
    json.dump(js, open(os.path.join(ROOT, "profiles", f"{tag}_gemm_traffic.json"), "w"), indent=1)
print("\n".join(md[:40]))
