"""Top stall sites of a kernel from `ncu -i <rep> --page source --csv` (SASS view).  usage: ncu_sass_hot.py <csv> [top_n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0] != "Address"]
tot = sum(int(r[ix["# Samples"]]) for r in data)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
print("kernel:", rows[0][1][:100] if rows[0] else "?")
print("total samples", tot, "instructions", len(data))
print("stall mix:", [(k, round(100 * v / max(tot, 1), 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]])
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:top_n]:
    st = {s: int(r[ix[s]]) for s in stalls if int(r[ix[s]] or 0) > 0}
    print(f"{int(r[ix['# Samples']]):6d} {100 * int(r[ix['# Samples']]) / max(tot, 1):5.1f}%  {r[1].strip()[:64]:64s} "
          f"{sorted(st.items(), key=lambda kv: -kv[1])[:2]} exec {r[ix['Instructions Executed']]}")
