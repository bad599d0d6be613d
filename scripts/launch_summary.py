"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list:
    python scripts/launch_summary.py launches.csv [skip_first_n] > summary.txt"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = collections.OrderedDict()
for row in rows[skip:]:
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void dicp::", "").replace("void ", "")[:90]
    t = float(row["Metric Value"].replace(",", ""))
    if row["Metric Unit"] == "ns":
        t /= 1e3
    elif row["Metric Unit"] == "ms":
        t *= 1e3
    a = agg.setdefault((name, row.get("Grid Size", "")), [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(a[1] for a in agg.values())
print(f"# {sys.argv[1]}: {len(rows) - skip} launches, {tot / 1e3:.3f} ms of kernel time (serialised, cold caches: shares matter, not absolutes)")
print(f"{'share':>7} {'total us':>12} {'count':>6} {'us/launch':>10}  kernel  grid")
for (name, grid), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{100 * a[1] / tot:6.2f}% {a[1]:12.1f} {a[0]:6d} {a[1] / a[0]:10.2f}  {name}  {grid}")
