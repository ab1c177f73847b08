"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: time share per kernel (and per phase)."""
import csv
import re
import sys
from collections import OrderedDict, defaultdict

path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [ln for ln in f if not ln.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "second": 1e9}.get(unit, 1)
    name = r["Kernel Name"]
    name = re.sub(r"\(.*", "", name)
    grid = r.get("Grid Size", "")
    rows.append((name, ns, grid))
tot = sum(ns for _, ns, _ in rows)
agg = defaultdict(lambda: [0, 0.0])
for n, ns, _ in rows:
    agg[n][0] += 1
    agg[n][1] += ns
print(f"launches: {len(rows)}   total kernel time: {tot / 1e6:.3f} ms")
print("| kernel | launches | total ms | share | avg us |")
print("|---|---|---|---|---|")
for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {n} | {c} | {ns / 1e6:.3f} | {100 * ns / tot:.1f}% | {ns / c / 1e3:.1f} |")
if len(sys.argv) > 2:
    print("\ntop individual launches:")
    for i, (n, ns, g) in sorted(enumerate(rows), key=lambda t: -t[1][1])[: int(sys.argv[2])]:
        print(f"  #{i} {n} grid={g} {ns / 1e3:.1f} us")
