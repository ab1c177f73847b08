"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the SECOND half of the launches
(tools/*_once.py run a warm-up pass and then the measured pass).  usage: launch_summary.py list.csv [--list N]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
H = rows[hdr]
ki, vi = H.index("Kernel Name"), H.index("Metric Value")
data = [(r[ki], float(r[vi].replace(",", ""))) for r in rows[hdr + 2:] if len(r) > vi]
second = data[len(data) // 2:]
agg = collections.OrderedDict()
for k, v in second:
    a = agg.setdefault(k.split("(")[0], [0, 0.0])
    a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"launches: {len(second)}   total kernel time: {tot / 1e6:.3f} ms")
print("| kernel | launches | total ms | share | avg us |\n|---|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {k} | {a[0]} | {a[1] / 1e6:.3f} | {100 * a[1] / tot:.1f}% | {a[1] / a[0] / 1e3:.1f} |")
if "--list" in sys.argv:
    n = int(sys.argv[sys.argv.index("--list") + 1])
    for i, (k, v) in enumerate(second[:n]):
        print(i, k.split("(")[0][-40:], f"{v / 1e3:.1f}")
