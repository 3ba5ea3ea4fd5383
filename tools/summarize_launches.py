"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name count, total and mean time."""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    name = r[ki].split("(")[0].replace("void ", "")[:90]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"| kernel | launches | total us | mean us | share |\n|---|---|---|---|---|")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name}` | {n} | {t:.1f} | {t / n:.2f} | {100 * t / tot:.1f} % |")
print(f"| total | {sum(a[0] for a in agg.values())} | {tot:.1f} | | |")
