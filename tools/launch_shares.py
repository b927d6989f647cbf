"""Per-kernel launch count, mean duration and share of the total from an ncu launch list (--metrics gpu__time_duration.sum --csv):
python tools/launch_shares.py profiles/rXX/launches.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[start]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[start + 1:]:
    if len(r) > vi:
        try:
            agg[r[ki].split("(")[0]].append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
tot = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k[:70]:70s} n={len(v):4d} avg_us={sum(v) / len(v) / 1e3:9.2f} share={sum(v) / tot:6.3f}")
