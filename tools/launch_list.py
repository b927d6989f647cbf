"""Per-kernel totals of an ncu launch list (csv with gpu__time_duration.sum): python tools/launch_list.py launches.csv"""
import csv
import sys
from collections import OrderedDict

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = None, OrderedDict()
for r in rows:
    if 'Kernel Name' in r:
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    data.setdefault((int(d['ID']), d['Kernel Name'][:60]), {})[d['Metric Name']] = d['Metric Value']
agg = {}
for (_, name), v in data.items():
    t = float(v['gpu__time_duration.sum'].replace(',', '')) / 1000
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(a[1] for a in agg.values())
for name, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{name:62s} n={n:3d} total {t:8.1f} us  avg {t / n:7.1f}  share {t / tot:.3f}")
print('total', round(tot, 1), 'us')
