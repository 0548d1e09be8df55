"""Summarise an ncu launch list (gpu__time_duration.sum CSV) per kernel: count, total, average, share."""
import collections
import csv
import sys

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0   # launches to skip (warm-up)
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
rows = list(csv.DictReader(lines))[skip:]
agg = collections.OrderedDict()
for x in rows:
    k = x['Kernel Name'].replace('void ', '').replace('papc::', '')[:64]
    v = float(x['Metric Value'].replace(',', ''))
    if x.get('Metric Unit', 'ns') in ('us', 'usecond'):
        v *= 1e3
    agg.setdefault(k, [0, 0.0])
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {tot / 1e3:.1f} us total")
for k, (n, t) in agg.items():
    print(f"{k:64s} {n:4d} {t / 1e3:10.1f} us  avg {t / n / 1e3:8.1f} us {100 * t / tot:5.1f}%")
