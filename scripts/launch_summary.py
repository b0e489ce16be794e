"""Summarise an ncu --csv launch list (gpu__time_duration.sum per launch) by kernel name."""
import csv
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith('==')]
rd = csv.DictReader(lines)
per = defaultdict(lambda: defaultdict(dict))
for r in rd:
    per[r['ID']][r['Metric Name']] = r['Metric Value']
    per[r['ID']]['name'] = r['Kernel Name']
agg = defaultdict(lambda: [0.0, 0, 0.0])
detail = []
for i, d in per.items():
    t = float(str(d.get('gpu__time_duration.sum', '0')).replace(',', '')) / 1e3  # ns -> us
    n = d['name'].split('(')[0]
    agg[n][0] += t
    agg[n][1] += 1
    agg[n][2] = max(agg[n][2], t)
    detail.append((int(i), n, t, d.get('launch__grid_size'), d.get('launch__block_size')))
tot = sum(a[0] for a in agg.values())
print('total kernel time %.1f us over %d launches' % (tot, len(per)))
for n, (t, c, mx) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print('%9.1f us %5.1f%%  x%-4d max %8.1f  %s' % (t, 100 * t / tot, c, mx, n[:90]))
if len(sys.argv) > 2:
    pat = sys.argv[2]
    for i, n, t, g, b in sorted(detail):
        if pat in n:
            print(i, '%.1f us' % t, 'grid', g, 'block', b)
