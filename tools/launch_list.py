"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one train step."""
import csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
rows = []
for row in csv.DictReader(lines):
    if row.get('Metric Name') == 'gpu__time_duration.sum':
        v = float(row['Metric Value'].replace(',', ''))
        if row['Metric Unit'] == 'ns': v /= 1000.0
        elif row['Metric Unit'] == 'ms': v *= 1000.0
        rows.append((row['Kernel Name'], row['Grid Size'], v))
idx = [i for i, r in enumerate(rows) if 'pack_input' in r[0]]
step = rows[idx[-2]:idx[-1]] if len(idx) >= 2 else rows
print('launches', len(step), 'sum us %.1f' % sum(t for _, _, t in step))
agg = {}
for k, g, t in step:
    k = re.sub(r'segb::', '', k).split('(')[0].replace('void ', '')
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += t
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-46s %3d %8.1f' % (k[:46], n, t))
if '-v' in sys.argv:
    for k, g, t in step:
        k = re.sub(r'segb::', '', k).split('(')[0].replace('void ', '')
        print('  %-44s %-14s %7.1f' % (k[:44], g, t))
