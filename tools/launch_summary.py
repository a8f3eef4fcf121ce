#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.  usage: launch_summary.py file.csv [nsearches]"""
import csv, sys, re, collections
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith('==')]
rd = csv.DictReader(lines)
agg = collections.OrderedDict()
for r in rd:
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = r['Kernel Name']
    name = re.sub(r'\(.*$', '', name)
    name = name.replace('ckm::', '').replace('void ', '')
    v = float(r['Metric Value'].replace(',', ''))
    unit = r['Metric Unit']
    ms = v / 1e6 if unit in ('ns', 'nsecond') else (v / 1e3 if unit in ('us', 'usecond') else v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += ms
tot = sum(a[1] for a in agg.values())
print('| kernel | launches | total ms | share |\n|---|---:|---:|---:|')
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('| `%s` | %d | %.1f | %.1f %% |' % (k, a[0], a[1], 100 * a[1] / tot))
print('| total | %d | %.1f | |' % (sum(a[0] for a in agg.values()), tot))
