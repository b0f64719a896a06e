"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: mean device time per kernel."""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    name = row['Kernel Name'].split('(')[0].replace('void ', '')
    v = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit']
    v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
    agg.setdefault(name, []).append(v)
tot = sum(sum(v) for v in agg.values())
for k, v in agg.items():
    print(f"{k[:44]:44s} n={len(v):3d} mean={sum(v)/len(v):10.1f} us  share={100*sum(v)/tot:5.1f}%")
