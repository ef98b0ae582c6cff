"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel.
usage: python profiles/launch_summary.py gpurun_out/<tag>_launches.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hi]
ki, vi, ui = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(',', ''))
    except ValueError:
        continue
    a = agg.setdefault(r[ki].split('(')[0][:64], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print("%-66s %6s %14s %10s %6s" % ("kernel", "n", "total_" + rows[hi + 1][ui], "avg", "share"))
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-66s %6d %14.0f %10.0f %5.1f%%" % (k, n, t, t / n, 100 * t / tot))
