#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for r in rows[1:]:
    v = float(r[vi].replace(',', ''))
    v = v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
    agg[r[ki]][0] += 1
    agg[r[ki]][1] += v
    tot += v
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-70s n=%4d total %9.1f us  %5.1f%%" % (k.replace('<unnamed>::', '')[:70], n, t, 100 * t / tot))
print("total %.1f us over %d launches" % (tot, len(rows) - 1))
