#!/usr/bin/env python
"""Aggregate an ncu launch list (gpu__time_duration.sum [+ dram__bytes_read/write.sum]) by kernel and, with --traffic-json,
write the per-launch DRAM traffic of conv_bf16x3_kernel that bench.py reports as roofline.traffic."""
import collections
import csv
import json
import sys

path = sys.argv[1]
rows = [r for r in csv.reader(open(path)) if len(r) > 10]
hdr = rows[0]
ki, vi, mi, ui, ii = (hdr.index(k) for k in ('Kernel Name', 'Metric Value', 'Metric Name', 'Metric Unit', 'ID'))
per = collections.defaultdict(dict)
for r in rows[1:]:
    v, u = float(r[vi].replace(',', '')), r[ui]
    if r[mi].startswith('gpu__time'):
        v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)                       # us
    else:
        v = v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)            # bytes
    per[r[ii]][r[mi]] = v
    per[r[ii]]['k'] = r[ki]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for d in per.values():
    k = d['k'].replace('<unnamed>::', '').replace('void ', '')[:56]
    agg[k][0] += 1
    agg[k][1] += d['gpu__time_duration.sum']
    agg[k][2] += d.get('dram__bytes_read.sum', 0.0) + d.get('dram__bytes_write.sum', 0.0)
tot = sum(a[1] for a in agg.values())
print("%-58s %5s %10s %6s %10s %8s" % ("kernel", "n", "us", "share", "dram MB", "GB/s"))
for k, (n, t, b) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-58s %5d %10.1f %5.1f%% %10.1f %8.0f" % (k, n, t, 100 * t / tot, b / 1e6, b / t / 1e3 if t else 0))
print("total %.1f us over %d launches" % (tot, len(per)))
if len(sys.argv) > 3 and sys.argv[2] == '--traffic-json':
    conv = [d for d in per.values() if 'conv_bf16x3' in d['k'] or 'conv3x3_halo' in d['k']]
    tb = sum(d['dram__bytes_read.sum'] + d['dram__bytes_write.sum'] for d in conv)
    json.dump({"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over every conv_bf16x3_kernel / conv3x3_halo_kernel launch of one joint "
                         "training step, batch 256 (%s)" % path, "launches": len(conv), "dram_bytes_total": tb,
               "dram_bytes_per_launch": tb / len(conv)}, open(sys.argv[3], 'w'), indent=1)
