#!/usr/bin/env python
"""Per-layer tensor-pipe utilisation of the Mixed_4 forward contractions from an ncu pass over tools/bench_conv.py:

  ncu --kernel-name "regex:conv_bf16x3|conv3x3_halo" --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,\
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum \
      --clock-control none --csv --log-file gpurun_out/m4.csv python tools/bench_conv.py --only Mixed_4 --reps 2

usage: tensor_pipe_report.py gpurun_out/m4.csv 3 > profiles/rNN_mixed4_tensor_pipe.csv      (3 = launches per shape: warm-up + reps)
"""
import collections
import csv
import sys

from tumblr_emotions_b200.topology import MIXED

path, per_shape = sys.argv[1], int(sys.argv[2])
rows = [r for r in csv.reader(open(path)) if len(r) > 10]
hdr = rows[0]
ki, vi, mi, ui, ii = (hdr.index(k) for k in ('Kernel Name', 'Metric Value', 'Metric Name', 'Metric Unit', 'ID'))
per = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(',', ''))
    if r[mi].startswith('gpu__time'):
        v = v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
    elif 'bytes' in r[mi]:
        v = v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(r[ui], 1) / 1e6
    per.setdefault(r[ii], {'k': r[ki]})[r[mi]] = v
launches = list(per.values())
names = []
for blk in ("Mixed_4b", "Mixed_4c", "Mixed_4d", "Mixed_4e", "Mixed_4f"):
    names += [blk + s for s in (" fused1x1", " b1 3x3", " b2 3x3", " b3 1x1")]
assert len(launches) == per_shape * len(names), (len(launches), per_shape, len(names))
print("shape,mode,launch_us,tensor_pipe_pct_active,tensor_pipe_pct_elapsed,dram_MB,l2_MB")
tw = tt = 0.0
for i, n in enumerate(names):
    d = launches[per_shape * i + per_shape - 1]
    t = d['gpu__time_duration.sum']
    a = d['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']
    e = d['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed']
    mode = "halo" if "halo" in d['k'] else ("cta-pair" if "<0, 1>" in d['k'] or "false, true" in d['k'] else "single")
    print("%s,%s,%.1f,%.1f,%.1f,%.1f,%.1f" % (n, mode, t, a, e, d['dram__bytes_read.sum'] + d['dram__bytes_write.sum'], d['lts__t_bytes.sum']))
    tw += a * t
    tt += t
print("time-weighted,,%.1f,%.1f,,," % (tt, tw / tt))
