#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and launch
count per kernel, share of the total.  usage: scripts/launch_summary.py launches.csv [steps]"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
hdr = rows[0]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(',', ''))
    except ValueError:
        continue
    v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(r[ui], 1.0)
    d = agg.setdefault(r[ki], [0, 0.0]); d[0] += 1; d[1] += v
tot = sum(d[1] for d in agg.values())
print(f"total {tot/1e3/steps:.3f} ms per step over {steps} step(s), {sum(d[0] for d in agg.values())//steps} launches per step")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t/1e3/steps:9.3f} ms {c//steps:4d}x {100*t/tot:5.1f}%  {k[:100]}")
