"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) over the WHOLE capture into per-kernel totals,
split into libnfb200 kernels (namespace nfb::) and library kernels (cuDNN / cuBLAS / ATen).
Usage: summarize_all_launches.py launches.csv out.md "title" """
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr, data = rows[hi], rows[hi + 1:]
kn, mv, mu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg, tot, ours = collections.OrderedDict(), 0.0, 0.0
n_ours = 0
for r in data:
    if len(r) <= mv:
        continue
    v = float(r[mv].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}[r[mu]]
    mine = 'nfb::' in r[kn]
    k = re.sub(r'^void ', '', re.sub(r'\(.*', '', r[kn])).replace('nfb::', '')
    k = re.sub(r'<.*', '<...>', k) if not mine else k
    a = agg.setdefault((mine, k), [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
    if mine:
        ours += v
        n_ours += 1
title = sys.argv[3] if len(sys.argv) > 3 else 'launch list'
lines = ['# ' + title + ' (ncu, cold-cache, serialised: compare SHARES)', '',
         '%d launches, %.1f us total; libnfb200: %d launches, %.1f us (%.1f%%); library (cuDNN/cuBLAS/ATen): %d launches, %.1f us (%.1f%%)'
         % (len(data), tot, n_ours, ours, 100 * ours / tot, len(data) - n_ours, tot - ours, 100 * (tot - ours) / tot), '',
         '| kernel | ours | launches | total us | us/launch | share |', '|---|---|---:|---:|---:|---:|']
for (mine, k), (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:40]:
    lines.append('| `%s` | %s | %d | %.1f | %.1f | %.1f%% |' % (k[:90], 'yes' if mine else '', c, t, t / c, 100 * t / tot))
open(sys.argv[2], 'w').write('\n'.join(lines) + '\n')
print('\n'.join(lines[:30]))
