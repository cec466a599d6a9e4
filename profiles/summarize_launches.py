"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) of profiles/glow_step.py into per-kernel
shares of ONE step (the last step in the capture).  Usage: summarize_launches.py launches.csv out.md"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr, data = rows[hi], rows[hi + 1:]
kn, mv, mu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
step = [r for r in data if len(r) > mv and r[mv]]  # the profiled range (cudaProfilerStart/Stop) = the step(s)
agg, tot = collections.OrderedDict(), 0.0
for r in step:
    v = float(r[mv].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}[r[mu]]
    k = re.sub(r'^void ', '', re.sub(r'\(.*', '', r[kn])).replace('nfb::', '')
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
lines = ['# ncu launch list of one step (eager, cold-cache, serialised: compare SHARES)', '',
         '`ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv python profiles/glow_step.py <workload> 1`', '',
         '%d launches, %.1f us total' % (len(step), tot), '', '| kernel | launches | total us | us/launch | share |', '|---|---:|---:|---:|---:|']
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    lines.append('| `%s` | %d | %.1f | %.1f | %.1f%% |' % (k, c, t, t / c, 100 * t / tot))
open(sys.argv[2], 'w').write('\n'.join(lines) + '\n')
print('\n'.join(lines))
