"""Per-kernel SM-time of one step from an ncu launch list with sm__cycles_active.sum (the cost of a kernel when several batches
are in flight): python profiles/summarize_smtime.py launches.csv [out.md]

    ncu --metrics gpu__time_duration.sum,sm__cycles_active.sum --clock-control none --profile-from-start off --csv \
        --log-file launches.csv python profiles/glow_step.py glow32 1 throughput
"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[hi]
kn, mn, mv, idc = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv or not r[mv]:
        continue
    k = re.sub(r'^void ', '', re.sub(r'\(.*', '', r[kn])).replace('nfb::', '')
    a = agg.setdefault(k, {'n': set(), 'dur': 0.0, 'act': 0.0})
    v = float(r[mv].replace(',', ''))
    a['n'].add(r[idc])
    if r[mn] == 'gpu__time_duration.sum':
        a['dur'] += v
    elif r[mn] == 'sm__cycles_active.sum':
        a['act'] += v
tot = sum(a['act'] for a in agg.values())
lines = ['# SM-time per kernel of one step (ncu, cold-cache, serialised; 148 SMs at 1.965 GHz)', '',
         '%d launches, %.2f ms of machine time (sum of sm__cycles_active over all SMs / 148 / 1.965 GHz)' % (
             sum(len(a['n']) for a in agg.values()), tot / 148 / 1.965e6), '',
         '| kernel | launches | duration us (sum) | machine time ms | share |', '|---|---:|---:|---:|---:|']
for k, a in sorted(agg.items(), key=lambda x: -x[1]['act']):
    lines.append('| `%s` | %d | %.1f | %.3f | %.1f%% |' % (k, len(a['n']), a['dur'] / 1e3, a['act'] / 148 / 1.965e6, 100 * a['act'] / tot))
print('\n'.join(lines))
if len(sys.argv) > 2:
    open(sys.argv[2], 'w').write('\n'.join(lines) + '\n')
