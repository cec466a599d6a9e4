"""Per-kernel GPU time of ONE eager training step (train-mode forward, loss, backward, Adam) via torch.profiler (CUPTI):
    python profiles/train_step_profile.py glow32 > profiles/rNN_train_kernels_glow32.md
Kernel durations are the device-side ones (warm caches, clocks as they are), so shares and absolute times are comparable
with the CUDA-graph replay of the same step."""
import collections
import os
import re
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402
import nfb200  # noqa: E402
import bench  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else 'glow32'
model, dims, datatype, cfg, batch, desc = bench.WORKLOADS[wl]
torch.manual_seed(0)
net = getattr(nfb200, {'glow': 'Glow', 'flowpp': 'Flowpp', 'realnvp': 'RealNVP'}[model])(
    dims, datatype, types.SimpleNamespace(**cfg)).cuda().train()
x = bench.make_inputs(dims, datatype, batch, 0).cuda()
opt = torch.optim.Adam(net.parameters(), lr=1e-4)
with torch.no_grad():
    net(x)
for _ in range(2):
    nfb200.parallel.train_step(net, opt, x)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    nfb200.parallel.train_step(net, opt, x)
    torch.cuda.synchronize()
agg = collections.OrderedDict()
tot = 0.0
for ev in prof.events():
    if ev.device_type != torch.autograd.DeviceType.CUDA:
        continue
    t = getattr(ev, 'device_time_total', None)
    if t is None:
        t = ev.cuda_time_total
    name = ev.name
    mine = 'nfb::' in name
    k = re.sub(r'^void ', '', re.sub(r'\(.*', '', name)).replace('nfb::', '')
    if not mine:
        k = re.sub(r'<.*', '<...>', k)
    a = agg.setdefault((mine, k), [0, 0.0])
    a[0] += 1
    a[1] += t
    tot += t
ours = sum(v[1] for (m, _), v in agg.items() if m)
n_all = sum(v[0] for v in agg.values())
n_ours = sum(v[0] for (m, _), v in agg.items() if m)
print('# device time of one eager %s TRAINING step (torch.profiler / CUPTI, warm)\n' % wl)
print('%d device activities, %.1f ms total; libnfb200: %d launches, %.1f ms (%.1f%%); library / ATen / memset: %d, %.1f ms\n'
      % (n_all, tot / 1e3, n_ours, ours / 1e3, 100 * ours / tot, n_all - n_ours, (tot - ours) / 1e3))
print('| kernel | ours | launches | total ms | us/launch | share |\n|---|---|---:|---:|---:|---:|')
for (mine, k), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print('| `%s` | %s | %d | %.2f | %.1f | %.1f%% |' % (k[:100], 'yes' if mine else '', c, t / 1e3, t / c, 100 * t / tot))
