"""Warm per-layer device time of one workload (CUDA-graph replay of 20 back-to-back calls per distinct layer shape) next to
the single-stream step time: where does a step go?   python profiles/layer_probe.py [workload]"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import nfb200  # noqa: E402
from nfb200.flows import modules as M  # noqa: E402
from nfb200.flows.coupling import AffineCoupling  # noqa: E402

torch.set_grad_enabled(False)
wl = sys.argv[1] if len(sys.argv) > 1 else 'glow32'
W = bench.WORKLOADS[wl]
torch.manual_seed(0)
net = getattr(nfb200, bench.MODEL_CLASS[W['model']])(W['dims'], W['datatype'], types.SimpleNamespace(**W['cfg'])).cuda().eval()
x = bench.make_inputs(W['dims'], W['datatype'], W['batch'], 0).cuda()
net(x)
torch.cuda.synchronize()

# walk the stack once, recording (kind, input shape) of every peephole unit and one representative callable per kind
layers = list(net.net.layers)
seen, order = {}, []
z, ldj = x, torch.zeros(x.size(0), device='cuda')
i = 0
while i < len(layers):
    l = layers[i]
    if type(l) is M.ActNorm and i + 1 < len(layers) and type(layers[i + 1]) is M.InvertibleConv1x1:
        key = ('actnorm+invconv', tuple(z.shape[1:]))
        fn = (lambda a=l, c=layers[i + 1], zz=z.clone(), ll=ldj.clone(): M._actnorm_invconv(a, c, zz, ll))
        z, ldj = M._actnorm_invconv(l, layers[i + 1], z, ldj)
        i += 2
    elif isinstance(l, AffineCoupling):
        key = ('conditioner+coupling', tuple(z.shape[1:]), l.mode)
        fn = (lambda c=l, zz=z.clone(), ll=ldj.clone(): c.forward_fused(zz, ll, inplace=True))
        z, ldj = l(z, ldj)
        i += 1
    else:
        key = (type(l).__name__, tuple(z.shape[1:]))
        fn = (lambda c=l, zz=z.clone(), ll=ldj.clone(): c(zz, ll))
        z, ldj = l(z, ldj)
        i += 1
    if key not in seen:
        seen[key] = [0, fn]
        order.append(key)
    seen[key][0] += 1
tot = 0.0
print('%-70s %6s %9s %9s' % ('layer', 'count', 'us/call', 'us/step'))
for key in order:
    n, fn = seen[key]
    us = bench.graph_time_us(fn)
    tot += n * us
    print('%-70s %6d %9.1f %9.1f' % (str(key), n, us, n * us))
print('sum of layers: %.2f ms' % (tot / 1e3))
H = bench.Harness(wl, W['batch'], 0, 1, torch.device('cuda', 0), 1)
ms = H.timed(10, 3, 1, False, False) / 10
print('single-stream step (graph replay): %.2f ms' % ms)
