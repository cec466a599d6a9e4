"""Timeline of CTA 0 of the tensor-core conditioner (developer instrumentation): python profiles/tc2_timeline.py CIN COUT HW B"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nfb200  # noqa: E402
import nfb200._lib as L  # noqa: E402

torch.set_grad_enabled(False)
cin, cout, hw, B = (int(a) for a in sys.argv[1:5])
dbg = int(sys.argv[5]) if len(sys.argv) > 5 else 0
net = nfb200.flows.ConvNet(cin, cout).to('cuda:0').eval()
x = torch.randn(B, cin, hw, hw, device='cuda:0')
for _ in range(3):
    net(x)
buf = torch.zeros(3 * 512, dtype=torch.int64, device='cuda:0')
fn = L.lib().nfb_debug_timeline
fn.argtypes = [ctypes.c_void_p]
fn(buf.data_ptr())
net.kernel_flags = L.conv_debug(dbg)
net(x)
torch.cuda.synchronize()
fn(None)
ev = buf.cpu().view(3, 512)
NAMES = {1: 'mma wait w_full', 2: 'mma got w_full', 10: 'mma wait act0', 11: 'mma wait act1', 12: 'mma got act0', 13: 'mma got act1',
         26: 'commit acc0', 27: 'commit acc1', 28: 'commit war0', 22: 'commit w_empty0', 23: 'commit w_empty1',
         30: 'epi wait acc', 31: 'epi got acc', 32: 'epi wait war', 33: 'epi got war', 34: 'epi signal..', 35: 'epi signalled',
         36: 'epi acc loaded', 40: 'epi unit start', 41: 'epi out layer', 42: 'epi done'}
rows = []
for role in range(3):
    for v in ev[role].tolist():
        if v == 0:
            continue
        v &= (1 << 64) - 1
        rows.append((v & 0xffffffffffff, role, v >> 48))
rows.sort()
t0 = rows[0][0]
last = {0: t0, 1: t0, 2: t0}
for t, role, tag in rows:
    print('%8d  (+%6d)  %s %s' % (t - t0, t - last[role], ['E0 ', ' E1', '  M'][role], NAMES.get(tag, str(tag))))
    last[role] = t
