"""Timeline of CTA 0 of the tensor-core conditioner (developer instrumentation; needs a library built with
`make -C normalizing-flows-pytorch_b200/csrc EXTRA=-DNFB_TC_TIMELINE` -- the stamps are compiled out otherwise):
python profiles/tc2_timeline.py CIN COUT HW B [dbg] [fused: checker|channel|none] [extra kernel flags, e.g. 0x10000 = 3xTF32]"""
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
fused = sys.argv[6] if len(sys.argv) > 6 and sys.argv[6] != 'none' else None
extra = int(sys.argv[7], 0) if len(sys.argv) > 7 else 0
if fused:
    dims = (cin // 2, 2 * hw, 2 * hw) if fused == 'checker' else (2 * cin, hw, hw)
    cpl = nfb200.flows.AffineCoupling(dims, masking='checkerboard' if fused == 'checker' else 'channelwise').to('cuda:0').eval()
    net = cpl.net
    x = torch.randn((B, ) + dims, device='cuda:0')
    l0 = torch.zeros(B, device='cuda:0')
    run = lambda: cpl.forward_fused(x, l0, inplace=True)
else:
    net = nfb200.flows.ConvNet(cin, cout).to('cuda:0').eval()
    x = torch.randn(B, cin, hw, hw, device='cuda:0')
    run = lambda: net(x)
for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    run()
e1.record()
torch.cuda.synchronize()
print('eager back-to-back launch period: %.1f us' % (e0.elapsed_time(e1) / 20 * 1e3))
buf = torch.zeros(4 * 512, dtype=torch.int64, device='cuda:0')
fn = L.lib().nfb_debug_timeline
fn.argtypes = [ctypes.c_void_p]
fn(buf.data_ptr())
net.kernel_flags = L.conv_debug(dbg) | extra
run()
torch.cuda.synchronize()
fn(None)
ev = buf.cpu().view(4, 512)
NAMES = {1: 'mma wait w_full', 2: 'mma got w_full', 10: 'mma wait act0', 11: 'mma wait act1', 12: 'mma got act0', 13: 'mma got act1',
         26: 'commit acc0', 27: 'commit acc1', 28: 'commit war0', 30: 'epi wait acc', 22: 'commit w_empty0', 23: 'commit w_empty1',
         31: 'epi got acc', 32: 'epi wait war', 33: 'epi got war', 34: 'epi signal..', 35: 'epi signalled',
         36: 'epi acc loaded', 0: 'kernel entry', 1: 'prologue done', 50: 'kernel exit', 40: 'epi unit start', 41: 'epi out layer', 42: 'epi done'}
rows = []
for role in range(4):
    for v in ev[role].tolist():
        if v == 0:
            continue
        v &= (1 << 64) - 1
        rows.append((v & 0xffffffffffff, role, v >> 48))
rows.sort()
t0 = rows[0][0]
last = {0: t0, 1: t0, 2: t0, 3: t0}
for t, role, tag in rows:
    print('%8d  (+%6d)  %s %s' % (t - t0, t - last[role], ['E0 ', ' E1', '  M', 'K  '][role], NAMES.get(tag, str(tag))))
    last[role] = t
