"""Where does the tensor-core conditioner spend its time?  Debug knobs (nfb_set_tuning key 4): 1 = no MMAs,
2 = one TMEM load per tile instead of four, 4 = single-pass TF32 (1 MMA per k-step instead of 3)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import nfb200  # noqa: E402
import nfb200._lib as L  # noqa: E402
from bench import graph_time_us  # noqa: E402

torch.manual_seed(0)
for cin, cout, hw in [(6, 12, 16), (24, 48, 8), (96, 192, 4)]:
    net = nfb200.flows.ConvNet(cin, cout).cuda().eval()
    for B in (148, 256):
        x = torch.randn(B, cin, hw, hw, device='cuda')
        L.check(L.lib().nfb_set_tuning(3, 1))
        res = []
        for dbg in (0, 1, 2, 3, 4):
            L.check(L.lib().nfb_set_tuning(4, dbg))
            res.append(graph_time_us(lambda: net(x)))
        L.lib().nfb_set_tuning(4, 0)
        L.lib().nfb_set_tuning(3, 0)
        ff = graph_time_us(lambda: net(x))
        print('%2dx%-2d B=%d: full %.1f | no-MMA %.1f | 1 TMEM ld %.1f | no-MMA+1ld %.1f | 1xTF32 %.1f | FFMA %.1f us' %
              (hw, hw, B, res[0], res[1], res[2], res[3], res[4], ff))
