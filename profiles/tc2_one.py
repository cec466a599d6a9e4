"""One shape of the tensor-core conditioner, a few launches (for ncu): python profiles/tc2_one.py CIN COUT HW B [fused]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nfb200  # noqa: E402

torch.set_grad_enabled(False)
cin, cout, hw, B = (int(a) for a in sys.argv[1:5])
fused = len(sys.argv) > 5
if fused:
    C = cin * 2  # channelwise coupling on (C, hw, hw): conditioner cin -> 2 cin
    cpl = nfb200.flows.AffineCoupling((C, hw, hw), masking='channelwise').to('cuda:0').eval()
    x = torch.randn(B, C, hw, hw, device='cuda:0')
    l0 = torch.zeros(B, device='cuda:0')
    for _ in range(5):
        cpl.forward_fused(x, l0, inplace=True)
else:
    net = nfb200.flows.ConvNet(cin, cout).to('cuda:0').eval()
    x = torch.randn(B, cin, hw, hw, device='cuda:0')
    for _ in range(5):
        net(x)
torch.cuda.synchronize()
print('ok')
