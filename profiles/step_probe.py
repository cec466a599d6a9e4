"""Per-shape device time of the fused flow-step kernel with and without the next step's ActNorm + 1x1 conv, next to the
separate ActNorm+conv kernel (CUDA-graph replay of 20 calls).   python profiles/step_probe.py [B]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import nfb200  # noqa: E402
import nfb200._lib as L  # noqa: E402
from nfb200.flows import modules as M  # noqa: E402

torch.set_grad_enabled(False)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for dims, masking in (((3, 32, 32), 'checkerboard'), ((12, 16, 16), 'channelwise'), ((12, 16, 16), 'checkerboard'),
                      ((48, 8, 8), 'channelwise'), ((48, 8, 8), 'checkerboard')):
    torch.manual_seed(0)
    F = nfb200.flows
    cpl = F.AffineCoupling(dims, masking=masking).cuda().eval()
    an = F.ActNorm(dims).cuda()
    an.initialized = True
    conv = F.InvertibleConv1x1(dims[0]).cuda()
    z = torch.randn((B, ) + dims, device='cuda')
    ldj = torch.zeros(B, device='cuda')
    for pair in (False, True):
        cpl.net.kernel_flags = L.CONV_PAIR if pair else 0
        t0 = bench.graph_time_us(lambda: cpl.forward_fused(z, ldj, inplace=True))
        t1 = bench.graph_time_us(lambda: cpl.forward_fused(z, ldj, inplace=True, post=(an, conv)))
        t2 = bench.graph_time_us(lambda: M._actnorm_invconv(an, conv, z, ldj))
        print('%-28s pair=%d  coupling %.1f us | +next ActNorm/conv in-kernel %.1f us | separate ActNorm+conv kernel %.1f us' %
              (str(dims) + ' ' + masking[:5], pair, t0, t1, t2), flush=True)
