"""Diagnostics of the tensor-core conditioner (csrc/conditioner_tc.cu) on a B200: error against the fp64 CPU oracle and
the FP32-FFMA kernel, fused conditioner+coupling against the two-kernel path, and CUDA-event timings per shape.
(The "knobs" lines need a library built with `make EXTRA=-DNFB_TC_DEBUG_KNOBS`; otherwise every knob column is the full kernel.)

    python profiles/tc2_probe.py [--quick]
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nfb200  # noqa: E402
import nfb200._lib as L  # noqa: E402
from oracle import flow_oracle as O  # noqa: E402

DEV = 'cuda:0'


def perturb_(m, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n_, p in list(m.named_parameters()) + list(m.named_buffers()):
            if not p.is_floating_point():
                continue
            if n_.endswith('running_var'):
                p.copy_(0.5 + torch.rand(p.shape, generator=g))
            elif n_.endswith('weight_g') or (n_.endswith('weight') and p.dim() == 1):
                p.copy_(0.5 + torch.rand(p.shape, generator=g))
            else:
                p.add_(0.1 * torch.randn(p.shape, generator=g))


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3  # us


def convnet_case(cin, cout, hw, B, groups=(3, 1)):
    F = nfb200.flows
    torch.manual_seed(cin + hw)
    net = F.ConvNet(cin, cout)
    perturb_(net, 9)
    net.eval()
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    x = torch.randn(B, cin, hw, hw)
    nref = min(B, 8)
    with torch.no_grad():
        ref64 = O.resnet_conditioner(O.to_dtype(sd, torch.float64), '', x[:nref].double())
        ref32 = O.resnet_conditioner(sd, '', x[:nref])
    net.to(DEV)
    for p_ in net.parameters():
        p_.requires_grad_(False)
    xd = x.to(DEV)
    net.kernel_flags = L.CONV_FFMA
    ffma = net(xd)
    t_ff = timeit(lambda: net(xd))
    net.kernel_flags = 0
    scale = float(ref64.abs().max())
    e_ff = float((ffma[:nref].cpu().double() - ref64).abs().max())
    e_32 = float((ref32.double() - ref64).abs().max())
    line = 'convnet cin=%3d cout=%3d %2dx%-2d B=%4d | scale %.2f | cpu32 %.2e ffma %.2e (%.1f us)' % (
        cin, cout, hw, hw, B, scale, e_32, e_ff, t_ff)
    for name, base_flags, gs in (('tf32', L.CONV_TF32, groups), ('f16', 0, (2, 1))):
        for G in gs:
            net.kernel_flags = base_flags | L.conv_groups(G)
            tc = net(xd)
            torch.cuda.synchronize()
            e_tc = float((tc[:nref].cpu().double() - ref64).abs().max())
            d_all = float((tc - ffma).abs().max())
            t_tc = timeit(lambda: net(xd))
            line += ' | %s G=%d %.2e, vs ffma(all) %.2e (%.1f us)' % (name, G, e_tc, d_all, t_tc)
    net.kernel_flags = 0
    print(line, flush=True)
    if B == 256:
        line = '    knobs (us):'
        for name, base_flags, G in (('tf32', L.CONV_TF32, 3), ('f16', 0, 2)):
            for dbg in (0, 1, 2, 4, 16, 32, 1 | 2, 1 | 2 | 16, 1 | 2 | 16 | 32, 2 | 16):
                net.kernel_flags = base_flags | L.conv_groups(G) | L.conv_debug(dbg)
                line += ' %s/%d=%.1f' % (name, dbg, timeit(lambda: net(xd)))
        net.kernel_flags = 0
        print(line, flush=True)


def fused_case(dims, masking, odd, B):
    F = nfb200.flows
    torch.manual_seed(7)
    cpl = F.AffineCoupling(dims, masking=masking, odd=odd)
    perturb_(cpl, 3)
    cpl.to(DEV).eval()
    x = torch.randn((B, ) + dims, device=DEV)
    l0 = torch.randn(B, device=DEV)
    with torch.no_grad():
        cpl.fused_conditioner = True
        cpl.net.kernel_flags = L.CONV_TF32
        z1, l1 = cpl(x, l0.clone())
        t_f = timeit(lambda: cpl.forward_fused(x.clone(), l0.clone(), inplace=True))
        t_clone = timeit(lambda: (x.clone(), l0.clone()))
        cpl.fused_conditioner = False
        z2, l2 = cpl(x, l0.clone())          # tensor-core conditioner + coupling kernel
        t_2 = timeit(lambda: cpl(x, l0.clone()))
        cpl.net.kernel_flags = L.CONV_FFMA
        z3, l3 = cpl(x, l0.clone())          # FFMA conditioner + coupling kernel
        t_3 = timeit(lambda: cpl(x, l0.clone()))
        cpl.net.kernel_flags = 0
        cpl.fused_conditioner = True
        z4, l4 = cpl(x, l0.clone())          # FP16-split fused kernel
        t_4 = timeit(lambda: cpl.forward_fused(x.clone(), l0.clone(), inplace=True))
        cpl.fused_conditioner = False
        z5, l5 = cpl(x, l0.clone())          # FP16-split conditioner + coupling kernel
        cpl.net.kernel_flags = L.CONV_PAIR
        cpl.fused_conditioner = True
        t_6 = timeit(lambda: cpl.forward_fused(x.clone(), l0.clone(), inplace=True))
        cpl.net.kernel_flags = 0
    print('fused %s %s odd=%d B=%d | z: fused-vs-2k %.2e, fused-vs-ffma %.2e | ldj %.2e / %.2e | fused %.1f us (clone %.1f) '
          '2-kernel %.1f us, ffma 2-kernel %.1f us || f16: z vs ffma %.2e, fused-vs-2k %.2e, ldj %.2e, fused %.1f us, pair %.1f us' %
          (dims, masking, odd, B, float((z1 - z2).abs().max()), float((z1 - z3).abs().max()),
           float((l1 - l2).abs().max()), float((l1 - l3).abs().max()), t_f, t_clone, t_2, t_3,
           float((z4 - z3).abs().max()), float((z4 - z5).abs().max()), float((l4 - l3).abs().max()), t_4, t_6), flush=True)


if __name__ == '__main__':
    torch.set_grad_enabled(False)
    quick = '--quick' in sys.argv
    t0 = time.time()
    print(torch.cuda.get_device_name(0))
    cases = [(6, 12, 16, 256), (24, 48, 8, 256), (96, 192, 4, 256)]
    if not quick:
        cases += [(6, 12, 16, 1), (6, 12, 16, 149), (6, 12, 16, 296), (24, 48, 8, 3), (96, 192, 4, 5), (96, 192, 4, 1184),
                  (3, 7, 8, 9), (40, 100, 4, 17), (24, 48, 16, 64), (96, 192, 8, 64), (384, 768, 4, 64), (1, 2, 16, 4)]
    for c in cases:
        convnet_case(*c)
    fcases = [((3, 32, 32), 'checkerboard', False, 256), ((12, 16, 16), 'channelwise', False, 256),
              ((12, 16, 16), 'checkerboard', True, 256), ((48, 8, 8), 'channelwise', True, 256),
              ((48, 8, 8), 'checkerboard', False, 256)]
    if not quick:
        fcases += [((3, 32, 32), 'checkerboard', True, 3), ((12, 16, 16), 'channelwise', True, 5),
                   ((48, 8, 8), 'checkerboard', True, 13), ((192, 8, 8), 'checkerboard', False, 16)]
    for c in fcases:
        fused_case(*c)
    print('done in %.1f s' % (time.time() - t0))
