"""A/B timing of the fused-ConvNet kernel variants (nfb_set_tuning) at the BASELINE cfg-2 shapes, CUDA events."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import nfb200  # noqa: E402
import nfb200._lib as L  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256


def timeit(fn, iters=15, per_graph=20):
    """median device time of ONE fn() in us: `per_graph` back-to-back launches captured in a CUDA graph (no Python
    launch overhead between them), CUDA events around each replay."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(per_graph):
            fn()
    g.replay()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record()
        g.replay()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2] * 1e3 / per_graph


torch.manual_seed(0)
for key, values, cin, cout, hw in [(0, (0, 1), 6, 12, 16), (1, (0, 1, 2, 3), 24, 48, 8), (2, (0, 1, 2, 3, 4), 96, 192, 4)]:
    net = nfb200.flows.ConvNet(cin, cout).cuda().eval()
    x = torch.randn(B, cin, hw, hw, device='cuda')
    mac = hw * hw * (cin * 288 + 4 * 9216 + 32 * cout)
    for v in values:
        L.check(L.lib().nfb_set_tuning(key, v))
        us = timeit(lambda: net(x))
        print('convnet %dx%d cin=%d cout=%d B=%d variant %d: %8.1f us  %6.2f TFLOP/s' %
              (hw, hw, cin, cout, B, v, us, 2 * mac * B / us * 1e-6))
    L.lib().nfb_set_tuning(key, 0)
    L.check(L.lib().nfb_set_tuning(3, 1))
    us = timeit(lambda: net(x))
    print('convnet %dx%d cin=%d cout=%d B=%d TENSOR-CORE (3xTF32): %8.1f us  %6.2f TFLOP/s (logical)' %
          (hw, hw, cin, cout, B, us, 2 * mac * B / us * 1e-6))
    L.lib().nfb_set_tuning(3, 0)

# 1x1 conv (+ fused ActNorm) and the elementwise layers at the cfg-2 shapes (L2-resident working set)
F = nfb200.flows
for dims in [(3, 32, 32), (12, 16, 16), (48, 8, 8)]:
    an = F.ActNorm(dims).cuda()
    an.initialized = True
    conv = F.InvertibleConv1x1(dims[0]).cuda()
    comp = F.Compose([an, conv]).cuda()
    z = torch.randn((B, ) + dims, device='cuda')
    ldj = torch.zeros(B, device='cuda')
    conv.matrices()
    t_f = timeit(lambda: comp(z, ldj))
    comp.fuse_steps = False
    t_u = timeit(lambda: comp(z, ldj))
    t_c = timeit(lambda: conv(z, ldj))
    t_a = timeit(lambda: an(z, ldj))
    print('dims %s: actnorm+invconv fused %.1f us, separate %.1f us (invconv %.1f, actnorm %.1f)' % (dims, t_f, t_u, t_c, t_a))
    for masking in (['checkerboard', 'channelwise'] if dims[0] % 2 == 0 else ['checkerboard']):
        cpl = F.AffineCoupling(dims, masking=masking).cuda().eval()
        params = cpl._params(z)
        out = torch.empty_like(z)
        C, H, W = dims
        fn = lambda: L.check(L.lib().nfb_affine_coupling_fwd(z.data_ptr(), out.data_ptr(), params.data_ptr(), ldj.data_ptr(),
                                                           ldj.data_ptr(), cpl.s_log_scale.data_ptr(), cpl.s_bias.data_ptr(),
                                                           B, C, H, W, cpl.mode, 0, L.stream()))
        print('   affine %s: %.1f us; full coupling layer %.1f us' % (masking, timeit(fn), timeit(lambda: cpl(z, ldj))))
