"""Per-kernel times of the train-mode conditioner kernels (csrc/conditioner_train.cu) at the BASELINE cfg-2 shapes, B = 256:
CUDA events around 20 back-to-back launches (L2-warm, as inside a training step)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import nfb200  # noqa: E402
import nfb200._lib as L  # noqa: E402
from nfb200.flows import conditioner_train as CT  # noqa: E402

torch.set_grad_enabled(False)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256


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
    return e0.elapsed_time(e1) / n * 1e3


for (cin, cout, hw) in ((6, 12, 16), (24, 48, 8), (96, 192, 4)):
    x = torch.randn(B, cin, hw, hw, device='cuda')
    h = torch.randn(B, 32, hw, hw, device='cuda')
    go = torch.randn(B, cout, hw, hw, device='cuda')

    def pack(O, I, KK):
        v = torch.randn(O, I, *((3, 3) if KK == 9 else (1, 1)), device='cuda') * 0.1
        g = torch.rand(I, *((3, 3) if KK == 9 else (1, 1)), device='cuda') + 0.5
        n = ((O + 31) // 32) * ((I + 31) // 32) * 32 * KK * 32
        wn, wf, wb = torch.empty_like(v), torch.empty(n, device='cuda'), torch.empty(n, device='cuda')
        f = lambda: L.check(L.lib().nfb_wn_pack_train(L.ptr(v), L.ptr(g), L.ptr(wn), L.ptr(wf), L.ptr(wb), O, I, KK, 1e-5, L.stream()))  # noqa: E731
        f()
        return v, g, wn, wf, wb, f

    v0, g0, _, wf0, wb0, f0 = pack(32, cin, 9)
    v1, g1, _, wf1, wb1, f1 = pack(32, 32, 9)
    v5, g5, _, wf5, wb5, f5 = pack(cout, 32, 1)
    bias = torch.zeros(32, device='cuda')
    gam, bet = torch.ones(32, device='cuda'), torch.zeros(32, device='cuda')
    rm, rv = torch.zeros(32, device='cuda'), torch.ones(32, device='cuda')
    _, st = CT._conv(h, wf1, bias, None, 32, 32, 3, True)
    a, mr = CT._bn_relu(h, st, gam, bet, rm, rv, 0.1, 1e-5)
    gw = torch.randn(32, 32, 3, 3, device='cuda')
    gv, gg = torch.empty_like(v1), torch.empty_like(g1)
    rows = [
        ('wn_pack 32x%dx9' % cin, f0), ('wn_pack 32x32x9', f1), ('wn_pack %dx32x1' % cout, f5),
        ('conv3 %d->32 (+stats)' % cin, lambda: CT._conv(x, wf0, bias, None, cin, 32, 3, True)),
        ('conv3 32->32 (+skip,+stats)', lambda: CT._conv(h, wf1, bias, h, 32, 32, 3, True)),
        ('conv1 32->%d' % cout, lambda: CT._conv(h, wf5, None, None, 32, cout, 1, False)),
        ('dgrad conv1 %d->32' % cout, lambda: CT._conv(go, wb5, None, None, cout, 32, 1, False)),
        ('dgrad conv3 32->32', lambda: CT._conv(h, wb1, None, None, 32, 32, 3, False)),
        ('dgrad conv3 32->%d' % cin, lambda: CT._conv(h, wb0, None, None, 32, cin, 3, False)),
        ('wgrad3 32x32', lambda: CT._wgrad(h, a, 32, 32, 3)),
        ('wgrad3 32x%d' % cin, lambda: CT._wgrad(h, x, cin, 32, 3)),
        ('wgrad1 %dx32' % cout, lambda: CT._wgrad(go, a, 32, cout, 1)),
        ('bn_relu fwd', lambda: CT._bn_relu(h, st, gam, bet, rm, rv, 0.1, 1e-5)),
        ('bn_relu bwd (reduce+apply)', lambda: CT._bn_relu_bwd(h, a, h, mr, gam, h)),
        ('wn_bwd 32x32x9', lambda: L.check(L.lib().nfb_wn_bwd(L.ptr(v1), L.ptr(g1), L.ptr(gw), L.ptr(gv), L.ptr(gg), 32, 288, 1e-5, L.stream()))),
    ]
    print('--- conditioner %d -> %d at %dx%d, B = %d (times include the torch.empty allocations of the wrappers)' % (cin, cout, hw, hw, B))
    for name, fn in rows:
        print('%-32s %8.1f us' % (name, timeit(fn)))
