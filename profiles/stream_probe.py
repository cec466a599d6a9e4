"""Streaming-size (inputs >> L2) launches of the HBM-bound layer kernels, for ncu captures and event timing."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import nfb200  # noqa: E402
import nfb200._lib as L  # noqa: E402
import bench  # noqa: E402

peaks = bench.load_peaks()
B, dims = 16384, (3, 32, 32)
D = 3072
z = torch.randn((B, ) + dims, device='cuda')
params = torch.randn((B, ) + dims, device='cuda')
out = torch.empty_like(z)
ldj = torch.zeros(B, device='cuda')
a = torch.tensor([0.3], device='cuda')
b = torch.tensor([0.01], device='cuda')
flush = torch.empty(256 * 1024 * 1024 // 4, device='cuda')
st = L.stream()


def affine(mode, C, H, W, inplace=False):
    o = z if inplace else out
    return lambda: L.check(L.lib().nfb_affine_coupling_fwd(z.data_ptr(), o.data_ptr(), params.data_ptr(), ldj.data_ptr(),
                                                         ldj.data_ptr(), a.data_ptr(), b.data_ptr(), B, C, H, W, mode, 0, st))


cases = [
    ('affine checker 3x32x32 out-of-place', affine(L.SPLIT_CHECKER, 3, 32, 32), 12 * D + 8),
    ('affine channel 12x16x16 out-of-place', affine(L.SPLIT_CHANNEL, 12, 16, 16), 12 * D + 8),
    ('affine 1d 3072 out-of-place', affine(L.SPLIT_1D, 3072, 1, 1), 12 * D + 8),
    ('affine checker 3x32x32 in-place', affine(L.SPLIT_CHECKER, 3, 32, 32, True), 8 * D + 8),
]
an = nfb200.flows.ActNorm(dims).cuda()
an.initialized = True
cases.append(('actnorm 3x32x32', lambda: an(z, ldj), 8 * D + 8))
lg = nfb200.flows.Logit(0.01)
zz = torch.rand((B, ) + dims, device='cuda')
cases.append(('logit 3x32x32', lambda: lg(zz, ldj), 8 * D + 8))
cases.append(('copy_ (torch) same bytes as 8D', lambda: out.copy_(z), 8 * D))
for name, fn, bytes_per_sample in cases:
    for _ in range(3):
        fn()
    mean, med, best = bench.time_kernel_stream(fn, 15, flush)
    gbs = bytes_per_sample * B / (med * 1e-3) / 1e9
    print('%-42s %8.1f us  %7.1f GB/s  %.3f of measured HBM peak' % (name, med * 1e3, gbs, gbs / peaks['hbm_gbs']))
