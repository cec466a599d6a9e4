"""Streaming-size (inputs >> L2) launches of the HBM-bound layer kernels, for ncu captures and event timing."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import nfb200  # noqa: E402
import nfb200._lib as L  # noqa: E402
import bench  # noqa: E402

torch.set_grad_enabled(False)  # inference kernels; the gradient kernels below are called through the C ABI directly

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
    # checkerboard in place: z0 and z1 alternate pixel by pixel, so every 32-byte DRAM sector of z is read AND written back
    # whole: the physical floor is 4D (z) + 4D (t, s) + 4D (write-back) = 12 D although only 8 D are algorithmically needed
    ('affine checker 3x32x32 in-place (alg. 8D; sectors 12D)', affine(L.SPLIT_CHECKER, 3, 32, 32, True), 8 * D + 8),
    ('affine channel 12x16x16 in-place', affine(L.SPLIT_CHANNEL, 12, 16, 16, True), 8 * D + 8),
]
an = nfb200.flows.ActNorm(dims).cuda()
an.initialized = True
cases.append(('actnorm 3x32x32', lambda: an(z, ldj), 8 * D + 8))
lg = nfb200.flows.Logit(0.01)
zz = torch.rand((B, ) + dims, device='cuda')
cases.append(('logit 3x32x32', lambda: lg(zz, ldj), 8 * D + 8))
cases.append(('copy_ (torch) same bytes as 8D', lambda: out.copy_(z), 8 * D))
# mixture-CDF (Flow++) and RQ-spline couplings at streaming size: params are (2+3K) / (3K-1) values per transformed element
K = 8
Bm = 4096
zm = torch.randn((Bm, ) + dims, device='cuda')
outm = torch.empty_like(zm)
ldjm = torch.zeros(Bm, device='cuda')
pm = torch.randn(Bm, (2 + 3 * K) * (D // 2), device='cuda') * 0.5
cases.append(('mixlog fwd K=8 checker 3x32x32 (B=4096)', lambda: L.check(L.lib().nfb_mixlog_coupling_fwd(
    zm.data_ptr(), outm.data_ptr(), pm.data_ptr(), ldjm.data_ptr(), ldjm.data_ptr(), a.data_ptr(), b.data_ptr(), Bm, 3, 32, 32,
    L.SPLIT_CHECKER, 0, K, st)), (12 + 6 * K) * D + 8, Bm))
pr = torch.randn(Bm, (3 * K - 1) * (D // 2), device='cuda')
cases.append(('rqs fwd K=8 checker 3x32x32 (B=4096)', lambda: L.check(L.lib().nfb_rqs_coupling_fwd(
    zm.data_ptr(), outm.data_ptr(), pr.data_ptr(), ldjm.data_ptr(), ldjm.data_ptr(), Bm, 3, 32, 32, L.SPLIT_CHECKER, 0, K, 3.0,
    st)), (8 + 2 * (3 * K - 1)) * D + 8, Bm))
z1d = torch.randn(1 << 20, 64, device='cuda')
o1d = torch.empty_like(z1d)
l1d = torch.zeros(1 << 20, device='cuda')
p1d = torch.randn(1 << 20, 23 * 32, device='cuda')
cases.append(('rqs fwd K=8 1-D D=64 (B=2^20)', lambda: L.check(L.lib().nfb_rqs_coupling_fwd(
    z1d.data_ptr(), o1d.data_ptr(), p1d.data_ptr(), l1d.data_ptr(), l1d.data_ptr(), 1 << 20, 64, 1, 1, L.SPLIT_1D, 0, K, 3.0,
    st)), (8 + 2 * 23) * 64 + 8, 1 << 20))
pa1 = torch.randn(1 << 20, 64, device='cuda')
cases.append(('affine 1-D D=64 (B=2^20)', lambda: L.check(L.lib().nfb_affine_coupling_fwd(
    z1d.data_ptr(), o1d.data_ptr(), pa1.data_ptr(), l1d.data_ptr(), l1d.data_ptr(), a.data_ptr(), b.data_ptr(), 1 << 20, 64, 1, 1,
    L.SPLIT_1D, 0, st)), 12 * 64 + 8, 1 << 20))
sq = torch.empty(B, 12, 16, 16, device='cuda')
cases.append(('squeeze2d 3x32x32', lambda: L.check(L.lib().nfb_squeeze2d(z.data_ptr(), sq.data_ptr(), B, 3, 32, 32, 0, st)), 8 * D, B))
# invertible 1x1 convolution, forward direction, alone and with the preceding ActNorm fused (8 D bytes per sample either way)
for Cc, Hh in ((3, 32), (12, 16), (48, 8), (192, 4)):
    zc = z.view(B, Cc, Hh, Hh)
    oc = out.view(B, Cc, Hh, Hh)
    Wc = torch.linalg.qr(torch.randn(Cc, Cc, device='cuda'))[0].contiguous()
    lsc = torch.zeros(Cc, device='cuda')
    cases.append(('invconv1x1 apply C=%d %dx%d' % (Cc, Hh, Hh), lambda zc=zc, oc=oc, Wc=Wc, lsc=lsc, Cc=Cc, Hh=Hh: L.check(L.lib().nfb_invconv1x1_apply(
        zc.data_ptr(), oc.data_ptr(), ldj.data_ptr(), ldj.data_ptr(), Wc.data_ptr(), lsc.data_ptr(), 1.0, B, Cc, Hh * Hh, st)), 8 * D + 8))
    cases.append(('actnorm+invconv fused C=%d %dx%d' % (Cc, Hh, Hh), lambda zc=zc, oc=oc, Wc=Wc, lsc=lsc, Cc=Cc, Hh=Hh: L.check(L.lib().nfb_actnorm_invconv_fwd(
        zc.data_ptr(), oc.data_ptr(), ldj.data_ptr(), ldj.data_ptr(), lsc.data_ptr(), lsc.data_ptr(), Wc.data_ptr(), lsc.data_ptr(), B, Cc, Hh * Hh, st)), 8 * D + 8))
nr = torch.empty(B, device='cuda')
tot = torch.empty(2, device='cuda', dtype=torch.float64)
cases.append(('gauss_nll 3x32x32', lambda: L.check(L.lib().nfb_gauss_nll(z.data_ptr(), ldj.data_ptr(), nr.data_ptr(), tot.data_ptr(), B, D, st)), 4 * D + 8, B))
# ---- gradient kernels (csrc/backward.cu): bytes = tensors read + written once ---------------------------------------
gy = torch.randn((B, ) + dims, device='cuda')
gz = torch.empty_like(z)
gp = torch.empty_like(params)
gl = torch.randn(B, device='cuda')
gs = torch.empty(2, device='cuda')
scr = torch.empty(2 * 3072, device='cuda', dtype=torch.float64)


def affine_bwd(mode, C, H, W):
    return lambda: L.check(L.lib().nfb_affine_coupling_bwd(z.data_ptr(), params.data_ptr(), gy.data_ptr(), gl.data_ptr(),
                                                         gz.data_ptr(), gp.data_ptr(), gs.data_ptr(), gs.data_ptr() + 4,
                                                         scr.data_ptr(), a.data_ptr(), b.data_ptr(), B, C, H, W, mode, 0, st))


# reads z (4D) + s_raw (2D) + gy (4D), writes gz (4D) + gparams (4D) = 18 D
cases.append(('affine BWD checker 3x32x32', affine_bwd(L.SPLIT_CHECKER, 3, 32, 32), 18 * D + 4))
cases.append(('affine BWD channel 12x16x16', affine_bwd(L.SPLIT_CHANNEL, 12, 16, 16), 18 * D + 4))
cases.append(('affine BWD 1d 3072', affine_bwd(L.SPLIT_1D, 3072, 1, 1), 18 * D + 4))
ls3 = torch.zeros(3, device='cuda')
g3 = torch.empty(6, device='cuda')
cases.append(('actnorm BWD 3x32x32', lambda: L.check(L.lib().nfb_actnorm_bwd(
    gy.data_ptr(), z.data_ptr(), gl.data_ptr(), ls3.data_ptr(), ls3.data_ptr(), gz.data_ptr(), g3.data_ptr(), g3.data_ptr() + 12,
    scr.data_ptr(), B, 3, 1024, st)), 12 * D + 4))
z48 = z.view(B, 48, 8, 8)
W48 = torch.randn(48, 48, device='cuda') / 7
gW = torch.empty(48, 48, device='cuda')
wscr = torch.empty(int(L.lib().nfb_invconv1x1_wgrad_scratch(B, 48, 64)), device='cuda')
cases.append(('invconv wgrad C=48 8x8', lambda: L.check(L.lib().nfb_invconv1x1_wgrad(
    gy.data_ptr(), z48.data_ptr(), gW.data_ptr(), wscr.data_ptr(), B, 48, 64, st)), 8 * D))
cases.append(('invconv apply (W^T gy) C=48 8x8', lambda: L.check(L.lib().nfb_invconv1x1_apply(
    gy.data_ptr(), gz.data_ptr(), None, None, W48.data_ptr(), None, 0.0, B, 48, 64, st)), 8 * D))
gym = torch.randn_like(zm)
gzm = torch.empty_like(zm)
glm = torch.randn(Bm, device='cuda')
gpm = torch.empty_like(pm)
cases.append(('mixlog BWD K=8 checker 3x32x32 (B=4096)', lambda: L.check(L.lib().nfb_mixlog_coupling_bwd(
    zm.data_ptr(), pm.data_ptr(), gym.data_ptr(), glm.data_ptr(), gzm.data_ptr(), gpm.data_ptr(), gs.data_ptr(),
    gs.data_ptr() + 4, scr.data_ptr(), a.data_ptr(), b.data_ptr(), Bm, 3, 32, 32, L.SPLIT_CHECKER, 0, K, st)),
    (12 + 2 * 2 * (2 + 3 * K)) * D + 4, Bm))  # z, gy, gz (12 D) + params and gparams ((2+3K) * D/2 * 4 each)
gpr = torch.empty_like(pr)
cases.append(('rqs BWD K=8 checker 3x32x32 (B=4096)', lambda: L.check(L.lib().nfb_rqs_coupling_bwd(
    zm.data_ptr(), pr.data_ptr(), gym.data_ptr(), glm.data_ptr(), gzm.data_ptr(), gpr.data_ptr(), Bm, 3, 32, 32,
    L.SPLIT_CHECKER, 0, K, 3.0, st)), (12 + 2 * 2 * (3 * K - 1)) * D + 4, Bm))
md = None
if '--md' in sys.argv:
    md = open(sys.argv[sys.argv.index('--md') + 1], 'w')
    md.write('| kernel | µs | GB/s (algorithmic bytes) | of measured HBM peak (%.1f GB/s) |\n|---|---:|---:|---:|\n' % peaks['hbm_gbs'])
for case in cases:
    name, fn, bytes_per_sample = case[:3]
    nb = case[3] if len(case) > 3 else B
    for _ in range(3):
        fn()
    mean, med, best = bench.time_kernel_stream(fn, 15, flush)
    gbs = bytes_per_sample * nb / (med * 1e-3) / 1e9
    print('%-58s %8.1f us  %7.1f GB/s  %.3f of measured HBM peak' % (name, med * 1e3, gbs, gbs / peaks['hbm_gbs']))
    if md:
        md.write('| %s | %.1f | %.0f | %.3f |\n' % (name, med * 1e3, gbs, gbs / peaks['hbm_gbs']))
if md:
    md.close()
