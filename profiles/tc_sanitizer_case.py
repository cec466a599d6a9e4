"""Small launches of every variant of the tensor-core conditioner (+ coupling) kernel, for compute-sanitizer:
    compute-sanitizer --tool memcheck python profiles/tc_sanitizer_case.py
one / two units per CTA, FP16-split / 3xTF32 operands, ragged batches (invalid trailing units), one-launch flow steps."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nfb200  # noqa: E402
import nfb200._lib as L  # noqa: E402

torch.set_grad_enabled(False)
F = nfb200.flows
n = 0
for dims, masking in (((3, 32, 32), 'checkerboard'), ((12, 16, 16), 'channelwise'), ((12, 16, 16), 'checkerboard'),
                      ((48, 8, 8), 'channelwise'), ((48, 8, 8), 'checkerboard')):
    for B in (1, 5):
        layers = [F.ActNorm(dims), F.InvertibleConv1x1(dims[0]), F.AffineCoupling(dims, masking=masking),
                  F.ActNorm(dims), F.InvertibleConv1x1(dims[0])]
        comp = F.Compose(layers).cuda().eval()
        for m in comp.modules():
            if isinstance(m, F.ActNorm):
                m.initialized = True
        x = torch.randn((B, ) + dims, device='cuda')
        l0 = torch.zeros(B, device='cuda')
        ref = None
        for flags in (0, L.CONV_PAIR, L.CONV_SINGLE, L.CONV_TF32, L.CONV_TF32 | L.CONV_PAIR):
            for fuse in (1, 2):
                for m in comp.modules():
                    if isinstance(m, F.ConvNet):
                        m.kernel_flags = flags
                comp.fuse_steps = fuse
                z, l = comp(x, l0.clone())
                torch.cuda.synchronize()
                if ref is None:
                    ref = z
                assert torch.isfinite(z).all() and float((z - ref).abs().max()) < 1e-3 * max(1.0, float(ref.abs().max()))
                n += 1
print('ok: %d configurations' % n)
