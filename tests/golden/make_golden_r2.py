"""Round-2 golden vectors from the UNMODIFIED reference (run in the build container; needs /root/reference):
AdditiveCoupling (coupling.py:52-79) and the standalone MixLogCDF module (modules.py:186-212).

    python tests/golden/make_golden_r2.py

Note: the reference's AdditiveCoupling builds ConvNet(dims[0], dims[0]) for the checkerboard masking (coupling.py:62-63)
although checker_split returns halves of 2*dims[0] channels -- it raises at the first forward, so only the channelwise
and 1-D maskings have a reference behaviour to pin.
"""
import os
import sys

import torch

sys.dont_write_bytecode = True
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import gen_layer_case, perturb, rc, rm, save  # noqa: E402,F401


def main():
    gen = torch.Generator().manual_seed(4321)
    n = lambda *s: torch.randn(*s, generator=gen)  # noqa: E731
    gen_layer_case(gen, 'additive_channel_0', rc.AdditiveCoupling, dict(dims=(12, 4, 4), masking='channelwise', odd=False), n(3, 12, 4, 4))
    gen_layer_case(gen, 'additive_channel_1', rc.AdditiveCoupling, dict(dims=(12, 8, 8), masking='channelwise', odd=True), n(2, 12, 8, 8))
    gen_layer_case(gen, 'additive_1d_0', rc.AdditiveCoupling, dict(dims=(16, ), odd=False), n(6, 16))
    gen_layer_case(gen, 'additive_1d_1', rc.AdditiveCoupling, dict(dims=(16, ), odd=True), n(6, 16))
    try:
        torch.manual_seed(7)
        rc.AdditiveCoupling(dims=(3, 8, 8), masking='checkerboard')(n(2, 3, 8, 8), torch.zeros(2))
        print('NOTE: the reference accepted the checkerboard AdditiveCoupling')
    except Exception as e:
        print('reference AdditiveCoupling(checkerboard) raises:', type(e).__name__)

    # MixLogCDF: x (B, *C), log_pi / mu / s (B, K, *C); log_pi log-softmaxed over K like its only call site (coupling.py:180)
    for name, shape, K in (('mixlogcdf_img', (3, 6, 4, 4), 4), ('mixlogcdf_1d', (5, 7), 8)):
        B = shape[0]
        x = n(*shape)
        log_pi = torch.log_softmax(n(B, K, *shape[1:]), dim=1)
        mu = n(B, K, *shape[1:])
        s = 0.5 * n(B, K, *shape[1:])
        ldj0 = n(B)
        m = rm.MixLogCDF()
        with torch.no_grad():
            y, l1 = m(x.clone(), log_pi, mu, s, ldj0.clone())
            xi, l2 = m.backward(y.clone(), log_pi, mu, s, ldj0.clone())
        save(name, dict(kind='MixLogCDF', K=K),
             dict(x=x.numpy(), log_pi=log_pi.numpy(), mu=mu.numpy(), s=s.numpy(), ldj0=ldj0.numpy(), fwd_y=y.numpy(),
                  fwd_ldj=l1.numpy(), inv_x=xi.numpy(), inv_ldj=l2.numpy()))


if __name__ == '__main__':
    main()
