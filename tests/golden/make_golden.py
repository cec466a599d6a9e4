"""Generate golden input/output vectors from the UNMODIFIED reference (run in the build container).

    python tests/golden/make_golden.py            # needs /root/reference (read-only mount)

The reference (tatsy/normalizing-flows-pytorch) ships no tests or fixtures, so the
oracle is pinned against what the reference itself computes here on CPU (torch 2.11.0, fp32).
Every case stores: the layer/model constructor arguments, the (perturbed, so that BatchNorm / ActNorm /
1x1-conv are not near-identity) state dict, seeded inputs, and the reference's forward and
``backward`` (inverse) outputs.  Files are small ``.npz`` archives committed under ``tests/golden/``.

Nothing at test/bench time imports the reference; only this script does.
"""
import json
import os
import sys
import types

import numpy as np
import torch

sys.dont_write_bytecode = True
REF = os.environ.get('NFB_REFERENCE', '/root/reference')
sys.path.insert(0, REF)
import warnings  # noqa: E402

warnings.filterwarnings('ignore')
import flows  # noqa: E402  (the reference package)
from flows import coupling as rc, modules as rm, squeeze as rs  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(4)


def perturb(module, gen):
    """Move every parameter/buffer away from its near-identity initial value (seeded)."""
    with torch.no_grad():
        for name, t in list(module.named_parameters()) + list(module.named_buffers()):
            leaf = name.split('.')[-1]
            if not t.is_floating_point() or leaf in ('P', 'I', 'L_mask', 'U_mask', 'sign_s', 'pivots'):
                continue
            noise = torch.randn(t.shape, generator=gen)
            if leaf in ('running_var', 'batch_var'):
                t.copy_(0.5 + torch.rand(t.shape, generator=gen))
            elif leaf in ('running_mean', 'batch_mean', 'beta', 'bias', 'log_gamma', 'log_scale'):
                t.add_(0.1 * noise)
            elif leaf in ('L', 'U', 'log_s'):
                t.add_(0.05 * noise)
            elif leaf in ('s_log_scale', 'a_log_scale'):
                t.copy_(0.3 + 0.1 * noise)
            elif leaf in ('s_bias', 'a_bias'):
                t.copy_(0.05 * noise)
            elif leaf == 'weight' and t.dim() == 1:  # BatchNorm / LayerNorm gains
                t.add_(0.1 * noise)
            else:
                t.add_(0.02 * noise)


def sd_np(module, prefix='sd/'):
    return {prefix + k: v.detach().cpu().numpy() for k, v in module.state_dict().items()}


def save(name, meta, arrays):
    arrays = dict(arrays)
    arrays['meta'] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **arrays)
    print('%-40s %8.1f KB' % (name, os.path.getsize(path) / 1024))


def run_layer(layer, x, ldj0):
    layer.eval()
    with torch.no_grad():
        z, l1 = layer(x.clone(), ldj0.clone())
        y, l2 = layer.backward(z.clone(), ldj0.clone())
    return dict(x=x.numpy(), ldj0=ldj0.numpy(), fwd_z=z.numpy(), fwd_ldj=l1.numpy(), inv_y=y.numpy(),
                inv_ldj=l2.numpy())


def gen_layer_case(gen, name, ctor, ctor_kwargs, x, pre=None):
    """One layer: seeded construction, perturbed parameters, forward + backward through the reference, saved as <name>.npz."""
    torch.manual_seed(7)
    layer = ctor(**ctor_kwargs)
    perturb(layer, gen)
    if pre is not None:
        pre(layer)
    ldj0 = torch.randn(x.shape[0], generator=gen)
    arrays = run_layer(layer, x, ldj0)
    arrays.update(sd_np(layer))
    kw = {k: (list(v) if isinstance(v, tuple) else v) for k, v in ctor_kwargs.items()}
    save(name, dict(kind=ctor.__name__, kwargs=kw), arrays)


def main():
    gen = torch.Generator().manual_seed(1234)

    # ---- index permutations: arange tensors, bit-exact -------------------------------
    arrs = {}
    for (C, H, W) in [(1, 2, 2), (3, 4, 4), (2, 6, 4), (12, 8, 8)]:
        z = torch.arange(2 * C * H * W, dtype=torch.float32).view(2, C, H, W)
        tag = '%dx%dx%d' % (C, H, W)
        arrs['in_' + tag] = z.numpy()
        for odd in (False, True):
            z0, z1 = rs.checker_split(z, odd)
            arrs['checker_%s_%d_z0' % (tag, odd)] = z0.contiguous().numpy()
            arrs['checker_%s_%d_z1' % (tag, odd)] = z1.contiguous().numpy()
            assert torch.equal(rs.checker_merge(z0, z1, odd), z)
            if C % 2 == 0:
                c0, c1 = rs.channel_split(z, 1, odd)
                arrs['channel_%s_%d_z0' % (tag, odd)] = c0.contiguous().numpy()
                arrs['channel_%s_%d_z1' % (tag, odd)] = c1.contiguous().numpy()
            sq = rs.Squeeze2d(odd)(z, None)[0]
            arrs['squeeze2d_%s_%d' % (tag, odd)] = sq.contiguous().numpy()
            assert torch.equal(rs.Unsqueeze2d(odd)(sq, None)[0], z)
    v = torch.arange(3 * 10, dtype=torch.float32).view(3, 10)
    arrs['in_1d'] = v.numpy()
    for odd in (False, True):
        a, b = rs.squeeze1d(v, odd)
        arrs['split1d_%d_z0' % odd] = a.contiguous().numpy()
        arrs['split1d_%d_z1' % odd] = b.contiguous().numpy()
        assert torch.equal(rs.unsqueeze1d(a, b, odd), v)
    save('permutations', dict(kind='permutations'), arrs)

    # ---- single layers ------------------------------------------------------------------
    def layer_case(name, ctor, ctor_kwargs, x, pre=None):
        torch.manual_seed(7)
        layer = ctor(**ctor_kwargs)
        perturb(layer, gen)
        if pre is not None:
            pre(layer)
        ldj0 = torch.randn(x.shape[0], generator=gen)
        arrays = run_layer(layer, x, ldj0)
        arrays.update(sd_np(layer))
        kw = {k: (list(v) if isinstance(v, tuple) else v) for k, v in ctor_kwargs.items()}
        save(name, dict(kind=ctor.__name__, kwargs=kw), arrays)

    def mark_init(layer):
        layer.initialized = True

    r = lambda *s: torch.rand(*s, generator=gen)  # noqa: E731
    n = lambda *s: torch.randn(*s, generator=gen)  # noqa: E731

    layer_case('logit_img', rm.Logit, dict(eps=0.01), r(4, 3, 8, 8))
    layer_case('logit_1e5', rm.Logit, dict(eps=1.0e-5), r(5, 6))
    layer_case('actnorm_img', rm.ActNorm, dict(num_features=(12, 4, 4)), n(4, 12, 4, 4), mark_init)
    layer_case('actnorm_1d', rm.ActNorm, dict(num_features=(6, )), n(5, 6), mark_init)
    layer_case('bnflow_img', rm.BatchNorm, dict(num_features=(3, 8, 8), affine=False), n(4, 3, 8, 8))
    layer_case('bnflow_1d_affine', rm.BatchNorm, dict(num_features=(6, ), affine=True), n(5, 6))
    layer_case('invconv_3', rm.InvertibleConv1x1, dict(in_out_channels=3), n(4, 3, 8, 8))
    layer_case('invconv_12', rm.InvertibleConv1x1, dict(in_out_channels=12), n(3, 12, 4, 4))
    layer_case('invconv_48', rm.InvertibleConv1x1, dict(in_out_channels=48), n(2, 48, 4, 4))
    layer_case('invconv_1d', rm.InvertibleConv1x1, dict(in_out_channels=6), n(5, 6))
    for odd in (False, True):
        layer_case('affine_checker_%d' % odd, rc.AffineCoupling,
                   dict(dims=(3, 8, 8), masking='checkerboard', odd=odd), n(3, 3, 8, 8))
        layer_case('affine_channel_%d' % odd, rc.AffineCoupling,
                   dict(dims=(12, 4, 4), masking='channelwise', odd=odd), n(3, 12, 4, 4))
        layer_case('affine_1d_%d' % odd, rc.AffineCoupling, dict(dims=(16, ), odd=odd), n(6, 16))
    layer_case('affine_1d_2', rc.AffineCoupling, dict(dims=(2, ), odd=False), n(8, 2))
    layer_case('mixlog_checker', rc.MixLogAttnCoupling,
               dict(dims=(3, 8, 8), masking='checkerboard', odd=False, n_mixtures=4), n(2, 3, 8, 8))
    layer_case('mixlog_channel', rc.MixLogAttnCoupling,
               dict(dims=(12, 4, 4), masking='channelwise', odd=True, n_mixtures=8), n(2, 12, 4, 4))

    # ActNorm data-dependent init and flow-BatchNorm train-mode statistics (cross-sample paths)
    torch.manual_seed(3)
    x = n(6, 12, 4, 4) * 1.7 + 0.3
    an = rm.ActNorm((12, 4, 4))
    with torch.no_grad():
        z, l = an(x.clone(), torch.zeros(6))
    bn = rm.BatchNorm((12, 4, 4), affine=False)
    bn.train()
    with torch.no_grad():
        zb, lb = bn(x.clone(), torch.zeros(6))
    save('stats_init', dict(kind='stats'),
         dict(x=x.numpy(), an_z=z.numpy(), an_ldj=l.numpy(), an_log_scale=an.log_scale.detach().numpy(),
              an_bias=an.bias.detach().numpy(), bn_z=zb.numpy(), bn_ldj=lb.numpy(),
              bn_batch_mean=bn.batch_mean.numpy(), bn_batch_var=bn.batch_var.numpy(),
              bn_running_mean=bn.running_mean.numpy(), bn_running_var=bn.running_var.numpy()))

    # ---- whole stacks (eval mode; ActNorm initialised by one reference forward) ---------------
    def model_case(name, cls, dims, datatype, layers, mixtures, x, do_perturb=True):
        torch.manual_seed(11)
        cfg = types.SimpleNamespace(layers=layers, mixtures=mixtures)
        net = cls(dims=dims, datatype=datatype, cfg=cfg)
        net.eval()
        with torch.no_grad():
            net(x.clone())  # ActNorm init (modules.py:238-244)
            if do_perturb:
                perturb(net, gen)
            z, ldj = net(x.clone())
            y, ldj_inv = net.backward(z.clone())
        arrays = dict(x=x.numpy(), fwd_z=z.numpy(), fwd_ldj=ldj.numpy(), inv_y=y.numpy(),
                      inv_ldj=ldj_inv.numpy())
        arrays.update(sd_np(net))
        D = int(np.prod(dims))
        nll = -(-0.5 * (z.view(z.size(0), -1).double()**2).sum(1) - 0.5 * D * np.log(2 * np.pi) + ldj.double())
        meta = dict(kind=cls.__name__, dims=list(dims), datatype=datatype, layers=layers, mixtures=mixtures,
                    bpd=float(nll.mean() / (D * np.log(2.0))))
        save(name, meta, arrays)

    model_case('model_realnvp_2d', flows.RealNVP, (2, ), None, 6, 4, r(512, 2) * 2 - 1)  # BASELINE cfg 1
    model_case('model_realnvp_64d', flows.RealNVP, (64, ), None, 8, 4, n(64, 64))  # cfg 4 proxy (affine)
    model_case('model_glow_16', flows.Glow, (3, 16, 16), 'image', 2, 4, r(4, 3, 16, 16))
    model_case('model_glow_1d', flows.Glow, (6, ), None, 3, 4, n(8, 6))
    model_case('model_realnvp_img', flows.RealNVP, (3, 16, 16), 'image', 1, 4, r(3, 3, 16, 16))
    model_case('model_flowpp_16', flows.Flowpp, (3, 16, 16), 'image', 1, 4, r(2, 3, 16, 16))


if __name__ == '__main__':
    main()
