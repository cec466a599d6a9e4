"""Generate GRADIENT golden vectors from the UNMODIFIED reference (run in the build container).

    python tests/golden/make_golden_grad.py       # needs /root/reference and the forward goldens next to this file

What the reference's training step computes (main.py:78-92: train mode, ``loss.backward()``), pinned layer by layer and
for whole stacks: every case reloads the state dict + inputs of an existing forward golden (``make_golden.py``), runs the
reference layer / model in TRAIN mode (conditioner BatchNorm on batch statistics, flow BatchNorm on batch statistics
through its buffers), and stores

  layers:  loss = sum(z_out * Rz) + sum(ldj_out * Rl) with seeded cotangents Rz, Rl;  d loss / d(x, ldj0, every parameter)
  models:  loss = -mean(log N(z; 0, I) + ldj)  (main.py:85);  d loss / d(every parameter), plus the BatchNorm running
           statistics after the step.

Files: ``tests/golden/grad_*.npz``.  Nothing at test/bench time imports the reference; only this script does.
"""
import json
import math
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
from flows import coupling as rc, modules as rm  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
from tests import _golden  # noqa: E402

torch.set_num_threads(4)


def save(name, meta, arrays):
    arrays = dict(arrays)
    arrays['meta'] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **arrays)
    print('%-40s %8.1f KB' % (name, os.path.getsize(path) / 1024))


def load_into(module, sd):
    module.load_state_dict(sd, strict=True)
    for m in module.modules():
        if isinstance(m, rm.ActNorm):
            m.initialized = True


def param_grads(module):
    return {'grad/' + k: p.grad.detach().numpy().copy() for k, p in module.named_parameters() if p.grad is not None}


def running_stats(module):
    return {'after/' + k: v.detach().numpy().copy() for k, v in module.state_dict().items()
            if k.split('.')[-1] in ('running_mean', 'running_var', 'batch_mean', 'batch_var')}


def layer_case(src, ctor, gen):
    meta, arrays, sd = _golden.load(src)
    kwargs = {k: (tuple(v) if isinstance(v, list) else v) for k, v in meta['kwargs'].items()}
    torch.manual_seed(0)
    layer = ctor(**kwargs)
    load_into(layer, sd)
    layer.train()
    x = arrays['x'].clone().requires_grad_(True)
    l0 = arrays['ldj0'].clone().requires_grad_(True)
    z, l = layer(x, l0 * 1.0)  # the reference accumulates the log-det in place: hand it a non-leaf
    Rz = torch.randn(z.shape, generator=gen)
    Rl = torch.randn(l.shape, generator=gen)
    loss = (z * Rz).sum() + (l * Rl).sum()
    loss.backward()
    out = dict(Rz=Rz.numpy(), Rl=Rl.numpy(), fwd_z=z.detach().numpy(), fwd_ldj=l.detach().numpy(),
               loss=np.float64(loss.item()), gx=x.grad.numpy(), gldj=l0.grad.numpy())
    out.update(param_grads(layer))
    out.update(running_stats(layer))
    save('grad_' + src, dict(kind=meta['kind'], kwargs=meta['kwargs'], source=src, mode='train'), out)


def model_case(src, cls):
    meta, arrays, sd = _golden.load(src)
    torch.manual_seed(0)
    cfg = types.SimpleNamespace(layers=meta['layers'], mixtures=meta['mixtures'])
    net = cls(dims=tuple(meta['dims']), datatype=meta['datatype'], cfg=cfg)
    load_into(net, sd)
    net.train()
    x = arrays['x'].clone()
    z, ldj = net(x)
    D = int(np.prod(meta['dims']))
    logp = -0.5 * (z.reshape(z.size(0), -1)**2).sum(1) - 0.5 * D * math.log(2.0 * math.pi)
    loss = -1.0 * torch.mean(logp + ldj)  # main.py:85
    loss.backward()
    out = dict(fwd_z=z.detach().numpy(), fwd_ldj=ldj.detach().numpy(), loss=np.float64(loss.item()))
    out.update(param_grads(net))
    out.update(running_stats(net))
    m = dict(meta)
    m.update(source=src, mode='train')
    save('grad_' + src, m, out)


def main():
    gen = torch.Generator().manual_seed(4321)
    layer_case('logit_img', rm.Logit, gen)
    layer_case('actnorm_img', rm.ActNorm, gen)
    layer_case('actnorm_1d', rm.ActNorm, gen)
    layer_case('bnflow_img', rm.BatchNorm, gen)
    layer_case('bnflow_1d_affine', rm.BatchNorm, gen)
    layer_case('invconv_3', rm.InvertibleConv1x1, gen)
    layer_case('invconv_12', rm.InvertibleConv1x1, gen)
    layer_case('invconv_48', rm.InvertibleConv1x1, gen)
    layer_case('invconv_1d', rm.InvertibleConv1x1, gen)
    for src in ('affine_checker_0', 'affine_checker_1', 'affine_channel_0', 'affine_channel_1', 'affine_1d_0',
                'affine_1d_1', 'affine_1d_2'):
        layer_case(src, rc.AffineCoupling, gen)
    layer_case('mixlog_checker', rc.MixLogAttnCoupling, gen)
    layer_case('mixlog_channel', rc.MixLogAttnCoupling, gen)
    model_case('model_realnvp_2d', flows.RealNVP)
    model_case('model_realnvp_64d', flows.RealNVP)
    model_case('model_glow_16', flows.Glow)
    model_case('model_glow_1d', flows.Glow)
    model_case('model_realnvp_img', flows.RealNVP)
    model_case('model_flowpp_16', flows.Flowpp)


if __name__ == '__main__':
    main()
