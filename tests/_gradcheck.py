"""Gradient oracle helpers: torch CPU autograd through oracle/flow_oracle.py (train-mode semantics of main.py:78-92),
shared by the CPU test that pins it against the reference's gradient goldens and by the GPU parity tests."""
import torch

from oracle import flow_oracle as O

TRAINABLE_SKIP = ('P', 'I', 'L_mask', 'U_mask', 'sign_s', 'pivots', 'running_mean', 'running_var', 'batch_mean',
                  'batch_var', 'num_batches_tracked')


def leaf_state(sd, dtype=torch.float32, frozen=()):
    """state dict -> leaf tensors; everything the reference trains requires grad."""
    out = {}
    for k, v in sd.items():
        t = v.clone()
        if t.is_floating_point():
            t = t.to(dtype)
            leaf = k.split('.')[-1]
            if leaf not in TRAINABLE_SKIP and k not in frozen:
                t.requires_grad_(True)
        out[k] = t
    return out


def oracle_layer_train(meta, sd, x, ldj):
    """forward of one layer in TRAIN mode (what make_golden_grad.py ran through the reference)."""
    kind, kw = meta['kind'], meta['kwargs']
    if kind == 'Logit':
        return O.logit_fwd(x, ldj, kw['eps'])
    if kind == 'ActNorm':
        return O.actnorm_fwd(x, ldj, sd['log_scale'], sd['bias'])
    if kind == 'BatchNorm':
        m, v = O.bnflow_batch_stats(x.detach())
        shape = sd['log_gamma'].shape
        return O.bnflow_fwd(x, ldj, m.reshape(shape), v.reshape(shape), sd['log_gamma'], sd['beta'])
    if kind == 'InvertibleConv1x1':
        return O.invconv_fwd(x, ldj, sd['P'], sd['L'], sd['U'], sd['log_s'], sd['sign_s'])
    opt = dict(dims=tuple(kw['dims']), masking=kw.get('masking', 'checkerboard'), odd=kw.get('odd', False),
               mixtures=kw.get('n_mixtures', 4))
    ck = 'affine' if kind == 'AffineCoupling' else 'mixlog'
    return O._coupling(ck, opt, sd, '', x, ldj, False, True)


def oracle_layer_grads(meta, sd, x, ldj0, Rz, Rl, dtype=torch.float32):
    """-> (z, ldj, gx, gldj, {param: grad}) of loss = sum(z Rz) + sum(ldj Rl)."""
    leaves = leaf_state(sd, dtype)
    xl = x.to(dtype).clone().requires_grad_(True)
    ll = ldj0.to(dtype).clone().requires_grad_(True)
    z, l = oracle_layer_train(meta, leaves, xl, ll)
    loss = (z * Rz.to(dtype)).sum() + (l * Rl.to(dtype)).sum()
    loss.backward()
    grads = {k: t.grad for k, t in leaves.items() if t.requires_grad and t.grad is not None}
    return z.detach(), l.detach(), xl.grad, ll.grad, grads


def oracle_model_grads(meta, sd, x, dtype=torch.float32, train=True, coupling=None):
    """-> (loss, {param: grad}) of the training loss of main.py:85 through the oracle stack."""
    model = {'RealNVP': 'realnvp', 'Glow': 'glow', 'Flowpp': 'flowpp'}[meta['kind']]
    spec = O.stack_spec(model, tuple(meta['dims']), meta['datatype'], meta['layers'], meta.get('mixtures', 4), coupling)
    leaves = leaf_state(sd, dtype)
    z, ldj = O.stack_forward(spec, leaves, x.to(dtype), train)
    loss = O.mean_nll(z, ldj)
    loss.backward()
    grads = {k: t.grad for k, t in leaves.items() if t.requires_grad and t.grad is not None}
    return loss.detach(), grads


def grad_floor(grads):
    """Absolute noise floor for one backward pass: 1e-5 of the largest gradient entry anywhere.  In train mode every
    conditioner bias that feeds a batch-statistics BatchNorm has a mathematically ZERO gradient (the mean is removed), so
    both sides hold only the rounding residue of large cancelling terms there."""
    return 1e-5 * max([float(v.detach().abs().max()) for v in grads.values() if v.numel()] + [0.0])


def grad_close(got, want, rtol, what='', floor=2e-6):
    """Gradients are compared relative to the LARGEST entry of the tensor (sums of many terms of mixed sign: entries
    near zero carry the rounding of the big ones).  Absolute floor 2e-6: a bias feeding a train-mode BatchNorm has a
    mathematically zero gradient, what both sides hold there is rounding noise of the cancelling terms."""
    got, want = got.detach().cpu().double().reshape(-1), want.detach().cpu().double().reshape(-1)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    scale = float(want.abs().max())
    err = float((got - want).abs().max())
    assert err <= rtol * max(scale, 1e-12) + 1e-30 or err <= max(floor, 2e-6), '%s: max err %.3e vs scale %.3e (rtol %.1e)' % (
        what, err, scale, rtol)


def grad_close_yardstick(got, ref32, truth64, rtol, what='', floor=2e-6, slack=4.0):
    """Deep / ill-conditioned stacks amplify fp32 rounding in ANY implementation (e.g. a weight in front of a train-mode
    BatchNorm is scale-invariant: its gradient is the small residue of large cancelling terms), so the yardstick is the
    fp64 run of the oracle: `got` may be at most `slack` times as far from the fp64 truth as the reference's own fp32
    result `ref32` is, plus the usual relative tolerance."""
    got, ref32, truth64 = [t.detach().cpu().double().reshape(-1) for t in (got, ref32, truth64)]
    assert got.shape == truth64.shape, (what, got.shape, truth64.shape)
    scale = float(truth64.abs().max())
    e_got = float((got - truth64).abs().max())
    e_ref = float((ref32 - truth64).abs().max())
    assert e_got <= slack * e_ref + rtol * max(scale, 1e-12) + max(floor, 2e-6), \
        '%s: |gpu - fp64| %.3e, |fp32 reference - fp64| %.3e, scale %.3e (rtol %.1e)' % (what, e_got, e_ref, scale, rtol)
