"""GPU parity tests (run on the B200 box with ``-m gpu``): the CUDA path through the C ABI against the CPU
oracle and against the committed golden vectors from the reference.

Tolerances: index permutations bit-exact (torch.equal); floating point: elementwise outputs 1e-5 (abs+rel; libdevice
vs Sleef transcendentals differ by ~1-2 ulp and go through exp), per-sample log-det 1e-5 relative; bits/dim 1e-5
relative (BASELINE.json north_star) -- in practice ~1e-7.
"""
import math
import types

import pytest
import torch

from tests import _golden
from oracle import flow_oracle as O

pytestmark = pytest.mark.gpu

DEV = 'cuda'


@pytest.fixture(autouse=True)
def _inference_path():
    """These tests cover the inference kernels (in-place log-det, fused peepholes); gradients are tests/test_gpu_backward.py."""
    with torch.no_grad():
        yield


def nfb():
    import nfb200
    return nfb200


def close(a, b, rtol=1e-5, atol=1e-5, what=''):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    assert a.shape == b.shape, (a.shape, b.shape)
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    assert bool((err <= tol).all()), '%s max err %.3e (tol %.1e/%.1e)' % (what, float(err.max()), rtol, atol)


# ---------------------------------------------------------------------------------------------------------
# permutations: bit exact
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('shape', [(2, 1, 2, 2), (2, 3, 4, 4), (3, 2, 6, 4), (2, 12, 8, 8), (2, 3, 32, 32),
                                   (1, 48, 8, 8), (2, 6, 16, 24)])
@pytest.mark.parametrize('odd', [False, True])
def test_permutations_bit_exact(shape, odd):
    F = nfb().flows
    z = torch.randn(shape)
    zg = z.to(DEV)
    z0, z1 = F.checker_split(zg, odd)
    o0, o1 = O.checker_split(z, odd)
    assert torch.equal(z0.cpu(), o0) and torch.equal(z1.cpu(), o1)
    assert torch.equal(F.checker_merge(z0, z1, odd).cpu(), z)
    if shape[1] % 2 == 0:
        c0, c1 = F.channel_split(zg, 1, odd)
        p0, p1 = O.channel_split(z, odd)
        assert torch.equal(c0.cpu(), p0) and torch.equal(c1.cpu(), p1)
        assert torch.equal(F.channel_merge(c0, c1, 1, odd).cpu(), z)
    sq = F.Squeeze2d(odd)(zg, None)[0]
    assert torch.equal(sq.cpu(), O.squeeze2d_layer(z, odd))
    assert torch.equal(F.Unsqueeze2d(odd)(sq, None)[0].cpu(), z)
    assert torch.equal(F.Squeeze2d(odd).backward(sq, None)[0].cpu(), z)


@pytest.mark.parametrize('C', [2, 10, 64])
@pytest.mark.parametrize('odd', [False, True])
def test_split1d_bit_exact(C, odd):
    F = nfb().flows
    z = torch.randn(5, C)
    a, b = F.squeeze1d(z.to(DEV), odd)
    oa, ob = O.split1d(z, odd)
    assert torch.equal(a.cpu(), oa) and torch.equal(b.cpu(), ob)
    assert torch.equal(F.unsqueeze1d(a, b, odd).cpu(), z)


def test_permutations_golden():
    F = nfb().flows
    _, a, _ = _golden.load('permutations')
    for tag in ('1x2x2', '3x4x4', '2x6x4', '12x8x8'):
        z = a['in_' + tag].to(DEV)
        for odd in (0, 1):
            z0, z1 = F.checker_split(z, bool(odd))
            assert torch.equal(z0.cpu(), a['checker_%s_%d_z0' % (tag, odd)])
            assert torch.equal(z1.cpu(), a['checker_%s_%d_z1' % (tag, odd)])
            assert torch.equal(F.Squeeze2d(bool(odd))(z, None)[0].cpu(), a['squeeze2d_%s_%d' % (tag, odd)])


# ---------------------------------------------------------------------------------------------------------
# single layers vs golden vectors from the reference
# ---------------------------------------------------------------------------------------------------------
def build_layer(meta):
    F = nfb().flows
    kind, kw = meta['kind'], dict(meta['kwargs'])
    for k in ('dims', 'num_features'):
        if k in kw:
            kw[k] = tuple(kw[k])
    layer = getattr(F, kind)(**kw)
    return layer


LAYER_CASES = [n for n in _golden.names() if not n.startswith(('model_', 'permutations', 'stats_', 'grad_', 'mixlogcdf_'))]


@pytest.mark.parametrize('name', _golden.names('mixlogcdf_'))
def test_mixlogcdf_module_vs_reference_golden(name):
    """nfb200.flows.MixLogCDF (standalone module of modules.py:186-212) against the reference's outputs."""
    F = nfb().flows
    _, a, _ = _golden.load(name)
    m = F.MixLogCDF()
    d = {k: v.to(DEV) for k, v in a.items()}
    y, l1 = m(d['x'], d['log_pi'], d['mu'], d['s'], d['ldj0'].clone())
    close(y, a['fwd_y'], what=name + ' fwd y')
    close(l1, a['fwd_ldj'], rtol=1e-5, atol=2e-5, what=name + ' fwd ldj')
    x, l2 = m.backward(d['fwd_y'], d['log_pi'], d['mu'], d['s'], d['ldj0'].clone())
    close(x, a['inv_x'], rtol=2e-4, atol=2e-4, what=name + ' inv x')  # the reference's bisection bracket is 1e-4 wide
    close(l2, a['inv_ldj'], rtol=2e-4, atol=4e-3, what=name + ' inv ldj')
    assert torch.equal(d['ldj0'], a['ldj0'].to(DEV))  # returns NEW log-det tensors (modules.py:194)


@pytest.mark.parametrize('name', LAYER_CASES)
def test_layer_vs_reference_golden(name):
    meta, a, sd = _golden.load(name)
    layer = build_layer(meta)
    layer.load_state_dict(sd)
    if hasattr(layer, 'initialized'):
        layer.initialized = True
    layer.to(DEV).eval()
    z, ldj = layer(a['x'].to(DEV), a['ldj0'].to(DEV).clone())
    close(z, a['fwd_z'], what=name + ' fwd z')
    close(ldj, a['fwd_ldj'], rtol=1e-5, atol=2e-5, what=name + ' fwd ldj')
    y, ldj2 = layer.backward(a['fwd_z'].to(DEV), a['ldj0'].to(DEV).clone())
    tol = 2e-4 if meta['kind'] == 'MixLogAttnCoupling' else 3e-5 if meta['kind'] == 'InvertibleConv1x1' else 1e-5
    close(y, a['inv_y'], rtol=tol, atol=tol, what=name + ' inv y')
    close(ldj2, a['inv_ldj'], rtol=max(tol, 1e-5), atol=20 * tol, what=name + ' inv ldj')


def test_stats_paths_vs_golden():
    F = nfb().flows
    _, a, _ = _golden.load('stats_init')
    x = a['x'].to(DEV)
    an = F.ActNorm((12, 4, 4)).to(DEV)
    z, l = an(x, torch.zeros(6, device=DEV))
    assert an.initialized
    close(an.log_scale, a['an_log_scale'], what='actnorm init log_scale')
    close(an.bias, a['an_bias'], what='actnorm init bias')
    close(z, a['an_z'])
    close(l, a['an_ldj'], atol=1e-4)
    bn = F.BatchNorm((12, 4, 4), affine=False).to(DEV)
    bn.train()
    zb, lb = bn(x, torch.zeros(6, device=DEV))
    close(bn.batch_mean, a['bn_batch_mean'])
    close(bn.batch_var, a['bn_batch_var'])
    close(bn.running_mean, a['bn_running_mean'])
    close(bn.running_var, a['bn_running_var'])
    close(zb, a['bn_z'])
    close(lb, a['bn_ldj'], atol=1e-4)


# ---------------------------------------------------------------------------------------------------------
# whole stacks vs golden (reference) and vs oracle
# ---------------------------------------------------------------------------------------------------------
def build_model(meta, **extra):
    n = nfb()
    cls = {'RealNVP': n.RealNVP, 'Glow': n.Glow, 'Flowpp': n.Flowpp}[meta['kind']]
    cfg = types.SimpleNamespace(layers=meta['layers'], mixtures=meta['mixtures'], **extra)
    return cls(tuple(meta['dims']), meta['datatype'], cfg)


@pytest.mark.parametrize('name', _golden.names('model_'))
def test_model_vs_reference_golden(name):
    meta, a, sd = _golden.load(name)
    net = build_model(meta)
    net.load_state_dict(sd)
    net.mark_initialized().to(DEV).eval()
    z, ldj = net(a['x'].to(DEV))
    close(z, a['fwd_z'], rtol=5e-5, atol=5e-5, what=name + ' z')
    close(ldj, a['fwd_ldj'], rtol=1e-5, atol=2e-4, what=name + ' ldj')
    bpd = net.bits_per_dim(a['x'].to(DEV))
    assert abs(bpd - meta['bpd']) <= 1e-5 * abs(meta['bpd']), (bpd, meta['bpd'])
    if meta['kind'] != 'Flowpp':
        y, ldj2 = net.backward(a['fwd_z'].to(DEV))
        close(y, a['inv_y'], rtol=2e-4, atol=2e-4, what=name + ' inv y')
        close(ldj2, a['inv_ldj'], rtol=1e-4, atol=2e-3, what=name + ' inv ldj')
    else:  # bisection: compare against the input instead (reference round-trip error ~2e-5)
        y, _ = net.backward(z)
        # the leading Logit(0.01) clamps to [0.01, 0.99] (glow.py:20), so that is what the inverse can recover
        close(y, a['x'].clamp(0.01, 0.99), rtol=5e-4, atol=5e-4, what=name + ' round trip')


def oracle_spec(model, dims, datatype, layers, mixtures=4, coupling=None):
    return O.stack_spec(model, dims, datatype, layers, mixtures, coupling)


@pytest.mark.parametrize('cfg', [
    dict(model='glow', dims=(3, 32, 32), datatype='image', layers=32, B=256),                      # BASELINE configs[1]
    dict(model='flowpp', dims=(3, 32, 32), datatype='image', layers=32, mixtures=8, B=32),         # configs[2] (CPU-bounded B)
    dict(model='realnvp', dims=(64, ), datatype=None, layers=8, coupling='rqs', B=65536),           # configs[3]
    dict(model='glow', dims=(3, 64, 64), datatype='image', layers=48, B=8),                        # configs[4] (CPU-bounded B)
    dict(model='realnvp', dims=(2, ), datatype=None, layers=6, B=512),                              # configs[0]
], ids=['glow32_K32_B256', 'flowpp32_K32_M8_B32', 'realnvp64_rqs_B65536', 'glow64_K48_B8', 'realnvp2_B512'])
def test_full_depth_config_bits_per_dim(cfg):
    """The BASELINE.json configurations at FULL depth: bits/dim of the GPU path within 1e-5 (relative) of the CPU oracle on
    the same weights and inputs -- the north-star parity bar, here inside pytest and not only in bench.py."""
    n = nfb()
    torch.manual_seed(0)
    cls = {'glow': n.Glow, 'realnvp': n.RealNVP, 'flowpp': n.Flowpp}[cfg['model']]
    extra = {'coupling': cfg['coupling']} if cfg.get('coupling') else {}
    net = cls(cfg['dims'], cfg['datatype'], types.SimpleNamespace(layers=cfg['layers'], mixtures=cfg.get('mixtures', 4), **extra))
    net.eval()
    g = torch.Generator().manual_seed(0)
    shape = (cfg['B'], ) + tuple(cfg['dims'])
    x = torch.rand(shape, generator=g) if cfg['datatype'] == 'image' else torch.randn(shape, generator=g)
    net.to(DEV)
    net(x.to(DEV))  # ActNorm data-dependent init on the device (modules.py:238-244); the oracle gets the resulting state
    sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    bpd_gpu = net.bits_per_dim(x.to(DEV))
    spec = O.stack_spec(cfg['model'], cfg['dims'], cfg['datatype'], cfg['layers'], cfg.get('mixtures', 4), cfg.get('coupling'))
    torch.set_num_threads(max(1, (__import__('os').cpu_count() or 1)))
    z, ldj = O.stack_forward(spec, sd, x)
    bpd_cpu = O.bits_per_dim(z, ldj)
    rel = abs(bpd_gpu - bpd_cpu) / abs(bpd_cpu)
    print('%s: bits/dim gpu %.9f cpu %.9f rel %.2e' % (cfg['model'], bpd_gpu, bpd_cpu, rel))
    assert rel <= 1e-5, (bpd_gpu, bpd_cpu, rel)


def test_no_library_fallbacks_on_baseline_configs():
    """Eval-mode forwards of the five BASELINE stacks stay on libnfb200 kernels: the library-path counter does not move."""
    n = nfb()
    n0 = n.library_path_calls()
    for model, dims, dt, layers, extra, B in (('Glow', (3, 32, 32), 'image', 2, {}, 4), ('Flowpp', (3, 32, 32), 'image', 1, {}, 2),
                                              ('RealNVP', (64, ), None, 2, {'coupling': 'rqs'}, 64),
                                              ('Glow', (3, 64, 64), 'image', 1, {}, 2), ('RealNVP', (2, ), None, 2, {}, 32)):
        torch.manual_seed(0)
        net = getattr(n, model)(dims, dt, types.SimpleNamespace(layers=layers, mixtures=8, **extra)).to(DEV).eval()
        x = torch.rand((B, ) + dims, device=DEV) if dt == 'image' else torch.randn((B, ) + dims, device=DEV)
        net(x)
        net(x)
    assert n.library_path_calls() == n0, n.library_path_log()


def test_reference_stack_rebound_to_nfb200_layers():
    """INTEGRATION.md section 1 executed on the device: the REFERENCE's own Glow builder (glow.py:17-60) with its layer
    classes rebound to nfb200's (glow.py:27-29,37-39,43-45), a reference state dict loaded strictly, forward on the GPU --
    against the output the unmodified reference produced on the CPU (tests/golden/model_glow_16.npz)."""
    import importlib
    import os
    import sys
    ref_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle', '_ref')
    if not os.path.isdir(os.path.join(ref_dir, 'flows')):
        pytest.skip('oracle/_ref (the staged reference package) is not present')
    n = nfb()
    nf = n.flows
    sys.path.insert(0, ref_dir)
    try:
        import flows.glow as rglow
        importlib.reload(rglow)
        saved = {}
        for name in ('Logit', 'ActNorm', 'BatchNorm', 'InvertibleConv1x1', 'Compose', 'AffineCoupling', 'AdditiveCoupling',
                     'MixLogAttnCoupling', 'Squeeze2d', 'Unsqueeze2d'):
            if hasattr(rglow, name):
                saved[name] = getattr(rglow, name)
                setattr(rglow, name, getattr(nf, name))
        try:
            meta, a, sd = _golden.load('model_glow_16')
            cfg = types.SimpleNamespace(layers=meta['layers'], mixtures=meta['mixtures'])
            net = rglow.Glow(dims=tuple(meta['dims']), datatype=meta['datatype'], cfg=cfg)  # the reference's constructor
            assert type(net.net) is nf.Compose and any(type(m) is nf.AffineCoupling for m in net.modules())
            net.load_state_dict(sd, strict=True)
            for m in net.modules():
                if isinstance(m, nf.ActNorm):
                    m.initialized = True
            net.to(DEV).eval()
            z, ldj = net(a['x'].to(DEV))  # Glow.forward of the reference (glow.py:62-64) over nfb200 layers
            close(z, a['fwd_z'], rtol=5e-5, atol=5e-5, what='rebound glow z')
            close(ldj, a['fwd_ldj'], rtol=1e-5, atol=2e-4, what='rebound glow ldj')
            y, ldj2 = net.backward(a['fwd_z'].to(DEV))
            close(y, a['inv_y'], rtol=2e-4, atol=2e-4, what='rebound glow inverse')
        finally:
            for name, cls in saved.items():
                setattr(rglow, name, cls)
    finally:
        sys.path.remove(ref_dir)


def perturb_(net, seed=0):
    """Move BatchNorm/ActNorm/conv parameters away from their near-identity initial values (seeded)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, t in list(net.named_parameters()) + list(net.named_buffers()):
            leaf = name.split('.')[-1]
            if not t.is_floating_point() or leaf in ('P', 'I', 'L_mask', 'U_mask', 'sign_s', 'pivots'):
                continue
            if leaf in ('running_var', 'batch_var'):
                t.copy_(0.5 + torch.rand(t.shape, generator=g))
            elif leaf in ('L', 'U'):
                t.add_(0.05 / math.sqrt(t.shape[0]) * torch.randn(t.shape, generator=g))
            elif leaf in ('running_mean', 'beta', 'bias', 'log_gamma', 'log_scale', 'log_s'):
                t.add_(0.05 * torch.randn(t.shape, generator=g))
            elif leaf in ('s_log_scale', 'a_log_scale'):
                t.copy_(0.3 + 0.1 * torch.randn(t.shape, generator=g))


@pytest.mark.parametrize('cfg', [
    dict(model='glow', dims=(3, 32, 32), datatype='image', layers=4, B=8),   # BASELINE cfg 2 stack at K=4
    dict(model='glow', dims=(3, 64, 64), datatype='image', layers=2, B=2),   # cfg 5 stack (C up to 192) at K=2
    dict(model='realnvp', dims=(64, ), datatype=None, layers=8, B=4096),      # cfg 4 proxy (affine)
    dict(model='realnvp', dims=(2, ), datatype=None, layers=6, B=512),        # cfg 1
    dict(model='flowpp', dims=(3, 32, 32), datatype='image', layers=1, B=2, mixtures=8),  # cfg 3 stack at K=1
    dict(model='glow', dims=(1, 32, 32), datatype='image', layers=2, B=4),   # the reference's padded MNIST (dataset.py:67-72): C = 1
])
def test_model_vs_oracle(cfg):
    n = nfb()
    torch.manual_seed(0)
    cls = {'glow': n.Glow, 'realnvp': n.RealNVP, 'flowpp': n.Flowpp}[cfg['model']]
    mixtures = cfg.get('mixtures', 4)
    net = cls(cfg['dims'], cfg['datatype'], types.SimpleNamespace(layers=cfg['layers'], mixtures=mixtures))
    perturb_(net)
    net.mark_initialized().eval()
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    gen = torch.Generator().manual_seed(1)
    x = torch.rand((cfg['B'], ) + cfg['dims'], generator=gen) if cfg['datatype'] == 'image' else \
        torch.randn((cfg['B'], ) + cfg['dims'], generator=gen)
    spec = oracle_spec(cfg['model'], cfg['dims'], cfg['datatype'], cfg['layers'], mixtures)
    with torch.no_grad():
        zo, lo = O.stack_forward(spec, sd, x)
        z64, l64 = O.stack_forward(spec, O.to_dtype(sd, torch.float64), x.double())
    net.to(DEV)
    z, ldj = net(x.to(DEV))
    # Deep stacks amplify fp32 rounding in ANY implementation, so the yardstick is the fp64 run of the oracle:
    # the CUDA path may be at most a few times further from the fp64 truth than the reference's own fp32 CPU path.
    scale = float(z64.abs().max())
    e_ref = float((zo.double() - z64).abs().max())
    e_gpu = float((z.cpu().double() - z64).abs().max())
    assert e_gpu <= 4.0 * e_ref + 2e-6 * scale, 'z: gpu-vs-fp64 %.3e, cpu-fp32-vs-fp64 %.3e, scale %.3e' % (e_gpu, e_ref, scale)
    lscale = float(l64.abs().max())
    l_ref = float((lo.double() - l64).abs().max())
    l_gpu = float((ldj.cpu().double() - l64).abs().max())
    assert l_gpu <= 4.0 * l_ref + 2e-6 * lscale, 'ldj: gpu %.3e ref %.3e scale %.3e' % (l_gpu, l_ref, lscale)
    bo = O.bits_per_dim(zo, lo)
    bg = net.bits_per_dim(x.to(DEV))
    assert abs(bg - bo) <= 1e-5 * abs(bo), (bg, bo)   # the BASELINE.json bar
    y, _ = net.backward(z)
    xr = x.clamp(0.01, 0.99) if cfg['datatype'] == 'image' else x
    yo, _ = O.stack_backward(spec, O.to_dtype(sd, torch.float64), z64)
    r_ref = float((yo - xr.double()).abs().max())   # fp64 round trip: only the conditioning of the stack
    r_gpu = float((y.cpu().double() - xr.double()).abs().max())
    assert r_gpu <= 1e-4 + 50.0 * max(r_ref, e_gpu / max(scale, 1.0)), 'round trip %.3e (fp64 %.3e)' % (r_gpu, r_ref)


# ---------------------------------------------------------------------------------------------------------
# RQ-spline coupling: no reference; oracle restatement of Durkan et al. in fp64
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('dims,masking', [((64, ), 'checkerboard'), ((4, 8, 8), 'checkerboard'),
                                          ((4, 8, 8), 'channelwise'), ((6, ), 'checkerboard')])
@pytest.mark.parametrize('odd', [False, True])
def test_rqs_coupling_vs_oracle(dims, masking, odd):
    F = nfb().flows
    torch.manual_seed(3)
    layer = F.RQSplineCoupling(dims, masking=masking, odd=odd, n_bins=8, tail_bound=3.0)
    perturb_(layer, 5)
    layer.eval()
    sd = {k: v.clone() for k, v in layer.state_dict().items()}
    x = torch.randn((16, ) + dims, generator=torch.Generator().manual_seed(2)) * 2.0  # some mass in the tails
    opt = dict(dims=dims, masking=masking, odd=odd, mixtures=4)
    sd64 = O.to_dtype(sd, torch.float64)
    with torch.no_grad():
        zo, lo = O._coupling('rqs', opt, sd64, '', x.double(), torch.zeros(16, dtype=torch.float64), False)
    layer.to(DEV)
    z, ldj = layer(x.to(DEV), torch.zeros(16, device=DEV))
    close(z, zo, rtol=2e-5, atol=2e-5, what='rqs fwd')
    close(ldj, lo, rtol=2e-5, atol=2e-4, what='rqs ldj')
    y, ldj2 = layer.backward(z, ldj.clone())
    close(y, x, rtol=1e-4, atol=1e-4, what='rqs round trip')
    close(ldj2, torch.zeros(16), atol=1e-3, what='rqs log-det cancels')


# ---------------------------------------------------------------------------------------------------------
# size-independent properties at BASELINE sizes (no CPU oracle needed)
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('dims,masking,B', [((3, 32, 32), 'checkerboard', 256), ((12, 16, 16), 'channelwise', 256),
                                            ((48, 8, 8), 'checkerboard', 256), ((64, ), 'checkerboard', 65536),
                                            ((3, 64, 64), 'checkerboard', 64), ((192, 8, 8), 'channelwise', 64)])
def test_affine_kernel_properties_full_size(dims, masking, B):
    import nfb200._lib as L
    mode = L.SPLIT_1D if len(dims) == 1 else (L.SPLIT_CHECKER if masking == 'checkerboard' else L.SPLIT_CHANNEL)
    C, H, W = (dims + (1, 1))[:3] if len(dims) == 1 else dims
    g = torch.Generator(device=DEV).manual_seed(0)
    z = torch.randn((B, ) + dims, device=DEV, generator=g)
    params = torch.randn((B, ) + dims, device=DEV, generator=g)  # (t | s_raw) has exactly D entries per sample
    a = torch.tensor([0.3], device=DEV)
    b = torch.tensor([-0.05], device=DEV)
    for odd in (0, 1):
        out = torch.empty_like(z)
        ldj = torch.zeros(B, device=DEV)
        L.check(L.lib().nfb_affine_coupling_fwd(z.data_ptr(), out.data_ptr(), params.data_ptr(), ldj.data_ptr(),
                                                ldj.data_ptr(), a.data_ptr(), b.data_ptr(), B, C, H, W, mode, odd,
                                                L.stream()))
        # (1) in-place == out-of-place, bit for bit
        z2 = z.clone()
        ldj2 = torch.zeros(B, device=DEV)
        L.check(L.lib().nfb_affine_coupling_fwd(z2.data_ptr(), z2.data_ptr(), params.data_ptr(), ldj2.data_ptr(),
                                                ldj2.data_ptr(), a.data_ptr(), b.data_ptr(), B, C, H, W, mode, odd,
                                                L.stream()))
        assert torch.equal(out, z2) and torch.equal(ldj, ldj2)
        # (2) the pass-through half is untouched bit for bit; the log-det equals sum(s) recomputed by torch
        F = nfb().flows
        from nfb200.flows.squeeze import coupling_split
        i0, i1 = coupling_split(z, mode, bool(odd))
        o0, o1 = coupling_split(out, mode, bool(odd))
        assert torch.equal(i1, o1)
        n0 = i0[0].numel()
        pf = params.view(B, -1)
        s = torch.tanh(pf[:, n0:]) * a + b
        close(ldj, s.sum(1), rtol=1e-5, atol=1e-3, what='ldj')
        close(o0.reshape(B, -1), i0.reshape(B, -1) * torch.exp(s) + pf[:, :n0], rtol=1e-5, atol=1e-5, what='z0')
        # (3) inverse undoes forward, log-det cancels
        back = torch.empty_like(z)
        L.check(L.lib().nfb_affine_coupling_inv(out.data_ptr(), back.data_ptr(), params.data_ptr(), ldj.data_ptr(),
                                                ldj.data_ptr(), a.data_ptr(), b.data_ptr(), B, C, H, W, mode, odd,
                                                L.stream()))
        close(back, z, rtol=1e-5, atol=1e-5, what='round trip')
        close(ldj, torch.zeros(B), atol=1e-3, what='ldj cancels')


def test_nll_matches_oracle():
    n = nfb()
    z = torch.randn(300, 3, 32, 32)
    ldj = torch.randn(300) * 100
    rows, total = n.gauss_nll(z.to(DEV), ldj.to(DEV))
    ref = O.nll_rows(z, ldj)
    close(rows, ref.float(), rtol=1e-6, atol=1e-3)
    s, cnt = total.tolist()
    assert cnt == 300 and abs(s - float(ref.sum())) <= 1e-6 * abs(float(ref.sum()))


def test_errors_are_loud():
    import nfb200._lib as L
    n = nfb()
    with pytest.raises(RuntimeError):
        n.flows.Logit()(torch.rand(2, 4), torch.zeros(2))  # CPU tensor: no fallback
    z = torch.randn(2, 3, 5, 5, device=DEV)
    rc = L.lib().nfb_coupling_split(z.data_ptr(), z.data_ptr(), None, 2, 3, 5, 5, L.SPLIT_CHECKER, 0, L.stream())
    assert rc == -3
    with pytest.raises(RuntimeError):
        L.check(rc)


# ---------------------------------------------------------------------------------------------------------
# fused conditioner kernels (ConvNet / MLP) vs the CPU oracle and vs the on-device library path
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('cin,cout,hw,B', [(6, 12, 16, 5), (24, 48, 8, 7), (96, 192, 4, 9), (2, 4, 8, 3),
                                           (40, 70, 16, 2), (6, 12, 16, 256), (6, 12, 32, 3), (40, 70, 32, 2),
                                           (6, 12, 32, 64)])
def test_convnet_fused_vs_oracle(cin, cout, hw, B):
    F = nfb().flows
    torch.manual_seed(cin)
    net = F.ConvNet(cin, cout)
    perturb_(net, 9)
    with torch.no_grad():
        for n_, p in net.named_parameters():
            if n_.endswith('weight') and p.dim() == 1:
                p.add_(0.2 * torch.randn(p.shape))  # BatchNorm gains
    net.eval()
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    x = torch.randn(B, cin, hw, hw)
    with torch.no_grad():
        ref = O.resnet_conditioner(sd, '', x)
    net.to(DEV)
    out = net(x.to(DEV))
    lib = net._forward_library(x.to(DEV))
    scale = float(ref.abs().max())
    close(out, ref, rtol=1e-5, atol=2e-6 * max(scale, 1.0), what='fused convnet vs oracle')
    close(out, lib, rtol=2e-5, atol=1e-5 * max(scale, 1.0), what='fused convnet vs cudnn path')


@pytest.mark.parametrize('cin,cout,B', [(1, 2, 512), (32, 64, 1000), (32, 32 * 23, 77), (3, 5, 4)])
def test_mlp_fused_vs_oracle(cin, cout, B):
    F = nfb().flows
    torch.manual_seed(cout)
    net = F.MLP(cin, cout)
    perturb_(net, 4)
    net.eval()
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    x = torch.randn(B, cin)
    with torch.no_grad():
        ref = O.resnet_conditioner(sd, '', x)
    net.to(DEV)
    out = net(x.to(DEV))
    close(out, ref, rtol=1e-5, atol=3e-6 * max(1.0, float(ref.abs().max())), what='fused mlp vs oracle')


@pytest.mark.parametrize('dims,masking', [((3, 32, 32), 'checkerboard'), ((12, 16, 16), 'channelwise'),
                                          ((12, 16, 16), 'checkerboard'), ((48, 8, 8), 'channelwise'),
                                          ((48, 8, 8), 'checkerboard'), ((64, ), 'checkerboard'),
                                          ((3, 64, 64), 'checkerboard'), ((12, 32, 32), 'channelwise')])
@pytest.mark.parametrize('odd', [False, True])
def test_conditioner_gathers_z1_in_kernel(dims, masking, odd):
    """forward_from_z (split addressing inside the kernel) == explicit split followed by the conditioner, bit for bit."""
    F = nfb().flows
    import nfb200._lib as L
    from nfb200.flows.squeeze import coupling_split
    torch.manual_seed(1)
    cpl = F.AffineCoupling(dims, masking=masking, odd=odd).to(DEV).eval()
    z = torch.randn((6, ) + dims, device=DEV)
    p1 = cpl.net.forward_from_z(z, cpl.mode, cpl.odd)
    _, z1 = coupling_split(z, cpl.mode, cpl.odd, want_z0=False)
    p2 = cpl.net(z1)
    assert p1 is not None and torch.equal(p1, p2)


def test_fused_actnorm_invconv_is_bit_identical():
    """Compose's ActNorm+InvertibleConv1x1 peephole must give exactly what the two layers give separately."""
    n = nfb()
    torch.manual_seed(2)
    for dims in [(3, 32, 32), (12, 16, 16), (48, 8, 8), (192, 8, 8), (6, )]:
        an = n.flows.ActNorm(dims)
        conv = n.flows.InvertibleConv1x1(dims[0])
        perturb_(an, 1)
        perturb_(conv, 2)
        an.initialized = True
        comp = n.flows.Compose([an, conv]).to(DEV)
        x = torch.randn((5, ) + dims, device=DEV)
        l0 = torch.randn(5, device=DEV)
        z1, l1 = comp(x, l0.clone())
        y1, m1 = comp.backward(z1, l1.clone())
        comp.fuse_steps = 0
        z2, l2 = comp(x, l0.clone())
        assert torch.equal(z1, z2) and torch.equal(l1, l2), dims
        y2, m2 = comp.backward(z1, l1.clone())
        assert torch.equal(y1, y2) and torch.equal(m1, m2), dims  # fused inverse == separate inverse layers


@pytest.mark.parametrize('C,hw,B', [(192, 8, 5), (192, 8, 256), (128, 4, 9), (192, 16, 3), (256, 8, 2)])
def test_invconv_large_channel_counts(C, hw, B):
    """C >= 128 (the last level of a 64x64 Glow has C = 192): the co-tiled persistent kernel against a plain fp32 matrix
    product, its fused-ActNorm variant against the two layers run one after the other (bit-identical), and the inverse."""
    F = nfb().flows
    torch.manual_seed(C + hw)
    an, conv = F.ActNorm((C, hw, hw)), F.InvertibleConv1x1(C)
    perturb_(an, 1)
    perturb_(conv, 2)
    an.initialized = True
    an.to(DEV)
    conv.to(DEV)
    x = torch.randn(B, C, hw, hw, device=DEV)
    l0 = torch.randn(B, device=DEV)
    z, l = conv(x, l0.clone())
    Wm = conv.matrices()[0]
    ref = torch.einsum('oc,bcp->bop', Wm.double(), x.view(B, C, -1).double()).view_as(x)
    close(z, ref.float(), rtol=2e-5, atol=2e-5 * float(ref.abs().max()), what='1x1 conv C=%d' % C)
    close(l, l0 + conv.log_s.sum() * hw * hw, rtol=1e-5, atol=1e-3, what='ldj')
    y, l2 = conv.backward(z, l.clone())
    close(y, x, rtol=2e-4, atol=2e-4, what='inverse')
    comp = F.Compose([an, conv]).to(DEV)
    z1, l1 = comp(x, l0.clone())
    comp.fuse_steps = 0
    z2, l2 = comp(x, l0.clone())
    assert torch.equal(z1, z2) and torch.equal(l1, l2)


@pytest.mark.parametrize('key,values,cin,cout,hw', [(0, (0, 1), 6, 12, 16), (1, (0, 1, 2, 3), 24, 48, 8),
                                                    (2, (0, 1, 2, 3, 4), 96, 192, 4)])
def test_convnet_variants_agree(key, values, cin, cout, hw):
    """All thread-tile variants of the fused ConvNet kernel compute the same sums in the same order."""
    import nfb200._lib as L
    F = nfb().flows
    torch.manual_seed(0)
    net = F.ConvNet(cin, cout).to(DEV).eval()
    x = torch.randn(37, cin, hw, hw, device=DEV)
    outs = []
    for v in values:
        net.kernel_flags = L.CONV_FFMA | L.conv_variant(v)
        outs.append(net(x))
    for o in outs[1:]:
        assert torch.equal(o, outs[0])


@pytest.mark.parametrize('cin,cout,hw,B', [(6, 12, 16, 3), (24, 48, 8, 5), (96, 192, 4, 7), (6, 12, 16, 256),
                                           (40, 70, 8, 2), (24, 48, 8, 256), (96, 192, 4, 256)])
@pytest.mark.parametrize('operands', ['f16split', 'tf32x3'])
def test_convnet_tensor_core_path(cin, cout, hw, B, operands):
    """tcgen05 implicit-GEMM conditioner (default FP16-split operands; 3xTF32 with NFB_CONV_TF32) vs the FP32-FFMA kernel and
    the CPU oracle (same tolerance class)."""
    import nfb200._lib as L
    fmt = L.CONV_TF32 if operands == 'tf32x3' else 0
    F = nfb().flows
    torch.manual_seed(cin + hw)
    net = F.ConvNet(cin, cout)
    perturb_(net, 9)
    net.eval()
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    x = torch.randn(B, cin, hw, hw)
    with torch.no_grad():
        ref = O.resnet_conditioner(sd, '', x[:8])
        ref64 = O.resnet_conditioner(O.to_dtype(sd, torch.float64), '', x[:8].double())
    net.to(DEV)
    net.kernel_flags = L.CONV_FFMA
    ffma = net(x.to(DEV))
    net.kernel_flags = fmt
    tc = net(x.to(DEV))
    net.kernel_flags = fmt | L.CONV_PAIR
    tc_pair = net(x.to(DEV))
    torch.cuda.synchronize()
    scale = max(1.0, float(ref.abs().max()))
    close(tc_pair, tc, rtol=2e-6, atol=1e-6 * scale, what='paired tiles vs single tile')
    e_tc = float((tc[:8].cpu().double() - ref64).abs().max())
    e_ff = float((ffma[:8].cpu().double() - ref64).abs().max())
    e_ref = float((ref.double() - ref64).abs().max())
    print('tensor-core err vs fp64 %.3e | ffma %.3e | cpu fp32 %.3e | scale %.2f' % (e_tc, e_ff, e_ref, scale))
    close(tc, ffma, rtol=2e-5, atol=4e-6 * scale, what='tensor-core vs ffma')
    assert e_tc <= 8.0 * max(e_ref, e_ff) + 1e-6 * scale


@pytest.mark.parametrize('dims,masking', [((3, 32, 32), 'checkerboard'), ((12, 16, 16), 'channelwise'),
                                          ((12, 16, 16), 'checkerboard'), ((48, 8, 8), 'channelwise'),
                                          ((48, 8, 8), 'checkerboard'), ((6, 8, 8), 'channelwise'),
                                          ((2, 8, 8), 'checkerboard'), ((96, 16, 16), 'channelwise'),
                                          ((192, 8, 8), 'checkerboard')])
@pytest.mark.parametrize('odd', [False, True])
@pytest.mark.parametrize('B', [1, 3, 19, 256])
def test_fused_conditioner_coupling(dims, masking, odd, B):
    """nfb_convnet_affine_fwd (conditioner on the tensor cores + affine coupling + log-det in ONE kernel, in place) vs the
    two-kernel path (nfb_convnet_fwd_ex + nfb_affine_coupling_fwd): same arithmetic -> identical z; the pass-through half is
    bit-identical to the input; the log-det differs only by the reduction order; and both agree with the FFMA conditioner."""
    n = nfb()
    import nfb200._lib as L
    torch.manual_seed(7)
    cpl = n.flows.AffineCoupling(dims, masking=masking, odd=odd)
    perturb_(cpl, 3)
    cpl.to(DEV).eval()
    x = torch.randn((B, ) + dims, device=DEV)
    l0 = torch.randn(B, device=DEV)
    x_keep = x.clone()
    n0 = n._lib.launch_count()
    cpl(x, l0.clone())  # first call packs the conditioner weights (one-time launches)
    n0 = n._lib.launch_count()
    z1, l1 = cpl(x, l0.clone())
    assert n._lib.launch_count() - n0 == 1  # ONE libnfb200 kernel for conditioner + coupling
    assert torch.equal(x, x_keep)           # the layer API leaves its input alone
    z1b, l1b = cpl.forward_fused(x.clone(), l0.clone(), inplace=True)
    assert torch.equal(z1, z1b) and torch.equal(l1, l1b)
    cpl.fused_conditioner = False
    z2, l2 = cpl(x, l0.clone())
    assert torch.equal(z1, z2)
    close(l1, l2, rtol=1e-6, atol=1e-4, what='ldj')
    # pass-through half untouched
    m = n.flows.squeeze
    _, z1_in = m.coupling_split(x, cpl.mode, cpl.odd, want_z0=False)
    _, z1_out = m.coupling_split(z1, cpl.mode, cpl.odd, want_z0=False)
    assert torch.equal(z1_in, z1_out)
    # two tiles per CTA (throughput mode): the second tile of a pair issues its taps in another order (other rounding of the
    # same sums), so equality is to rounding, and the fused / two-kernel paths of THIS mode are again bit-identical
    cpl.fused_conditioner = True
    cpl.net.kernel_flags = L.CONV_PAIR
    z4, l4 = cpl(x, l0.clone())
    sc = max(1.0, float(z1.abs().max()))
    close(z4, z1, rtol=2e-6, atol=1e-6 * sc, what='paired tiles vs single tile')
    close(l4, l1, rtol=2e-6, atol=1e-4, what='ldj (paired tiles)')
    cpl.fused_conditioner = False
    z5, _ = cpl(x, l0.clone())
    assert torch.equal(z4, z5)
    cpl.net.kernel_flags = L.CONV_FFMA
    z3, l3 = cpl(x, l0.clone())
    scale = max(1.0, float(z3.abs().max()))
    close(z1, z3, rtol=2e-5, atol=4e-6 * scale, what='tensor-core vs ffma conditioner')
    close(l1, l3, rtol=2e-5, atol=4e-6 * max(1.0, float(l3.abs().max())), what='ldj tensor-core vs ffma')
    # 3xTF32 operands (NFB_CONV_TF32): fused and two-kernel paths bit-identical again, and within rounding of the default
    cpl.net.kernel_flags = L.CONV_TF32
    z6, l6 = cpl(x, l0.clone())
    cpl.fused_conditioner = True
    z7, l7 = cpl(x, l0.clone())
    assert torch.equal(z6, z7)
    close(z7, z3, rtol=2e-5, atol=4e-6 * scale, what='3xTF32 vs ffma conditioner')
    close(l7, l3, rtol=2e-5, atol=4e-6 * max(1.0, float(l3.abs().max())), what='ldj 3xTF32 vs ffma')


@pytest.mark.parametrize('hw', [16, 8, 4])
def test_fp16_split_range(hw):
    """The default operand format of the tensor-core conditioner represents x as fp16 hi + 2^-10 fp16 lo: tiny inputs (hi is a
    subnormal or zero) keep full precision through lo; an input beyond the fp16 range cannot be represented and must come
    back as NaN for that sample (never a finite wrong number), while NFB_CONV_TF32 and the FFMA kernel handle it."""
    import nfb200._lib as L
    F = nfb().flows
    torch.manual_seed(hw)
    net = F.ConvNet(6, 12)
    perturb_(net, 4)
    net.eval()
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    x = torch.randn(9, 6, hw, hw)
    x[1] *= 1e-6          # far below the fp16 normal range
    x[2] *= 3e4           # |x| up to ~1e5: beyond 65504
    x[2, 0, 0, 0] = 7e4
    with torch.no_grad():
        ref64 = O.resnet_conditioner(O.to_dtype(sd, torch.float64), '', x.double())
    net.to(DEV)
    out = net(x.to(DEV)).cpu()
    ok = [0, 1, 3, 4, 5, 6, 7, 8]
    sc = float(ref64[ok].abs().max())
    assert torch.isfinite(out[ok]).all()
    close(out[ok], ref64[ok].float(), rtol=2e-5, atol=4e-6 * sc, what='fp16 split incl. tiny inputs')
    assert torch.isnan(out[2]).any() and not torch.isinf(out[2]).any()
    for flags in (L.CONV_TF32, L.CONV_FFMA):
        net.kernel_flags = flags
        o2 = net(x.to(DEV)).cpu()
        close(o2[2], ref64[2].float(), rtol=2e-5, atol=4e-6 * float(ref64[2].abs().max()), what='large inputs, flags %#x' % flags)


@pytest.mark.parametrize('dims,masking', [((3, 32, 32), 'checkerboard'), ((12, 16, 16), 'channelwise'),
                                          ((12, 16, 16), 'checkerboard'), ((48, 8, 8), 'channelwise'),
                                          ((48, 8, 8), 'checkerboard'), ((6, 16, 16), 'checkerboard')])
@pytest.mark.parametrize('B', [1, 5, 256])
@pytest.mark.parametrize('pair', [False, True])
def test_whole_flow_step_in_one_launch(dims, masking, B, pair):
    """Compose.fuse_steps = 2: [coupling i + ActNorm i+1 + 1x1 conv i+1] is ONE kernel (nfb_convnet_affine_step_fwd) -- against
    fuse_steps = 1 (two kernels per step) and 0 (every layer on its own): z bit-identical (same fmaf chains), log-det to
    rounding.  (6, 16, 16): a channel count without a fused variant -> falls back to two kernels, same result."""
    n = nfb()
    import nfb200._lib as L
    torch.manual_seed(3)
    layers = []
    for i in range(3):
        layers += [n.flows.ActNorm(dims), n.flows.InvertibleConv1x1(dims[0]), n.flows.AffineCoupling(dims, masking=masking, odd=i % 2 == 1)]
    layers += [n.flows.ActNorm(dims), n.flows.InvertibleConv1x1(dims[0])]
    comp = n.flows.Compose(layers)
    perturb_(comp, 5)
    for m in comp.modules():
        if isinstance(m, n.flows.ActNorm):
            m.initialized = True
        if isinstance(m, n.flows.ConvNet) and pair:
            m.kernel_flags = L.CONV_PAIR
    comp.to(DEV).eval()
    x = torch.randn((B, ) + dims, device=DEV)
    l0 = torch.randn(B, device=DEV)
    comp.fuse_steps = 2
    comp(x, l0.clone())  # packs weights, builds W
    n0 = n._lib.launch_count()
    z2, l2 = comp(x, l0.clone())
    launches = n._lib.launch_count() - n0
    comp.fuse_steps = 1
    z1, l1 = comp(x, l0.clone())
    comp.fuse_steps = 0
    z0, l0_ = comp(x, l0.clone())
    assert torch.equal(z2, z1), float((z2 - z1).abs().max())
    close(l2, l1, rtol=2e-6, atol=2e-3, what='ldj step-fused vs two kernels')
    close(z2, z0, rtol=1e-5, atol=1e-5, what='z step-fused vs separate layers')
    close(l2, l0_, rtol=2e-6, atol=2e-3, what='ldj step-fused vs separate layers')
    if dims[0] in (3, 12, 48):
        assert launches == 4, launches  # ActNorm+conv | 3 x (coupling [+ next ActNorm+conv])
    assert torch.equal(x, x)  # inputs untouched (Compose works on its own tensors)


# ---------------------------------------------------------------------------------------------------------
# edge cases: ragged / odd shapes take the scalar kernels, B = 1, non-power-of-two widths, big channel counts
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('dims,masking', [((2, 6, 6), 'checkerboard'), ((4, 2, 2), 'channelwise'), ((3, 24, 40), 'checkerboard'),
                                          ((6, 10, 12), 'channelwise'), ((10, ), 'checkerboard'), ((1, 4, 4), 'checkerboard'),
                                          ((5, 8, 8), 'checkerboard')])
@pytest.mark.parametrize('B', [1, 5])
def test_affine_bijection_odd_shapes_vs_oracle(dims, masking, B):
    """The bijection kernel alone (conditioner output given) on shapes that miss the vector paths."""
    import nfb200._lib as L
    mode = L.SPLIT_1D if len(dims) == 1 else (L.SPLIT_CHECKER if masking == 'checkerboard' else L.SPLIT_CHANNEL)
    C, H, W = (dims[0], 1, 1) if len(dims) == 1 else dims
    g = torch.Generator().manual_seed(B)
    z = torch.randn((B, ) + dims, generator=g)
    params = torch.randn((B, ) + dims, generator=g)  # (t | s_raw): D entries per sample
    a, b = torch.tensor([0.4]), torch.tensor([0.02])
    ldj0 = torch.randn(B, generator=g)
    for odd in (False, True):
        split, merge = O.split_fn(len(dims), masking, odd)
        z0, z1 = split(z)
        p = params.reshape((B, 2 * z0.shape[1]) + tuple(z0.shape[2:]))
        for inverse in (False, True):
            o0, lo = O.affine_transform(z0, p, ldj0, a, b, inverse)
            ref = merge(o0, z1)
            zg, out, lg = z.to(DEV), torch.empty_like(z, device=DEV), ldj0.to(DEV).clone()
            pg, ag, bg = p.to(DEV).contiguous(), a.to(DEV), b.to(DEV)  # keep alive until the kernel has run
            fn = L.lib().nfb_affine_coupling_inv if inverse else L.lib().nfb_affine_coupling_fwd
            L.check(fn(zg.data_ptr(), out.data_ptr(), pg.data_ptr(), lg.data_ptr(), lg.data_ptr(), ag.data_ptr(),
                       bg.data_ptr(), B, C, H, W, mode, int(odd), L.stream()))
            torch.cuda.synchronize()
            close(out, ref, what='z %s odd=%s inv=%s' % (dims, odd, inverse))
            close(lg, lo, rtol=1e-5, atol=1e-5, what='ldj')


@pytest.mark.parametrize('shape', [(1, 3, 2, 2), (2, 5, 6, 10), (1, 7, 24, 40), (3, 2, 14, 6)])
def test_simple_layers_odd_shapes_vs_oracle(shape):
    F = nfb().flows
    torch.manual_seed(shape[1])
    B, C = shape[:2]
    x = torch.rand(shape)
    l0 = torch.randn(B)
    z, l = F.Logit(0.01)(x.to(DEV), l0.to(DEV))
    zo, lo = O.logit_fwd(x, l0, 0.01)
    close(z, zo)
    close(l, lo, rtol=1e-5, atol=1e-4)
    an = F.ActNorm(shape[1:])
    perturb_(an, 3)
    an.initialized = True
    zo, lo = O.actnorm_fwd(x, l0, an.log_scale.data, an.bias.data)
    z, l = an.to(DEV)(x.to(DEV), l0.to(DEV).clone())
    close(z, zo)
    close(l, lo, rtol=1e-5, atol=1e-4)
    conv = F.InvertibleConv1x1(C)
    perturb_(conv, 4)
    sd = {k: v.clone() for k, v in conv.state_dict().items()}
    zo, lo = O.invconv_fwd(x, l0, sd['P'], sd['L'], sd['U'], sd['log_s'], sd['sign_s'])
    conv.to(DEV)
    z, l = conv(x.to(DEV), l0.to(DEV).clone())
    close(z, zo)
    close(l, lo, rtol=1e-5, atol=1e-4)
    rows, total = nfb().gauss_nll(z, l)
    y, l2 = conv.backward(z, l.clone())  # backward accumulates into the tensor it is given
    close(y, x, rtol=1e-4, atol=1e-4)
    close(l2, l0, rtol=1e-5, atol=1e-4)
    close(rows, O.nll_rows(zo, lo).float(), rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize('dims,masking,K', [((3, 32, 32), 'checkerboard', 8), ((12, 16, 16), 'channelwise', 8),
                                            ((12, 16, 16), 'checkerboard', 4), ((48, 8, 8), 'channelwise', 8),
                                            ((48, 8, 8), 'checkerboard', 8)])
@pytest.mark.parametrize('B', [3, 32])
def test_flowpp_conditioner_kernel_vs_oracle(dims, masking, K, B):
    """One-kernel Flow++ conditioner (conv -> gated conv -> LN -> gated attention -> LN -> conv) vs the CPU oracle and vs
    the torch-ops path on the device."""
    F = nfb().flows
    from nfb200.flows.squeeze import coupling_split
    torch.manual_seed(K + dims[0])
    cpl = F.MixLogAttnCoupling(dims, masking=masking, odd=False, n_mixtures=K)
    with torch.no_grad():
        for n_, p in cpl.named_parameters():
            if 'pos_emb' in n_:
                p.add_(0.1 * torch.randn(p.shape))
            elif n_.endswith(('2.weight', '4.weight')):  # LayerNorm gains
                p.add_(0.2 * torch.randn(p.shape))
            elif n_.endswith(('2.bias', '4.bias')):
                p.add_(0.1 * torch.randn(p.shape))
    cpl.eval()
    sd = {k: v.clone() for k, v in cpl.state_dict().items()}
    z = torch.randn((B, ) + dims)
    split, _ = O.split_fn(3, masking, False)
    with torch.no_grad():
        ref = O.flowpp_conditioner(sd, 'net.', split(z)[1])
        ref64 = O.flowpp_conditioner(O.to_dtype(sd, torch.float64), 'net.', split(z)[1].double())
    cpl.to(DEV)
    fused = cpl._params(z.to(DEV))
    cpl.fused_conditioner = False
    lib = cpl._params(z.to(DEV))
    scale = max(1.0, float(ref.abs().max()))
    e_f = float((fused.cpu().double() - ref64).abs().max())
    e_r = float((ref.double() - ref64).abs().max())
    e_l = float((lib.cpu().double() - ref64).abs().max())
    print('flow++ conditioner err vs fp64: kernel %.3e | torch-gpu %.3e | cpu fp32 %.3e | scale %.2f' % (e_f, e_l, e_r, scale))
    close(fused, ref, rtol=2e-5, atol=5e-6 * scale, what='fused flow++ conditioner vs oracle')
    assert e_f <= 6.0 * max(e_r, e_l) + 1e-6 * scale


@pytest.mark.parametrize('D,K,B', [(2, 4, 512), (6, 8, 257), (64, 8, 1000), (2, 1, 5)])
@pytest.mark.parametrize('odd', [False, True])
def test_flowpp_conditioner_1d_kernel_vs_oracle(D, K, B, odd):
    """One-kernel Flow++ conditioner of the 1-D couplings (single attention token: A = Q) vs the CPU oracle, and the whole
    1-D Flow++ stack (flowpp.py:64-66: ActNorm + coupling per step) forward / inverse vs the oracle."""
    n = nfb()
    torch.manual_seed(K + D)
    cpl = n.flows.MixLogAttnCoupling((D, ), odd=odd, n_mixtures=K)
    with torch.no_grad():
        for n_, p in cpl.named_parameters():
            if 'pos_emb' in n_:
                p.add_(0.1 * torch.randn(p.shape))
            elif n_.endswith(('2.weight', '4.weight')):
                p.add_(0.2 * torch.randn(p.shape))
            elif n_.endswith(('2.bias', '4.bias')):
                p.add_(0.1 * torch.randn(p.shape))
    cpl.eval()
    sd = {k: v.clone() for k, v in cpl.state_dict().items()}
    z = torch.randn(B, D)
    split, _ = O.split_fn(1, 'checkerboard', odd)
    with torch.no_grad():
        ref = O.flowpp_conditioner(sd, 'net.', split(z)[1])
        ref64 = O.flowpp_conditioner(O.to_dtype(sd, torch.float64), 'net.', split(z)[1].double())
    cpl.to(DEV)
    n0 = n._lib.launch_count()
    fused = cpl._params(z.to(DEV))
    assert n._lib.launch_count() - n0 == 1  # one libnfb200 kernel, no library ops
    scale = max(1.0, float(ref.abs().max()))
    e_f = float((fused.cpu().double() - ref64).abs().max())
    e_r = float((ref.double() - ref64).abs().max())
    close(fused, ref, rtol=2e-5, atol=5e-6 * scale, what='fused 1-D flow++ conditioner vs oracle')
    assert e_f <= 6.0 * e_r + 1e-6 * scale
    if D == 2:  # the 2-D toy densities of the reference (dataset.py:18-60): whole stack
        torch.manual_seed(1)
        net = n.Flowpp((2, ), None, types.SimpleNamespace(layers=4, mixtures=K))
        perturb_(net)
        net.mark_initialized().eval()
        msd = {k: v.clone() for k, v in net.state_dict().items()}
        x = torch.rand(B, 2) * 2 - 1
        spec = oracle_spec('flowpp', (2, ), None, 4, K)
        with torch.no_grad():
            zo, lo = O.stack_forward(spec, msd, x)
        net.to(DEV)
        zg, lg = net(x.to(DEV))
        close(zg, zo, rtol=2e-5, atol=2e-5, what='1-D flow++ stack z')
        close(lg, lo, rtol=2e-5, atol=2e-4, what='1-D flow++ stack ldj')
        y, _ = net.backward(zg)
        close(y, x, rtol=1e-3, atol=1e-3, what='1-D flow++ round trip')


def test_sharded_statistics_match_global_batch():
    """ActNorm init / BatchNorm train statistics from per-shard moments (what the ranks all-reduce) == full-batch pass."""
    n = nfb()
    from nfb200 import parallel
    torch.manual_seed(0)
    x = (torch.randn(10, 12, 4, 4) * 1.7 + 0.3).to(DEV)
    ref = n.flows.ActNorm((12, 4, 4)).to(DEV)
    ref(x, torch.zeros(10, device=DEV))  # single-process data-dependent init
    m = parallel.channel_moments(x[:3].contiguous()) + parallel.channel_moments(x[3:].contiguous())  # = all-reduce(SUM)
    mean, var_b, var_u, cnt = parallel.finalize_moments(m)
    assert float(cnt) == 160.0
    close(torch.log(torch.sqrt(var_u) + 1e-5).float(), ref.log_scale.view(-1), rtol=1e-6, atol=1e-6)
    close(mean.float(), ref.bias.view(-1), rtol=1e-6, atol=1e-6)
    sharded = parallel.actnorm_init_sharded(n.flows.ActNorm((12, 4, 4)).to(DEV), x)  # world size 1: no collective
    close(sharded.log_scale, ref.log_scale, rtol=1e-6, atol=1e-6)
    bn = n.flows.BatchNorm((12, 4, 4), affine=False).to(DEV).train()
    bn(x, torch.zeros(10, device=DEV))
    bn2 = parallel.batchnorm_stats_sharded(n.flows.BatchNorm((12, 4, 4), affine=False).to(DEV), x)
    close(bn2.batch_var, bn.batch_var, rtol=1e-6, atol=1e-6)
    close(bn2.running_mean, bn.running_mean, rtol=1e-6, atol=1e-6)
