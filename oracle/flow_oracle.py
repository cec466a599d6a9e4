"""CPU oracle for the flow forward/inverse + log|det J| hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``normalizing-flows-pytorch_b200/`` may import this
module: it exists so that ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` can check (and time) a CPU restatement of what the
reference computes.  The product path is the CUDA library and has no CPU fallback.

What it restates: tatsy/normalizing-flows-pytorch (``/root/reference``) ``flows/coupling.py``,
``flows/modules.py``, ``flows/squeeze.py``, ``flows/glow.py``, ``flows/flowpp.py``,
``flows/realnvp.py``, ``flows/weight_norm.py`` and the NLL of ``main.py:85``.  The reference's
arithmetic lives in PyTorch's CPU kernels (``environment.yml:14`` asks ``pytorch>=1.6``; here
torch 2.11.0), so the restatement is written over the same ``torch`` CPU primitives (conv2d,
batch_norm, logsumexp ...) -- that keeps the CPU baseline representative of the reference's
own speed -- but it is a *functional* re-derivation: layers are pure functions of an explicit
state dict, the squeeze/checkerboard permutations are written from the index formula
``k = 4c + 2dy + dx`` with strided slices instead of view/permute, and stacks are replayed
from a layer-spec list.

Pinning: the reference ships no tests or golden vectors, so the oracle is
pinned against outputs of the reference itself, generated in the build container by
``tests/golden/make_golden.py`` (imports ``/root/reference``) and committed as
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays them (and, when
``/root/reference`` is mounted, compares live).  The RQ-spline coupling has no reference
implementation at all: **parity unpinned** for that one bijection -- its oracle
is the restatement of Durkan et al. 2019 below, validated by invertibility and an autograd
Jacobian check.

Every function is dtype-generic: pass fp64 tensors/state for a high-precision run.
"""
import math

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# index permutations (squeeze.py)
# --------------------------------------------------------------------------------------


def space_to_depth(z):
    """S[b, 4c+2dy+dx, i, j] = z[b, c, 2i+dy, 2j+dx]   (squeeze.py:36-38 / 90-92)."""
    B, C, H, W = z.shape
    out = z.new_empty(B, 4 * C, H // 2, W // 2)
    for dy in (0, 1):
        for dx in (0, 1):
            out[:, (2 * dy + dx)::4] = z[:, :, dy::2, dx::2]
    return out


def depth_to_space(s):
    """inverse of :func:`space_to_depth` (squeeze.py:59-60 / 109-110)."""
    B, C4, h, w = s.shape
    C = C4 // 4
    out = s.new_empty(B, C, 2 * h, 2 * w)
    for dy in (0, 1):
        for dx in (0, 1):
            out[:, :, dy::2, dx::2] = s[:, (2 * dy + dx)::4]
    return out


def checker_split(z, odd=False):
    """squeeze.py:32-44: blocks a,b,c,d of C squeezed channels; z0 = (a,d), z1 = (b,c)."""
    s = space_to_depth(z)
    C = z.shape[1]
    a, b, c, d = s[:, :C], s[:, C:2 * C], s[:, 2 * C:3 * C], s[:, 3 * C:]
    z0, z1 = torch.cat([a, d], 1), torch.cat([b, c], 1)
    return (z1, z0) if odd else (z0, z1)


def checker_merge(z0, z1, odd=False):
    """squeeze.py:47-61."""
    if odd:
        z0, z1 = z1, z0
    C = z0.shape[1] // 2
    s = torch.cat([z0[:, :C], z1[:, :C], z1[:, C:], z0[:, C:]], 1)
    return depth_to_space(s)


def channel_split(z, odd=False):
    """squeeze.py:5-10."""
    h = z.shape[1] // 2
    z0, z1 = z[:, :h], z[:, h:]
    return (z1, z0) if odd else (z0, z1)


def channel_merge(z0, z1, odd=False):
    """squeeze.py:13-17."""
    return torch.cat([z1, z0] if odd else [z0, z1], 1)


def split1d(z, odd=False):
    """squeeze.py:64-72: even entries / odd entries."""
    z0, z1 = z[:, 0::2], z[:, 1::2]
    return (z1, z0) if odd else (z0, z1)


def merge1d(z0, z1, odd=False):
    """squeeze.py:75-83."""
    if odd:
        z0, z1 = z1, z0
    out = z0.new_empty(z0.shape[0], z0.shape[1] * 2)
    out[:, 0::2] = z0
    out[:, 1::2] = z1
    return out


def squeeze2d_layer(z, odd=False):
    """Squeeze2d.forward (squeeze.py:162-165)."""
    s = space_to_depth(z)
    if odd:
        h = s.shape[1] // 2
        s = torch.cat([s[:, h:], s[:, :h]], 1)
    return s


def unsqueeze2d_layer(s, odd=False):
    """Unsqueeze2d.forward (squeeze.py:181-184)."""
    if odd:
        h = s.shape[1] // 2
        s = torch.cat([s[:, h:], s[:, :h]], 1)
    return depth_to_space(s)


def split_fn(ndim, masking, odd):
    """AbstractCoupling.__init__ dispatch (coupling.py:19-30)."""
    if ndim == 1:
        return (lambda z: split1d(z, odd)), (lambda a, b: merge1d(a, b, odd))
    if ndim == 3 and masking == 'checkerboard':
        return (lambda z: checker_split(z, odd)), (lambda a, b: checker_merge(a, b, odd))
    if ndim == 3 and masking == 'channelwise':
        return (lambda z: channel_split(z, odd)), (lambda a, b: channel_merge(a, b, odd))
    raise ValueError('unsupported combination of masking and dimension')


# --------------------------------------------------------------------------------------
# scalar helpers (modules.py:19-97)
# --------------------------------------------------------------------------------------


def _rowsum(x):
    return x.reshape(x.shape[0], -1).sum(1)


def log_dsigmoid(x):
    """modules.py:19-21."""
    return x - 2.0 * F.softplus(x)


def mixlog_logpdf(x, logpi, mu, s):
    """modules.py:64-67, 76-85.  x (B,*C); logpi/mu/s (B,K,*C)."""
    u = (x.unsqueeze(1) - mu) * torch.exp(-s)
    return torch.logsumexp(logpi + (u - s - 2.0 * F.softplus(u)), dim=1)


def mixlog_logcdf(x, logpi, mu, s):
    """modules.py:70-73, 88-97."""
    u = (x.unsqueeze(1) - mu) * torch.exp(-s)
    return torch.logsumexp(logpi + F.logsigmoid(u), dim=1)


# --------------------------------------------------------------------------------------
# simple bijective layers
# --------------------------------------------------------------------------------------


def logit_fwd(x, ldj, eps):
    """Logit.forward (modules.py:146-150 with 29-32)."""
    x = torch.clamp(x, eps, 1.0 - eps)
    y = torch.logit(torch.clamp(x, 1.0e-8, 1.0 - 1.0e-8))
    return torch.logit(x), ldj + _rowsum(-log_dsigmoid(y))


def logit_inv(x, ldj):
    """Logit.backward (modules.py:152-155)."""
    return torch.sigmoid(x), ldj + _rowsum(log_dsigmoid(x))


def _npix(z):
    return z.numel() // (z.shape[0] * z.shape[1])


def actnorm_stats(z, eps=1.0e-5):
    """ActNorm first-call init (modules.py:238-244): log(unbiased std + eps), mean."""
    r = z.reshape(z.shape[0], z.shape[1], -1)
    return torch.log(torch.std(r, dim=[0, 2]) + eps), torch.mean(r, dim=[0, 2])


def actnorm_fwd(z, ldj, log_scale, bias):
    """ActNorm.forward (modules.py:246-250).  log_scale/bias broadcastable (1,C,1,1)/(1,C)."""
    return (z - bias) / torch.exp(log_scale), ldj - torch.sum(log_scale) * _npix(z)


def actnorm_inv(y, ldj, log_scale, bias):
    """ActNorm.backward (modules.py:252-256)."""
    return y * torch.exp(log_scale) + bias, ldj + torch.sum(log_scale) * _npix(y)


def bnflow_batch_stats(x, eps=1.0e-5):
    """flow BatchNorm train-mode statistics (modules.py:285-287): mean, biased var + eps."""
    r = x.reshape(x.shape[0], x.shape[1], -1)
    m = torch.mean(r, dim=[0, 2], keepdim=True)
    v = torch.mean((r - m).pow(2), dim=[0, 2], keepdim=True) + eps
    return m.reshape(-1), v.reshape(-1)


def bnflow_fwd(x, ldj, mean, var, log_gamma, beta):
    """flow BatchNorm.forward normalisation part (modules.py:300-305)."""
    y = (x - mean) / torch.sqrt(var)
    y = y * torch.exp(log_gamma) + beta
    return y, ldj + torch.sum(log_gamma - 0.5 * torch.log(var)) * _npix(x)


def bnflow_inv(x, ldj, mean, var, log_gamma, beta):
    """flow BatchNorm.backward (modules.py:315-320)."""
    y = (x - beta) / torch.exp(log_gamma)
    y = y * torch.sqrt(var) + mean
    return y, ldj + torch.sum(-log_gamma + 0.5 * torch.log(var)) * _npix(x)


def invconv_weight(P, L, U, log_s, sign_s):
    """W = P (L o tril_-1 + I) (U o triu_1 + diag(sign_s exp(log_s)))   (modules.py:471-473)."""
    C = L.shape[0]
    eye = torch.eye(C, dtype=L.dtype)
    Lm = torch.tril(L, -1) + eye
    Um = torch.triu(U, 1) + torch.diag(sign_s * torch.exp(log_s))
    return P @ Lm @ Um


def invconv_fwd(z, ldj, P, L, U, log_s, sign_s):
    """InvertibleConv1x1.forward (modules.py:470-482)."""
    W = invconv_weight(P, L, U, log_s, sign_s)
    B, C = z.shape[:2]
    out = torch.matmul(W, z.reshape(B, C, -1)).reshape(z.shape)
    return out, ldj + torch.sum(log_s) * _npix(z)


def invconv_inv(y, ldj, P, L, U, log_s, sign_s):
    """InvertibleConv1x1.backward (modules.py:484-497): solve W x = y.

    The reference uses ``lu_solve`` with its stored LU/pivots; mathematically that is
    x = U^-1 L^-1 P^T y, restated here with two triangular solves.
    """
    C = L.shape[0]
    eye = torch.eye(C, dtype=L.dtype)
    Lm = torch.tril(L, -1) + eye
    Um = torch.triu(U, 1) + torch.diag(sign_s * torch.exp(log_s))
    B = y.shape[0]
    r = y.reshape(B, C, -1)
    r = torch.matmul(P.t(), r)
    r = torch.linalg.solve_triangular(Lm.expand(B, C, C), r, upper=False, unitriangular=True)
    r = torch.linalg.solve_triangular(Um.expand(B, C, C), r, upper=True)
    return r.reshape(y.shape).contiguous(), ldj - torch.sum(log_s) * _npix(y)


# --------------------------------------------------------------------------------------
# conditioners (modules.py:342-438, 500-578; weight_norm.py)
# --------------------------------------------------------------------------------------


def wn_weight(v, g, eps=1.0e-5):
    """weight_norm.py:40: w = v * g / (||v||_{dim 0} + eps)."""
    return v * (g / (torch.norm(v, dim=0) + eps))


def _bn_eval(sd, key, x, train=False):
    """nn.BatchNorm1d/2d inside the conditioners: running statistics (eval) or batch statistics (train mode,
    modules.py:349-352 under net.train(); functional -- the running buffers are left alone)."""
    if train:
        return F.batch_norm(x, None, None, sd[key + '.weight'], sd[key + '.bias'], True, 0.1, 1.0e-5)
    return F.batch_norm(x, sd[key + '.running_mean'], sd[key + '.running_var'], sd[key + '.weight'],
                        sd[key + '.bias'], False, 0.1, 1.0e-5)


def _wn_layer(sd, key, x, pad):
    w = wn_weight(sd[key + '.module.weight_v'], sd[key + '.module.weight_g'])
    b = sd[key + '.module.bias']
    if w.dim() == 4:
        return F.conv2d(x, w, b, 1, pad)
    return F.linear(x, w, b)


def resnet_conditioner(sd, prefix, x, train=False):
    """ConvNet.forward / MLP.forward (modules.py:391-438, 342-388); eval mode unless ``train``.

    in_block.0 (weight-normed conv3x3 or linear) -> 2 x [BN, ReLU, WN, BN, ReLU, WN] + skip
    -> BN, ReLU, weight-normed conv1x1 / linear.  The same key layout serves both.
    """
    x = _wn_layer(sd, prefix + 'in_block.0', x, 1)
    i = 0
    while (prefix + 'mid_block.%d.net.0.running_mean' % i) in sd:
        p = prefix + 'mid_block.%d.net.' % i
        y = F.relu(_bn_eval(sd, p + '0', x, train))
        y = _wn_layer(sd, p + '2', y, 1)
        y = F.relu(_bn_eval(sd, p + '3', y, train))
        y = _wn_layer(sd, p + '5', y, 1)
        x = x + y
        i += 1
    x = F.relu(_bn_eval(sd, prefix + 'out_block.0', x, train))
    return _wn_layer(sd, prefix + 'out_block.2', x, 0)


def _gated(sd, key, x):
    """GatedConv2d / GatedLinear (modules.py:500-535)."""
    C = x.shape[1]
    y = F.elu(torch.cat([x, -x], 1))
    w, b = sd[key + '.op.weight'], sd[key + '.op.bias']
    y = F.conv2d(y, w, b, 1, 1) if w.dim() == 4 else F.linear(y, w, b)
    y = F.elu(torch.cat([y, -y], 1))
    return x + y[:, :C] * torch.sigmoid(y[:, C:])


def _gated_attn(sd, key, x, heads=4):
    """GatedAttn.forward (modules.py:556-578); note the (V, K, Q) split order of line 566."""
    shape = x.shape
    B, C = shape[:2]
    w1, b1 = sd[key + '.conv1.weight'], sd[key + '.conv1.bias']
    w2, b2 = sd[key + '.conv2.weight'], sd[key + '.conv2.bias']
    filters = w1.shape[0] // 3
    D = filters // heads
    xr = (x + sd[key + '.pos_emb']).reshape(B, C, -1)
    p = F.conv1d(xr, w1, b1).reshape(B, 3 * heads, D, -1)
    V, K, Q = p[:, :heads], p[:, heads:2 * heads], p[:, 2 * heads:]
    Wm = torch.matmul(V.transpose(2, 3), K) / math.sqrt(D)
    Wm = F.softmax(Wm, dim=2)
    A = torch.matmul(Q, Wm).reshape(B, C, -1)
    y = F.conv1d(A, w2, b2)
    y = y[:, :C] * torch.sigmoid(y[:, C:])
    return x + y.reshape(shape)


def flowpp_conditioner(sd, prefix, x):
    """MixLogAttnCoupling.net (coupling.py:142-167)."""
    w0, b0 = sd[prefix + '0.weight'], sd[prefix + '0.bias']
    y = F.conv2d(x, w0, b0, 1, 1) if w0.dim() == 4 else F.linear(x, w0, b0)
    y = _gated(sd, prefix + '1', y)
    ln = sd[prefix + '2.weight']
    y = F.layer_norm(y, ln.shape, ln, sd[prefix + '2.bias'], 1.0e-5)
    y = _gated_attn(sd, prefix + '3', y)
    ln = sd[prefix + '4.weight']
    y = F.layer_norm(y, ln.shape, ln, sd[prefix + '4.bias'], 1.0e-5)
    w5, b5 = sd[prefix + '5.weight'], sd[prefix + '5.bias']
    return F.conv2d(y, w5, b5, 1, 1) if w5.dim() == 4 else F.linear(y, w5, b5)


# --------------------------------------------------------------------------------------
# coupling bijections (given the conditioner output ``params``)
# --------------------------------------------------------------------------------------


def affine_transform(z0, params, ldj, s_log_scale, s_bias, inverse=False):
    """AffineCoupling._transform / _inverse_transform (coupling.py:104-122)."""
    n = z0.shape[1]
    t = params[:, :n]
    s = torch.tanh(params[:, n:]) * s_log_scale + s_bias
    if not inverse:
        return z0 * torch.exp(s) + t, ldj + _rowsum(s)
    return torch.exp(-s) * (z0 - t), ldj - _rowsum(s)


def additive_transform(z0, t, ldj, inverse=False):
    """AdditiveCoupling._transform / _inverse_transform (coupling.py:69-79): z0 +- net_t(z1); no log-det."""
    return (z0 - t if inverse else z0 + t), ldj


def mixlogcdf_fwd(x, log_pi, mu, s, ldj):
    """MixLogCDF.forward (modules.py:190-194); log_pi / mu / s are (B, K, *x.shape[1:])."""
    return torch.exp(mixlog_logcdf(x, log_pi, mu, s)), ldj + _rowsum(mixlog_logpdf(x, log_pi, mu, s))


def mixlogcdf_inv(y, log_pi, mu, s, ldj):
    """MixLogCDF.backward (modules.py:196-212): bisection, then ldj - sum log-pdf at the root."""
    x, _ = mixlog_bisect(y, log_pi, mu, s)
    return x, ldj - _rowsum(mixlog_logpdf(x, log_pi, mu, s))


def _mix_params(z0, params, K, a_log_scale, a_bias):
    B, c0 = z0.shape[:2]
    tail = z0.shape[2:]
    a = torch.tanh(params[:, :c0]) * a_log_scale + a_bias
    b = params[:, c0:2 * c0]
    o = 2 * c0
    logpi = F.log_softmax(params[:, o:o + K * c0].reshape(B, K, c0, *tail), dim=1)
    mu = params[:, o + K * c0:o + 2 * K * c0].reshape(B, K, c0, *tail)
    s = params[:, o + 2 * K * c0:o + 3 * K * c0].reshape(B, K, c0, *tail)
    return a, b, logpi, mu, s


def mixlog_transform(z0, params, ldj, K, a_log_scale, a_bias):
    """MixLogAttnCoupling._transform (coupling.py:172-190), MixLogCDF.forward (modules.py:190-194)."""
    a, b, logpi, mu, s = _mix_params(z0, params, K, a_log_scale, a_bias)
    ldj = ldj + _rowsum(mixlog_logpdf(z0, logpi, mu, s))
    x = torch.exp(mixlog_logcdf(z0, logpi, mu, s))
    x, ldj = logit_fwd(x, ldj, 1.0e-5)
    return x * torch.exp(a) + b, ldj + _rowsum(a)


def mixlog_bisect(x, logpi, mu, s, max_iter=100, tol=1.0e-4):
    """MixLogCDF.backward search (modules.py:197-208), including its *global* stop rule."""
    lo = torch.full_like(x, -1.0e3)
    hi = torch.full_like(x, 1.0e3)
    n_iter = 0
    for _ in range(max_iter):
        mid = (lo + hi) * 0.5
        val = torch.exp(mixlog_logcdf(mid, logpi, mu, s))
        lo = torch.where(val < x, mid, lo)
        hi = torch.where(val > x, mid, hi)
        n_iter += 1
        if bool(torch.all(torch.abs(hi - lo) < tol)):
            break
    return (lo + hi) * 0.5, n_iter


def mixlog_inverse(z0, params, ldj, K, a_log_scale, a_bias):
    """MixLogAttnCoupling._inverse_transform (coupling.py:192-210)."""
    a, b, logpi, mu, s = _mix_params(z0, params, K, a_log_scale, a_bias)
    x = torch.exp(-a) * (z0 - b)
    ldj = ldj - _rowsum(a)
    x, ldj = logit_inv(x, ldj)
    x, _ = mixlog_bisect(x, logpi, mu, s)
    return x, ldj - _rowsum(mixlog_logpdf(x, logpi, mu, s))


# ---- rational-quadratic spline (Durkan et al. 2019; NOT in the reference: parity unpinned) ----

RQS_MIN_W = 1.0e-3
RQS_MIN_H = 1.0e-3
RQS_MIN_D = 1.0e-3


def _rqs_knots(params, z0, K, bound):
    """params (B, (3K-1) c0, *tail): sections [K c0 widths | K c0 heights | (K-1) c0 derivatives],
    each viewed (B, K, c0, *tail) -- bin-major like the reference's mixture layout
    (coupling.py:180-182).  Returns knot x, knot y (B,K+1,...) and knot derivatives (B,K+1,...)."""
    B, c0 = z0.shape[:2]
    tail = z0.shape[2:]
    uw = params[:, :K * c0].reshape(B, K, c0, *tail)
    uh = params[:, K * c0:2 * K * c0].reshape(B, K, c0, *tail)
    ud = params[:, 2 * K * c0:].reshape(B, K - 1, c0, *tail)

    def knots(u, mn):
        p = mn + (1.0 - mn * K) * F.softmax(u, dim=1)
        c = torch.cumsum(p, dim=1)
        c = torch.cat([torch.zeros_like(c[:, :1]), c], 1)
        c = 2.0 * bound * c - bound
        c[:, 0] = -bound
        c[:, K] = bound
        return c

    xk, yk = knots(uw, RQS_MIN_W), knots(uh, RQS_MIN_H)
    one = torch.ones_like(xk[:, :1])
    dk = torch.cat([one, RQS_MIN_D + F.softplus(ud), one], 1)
    return xk, yk, dk


def _rqs_gather(t, idx):
    return torch.gather(t, 1, idx.unsqueeze(1)).squeeze(1)


def rqs_transform(z0, params, ldj, K=8, bound=3.0, inverse=False):
    """Monotonic RQ spline with linear (identity) tails, elementwise; log-det summed per sample."""
    xk, yk, dk = _rqs_knots(params, z0, K, bound)
    inside = (z0 >= -bound) & (z0 <= bound)
    x = torch.clamp(z0, -bound, bound)
    edges = yk if inverse else xk
    idx = (x.unsqueeze(1) >= edges[:, 1:K]).sum(1)  # bin index in [0, K-1]
    x0, x1 = _rqs_gather(xk, idx), _rqs_gather(xk, idx + 1)
    y0, y1 = _rqs_gather(yk, idx), _rqs_gather(yk, idx + 1)
    d0, d1 = _rqs_gather(dk, idx), _rqs_gather(dk, idx + 1)
    w, h = x1 - x0, y1 - y0
    sk = h / w
    if not inverse:
        xi = (x - x0) / w
        om = xi * (1.0 - xi)
        den = sk + (d1 + d0 - 2.0 * sk) * om
        out = y0 + h * (sk * xi * xi + d0 * om) / den
        ld = 2.0 * torch.log(sk) + torch.log(d1 * xi * xi + 2.0 * sk * om + d0 * (1.0 - xi) * (1.0 - xi)) \
            - 2.0 * torch.log(den)
    else:
        dy = x - y0
        t = d0 + d1 - 2.0 * sk
        a = dy * t + h * (sk - d0)
        b = h * d0 - dy * t
        c = -sk * dy
        disc = b * b - 4.0 * a * c
        xi = 2.0 * c / (-b - torch.sqrt(disc))
        out = xi * w + x0
        om = xi * (1.0 - xi)
        den = sk + t * om
        ld = -(2.0 * torch.log(sk) + torch.log(d1 * xi * xi + 2.0 * sk * om + d0 * (1.0 - xi) * (1.0 - xi))
               - 2.0 * torch.log(den))
    out = torch.where(inside, out, z0)
    ld = torch.where(inside, ld, torch.zeros_like(ld))
    return out, ldj + _rowsum(ld)


# --------------------------------------------------------------------------------------
# stacks (glow.py:17-60, flowpp.py:17-70, realnvp.py:17-55) as layer-spec lists
# --------------------------------------------------------------------------------------


def stack_spec(model, dims, datatype, layers, mixtures=4, coupling=None):
    """Layer list of the reference's Glow / Flowpp / RealNVP constructors.

    Each entry is (kind, options); entry i owns state-dict prefix ``net.layers.<i>.``.
    ``coupling`` overrides the coupling kind ('affine' | 'mixlog' | 'rqs').
    """
    dims = tuple(dims)
    cpl = coupling or {'glow': 'affine', 'realnvp': 'affine', 'flowpp': 'mixlog'}[model]
    spec = []

    def step(d, masking, odd):
        if model == 'realnvp':
            spec.append(('bnflow', dict(dims=d)))
        else:
            spec.append(('actnorm', dict(dims=d)))
            if not (model == 'flowpp' and datatype != 'image'):  # flowpp.py:65-66: no 1x1 conv
                spec.append(('invconv', dict(dims=d)))
        spec.append((cpl, dict(dims=d, masking=masking, odd=odd, mixtures=mixtures)))

    if datatype == 'image':
        spec.append(('logit', dict(eps=0.01)))
        d = dims
        n_sq = 0
        while max(d[1], d[2]) > 8:
            for i in range(layers):
                step(d, 'checkerboard', i % 2 != 0)
            spec.append(('squeeze2d', {}))
            n_sq += 1
            d = (d[0] * 4, d[1] // 2, d[2] // 2)
            for i in range(layers):
                step(d, 'channelwise', i % 2 != 0)
        for i in range(layers + 1):
            step(d, 'checkerboard', i % 2 != 0)
        for _ in range(n_sq):
            spec.append(('unsqueeze2d', {}))
    else:
        for i in range(layers):
            step(dims, 'checkerboard', i % 2 != 0)
    return spec


def _coupling(kind, opt, sd, p, z, ldj, inverse, train=False):
    d = opt['dims']
    split, merge = split_fn(len(d), opt['masking'], opt['odd'])
    z0, z1 = split(z)
    if kind == 'mixlog':
        params = flowpp_conditioner(sd, p + 'net.', z1)
        f = mixlog_inverse if inverse else mixlog_transform
        z0, ldj = f(z0, params, ldj, opt['mixtures'], sd[p + 'a_log_scale'], sd[p + 'a_bias'])
    elif kind == 'additive':
        z0, ldj = additive_transform(z0, resnet_conditioner(sd, p + 'net_t.', z1, train), ldj, inverse)
    else:
        params = resnet_conditioner(sd, p + 'net.', z1, train)
        if kind == 'affine':
            z0, ldj = affine_transform(z0, params, ldj, sd[p + 's_log_scale'], sd[p + 's_bias'], inverse)
        else:
            z0, ldj = rqs_transform(z0, params, ldj, opt.get('bins', 8), opt.get('bound', 3.0), inverse)
    return merge(z0, z1), ldj


def _apply(kind, opt, sd, p, z, ldj, inverse, train=False):
    if kind == 'logit':
        return logit_inv(z, ldj) if inverse else logit_fwd(z, ldj, opt['eps'])
    if kind == 'actnorm':
        f = actnorm_inv if inverse else actnorm_fwd
        return f(z, ldj, sd[p + 'log_scale'], sd[p + 'bias'])
    if kind == 'bnflow':
        f = bnflow_inv if inverse else bnflow_fwd
        if train and not inverse:  # modules.py:284-296: batch statistics through buffers (.data.copy_): no gradient
            m, v = bnflow_batch_stats(z.detach())
            shape = sd[p + 'log_gamma'].shape
            return f(z, ldj, m.reshape(shape), v.reshape(shape), sd[p + 'log_gamma'], sd[p + 'beta'])
        return f(z, ldj, sd[p + 'running_mean'], sd[p + 'running_var'], sd[p + 'log_gamma'], sd[p + 'beta'])
    if kind == 'invconv':
        f = invconv_inv if inverse else invconv_fwd
        return f(z, ldj, sd[p + 'P'], sd[p + 'L'], sd[p + 'U'], sd[p + 'log_s'], sd[p + 'sign_s'])
    if kind == 'squeeze2d':
        return (unsqueeze2d_layer(z) if inverse else squeeze2d_layer(z)), ldj
    if kind == 'unsqueeze2d':
        return (squeeze2d_layer(z) if inverse else unsqueeze2d_layer(z)), ldj
    return _coupling(kind, opt, sd, p, z, ldj, inverse, train)


def stack_forward(spec, sd, x, train=False):
    """Model.forward -> Compose.forward (glow.py:62-64, modules.py:331-334); eval mode unless ``train`` (batch
    statistics in every BatchNorm, as under ``net.train()`` in main.py:71-82).  Differentiable by torch autograd: with
    leaf tensors in ``sd`` that require grad this is the gradient oracle of the training step."""
    z, ldj = x, torch.zeros(x.shape[0], dtype=x.dtype)
    for i, (kind, opt) in enumerate(spec):
        z, ldj = _apply(kind, opt, sd, 'net.layers.%d.' % i, z, ldj, False, train)
    return z, ldj


def mean_nll(z, ldj):
    """main.py:85: loss = -mean(log N(z; 0, I) + ldj), in the dtype of z (differentiable)."""
    zf = z.reshape(z.shape[0], -1)
    logp = -0.5 * (zf * zf).sum(1) - 0.5 * zf.shape[1] * math.log(2.0 * math.pi)
    return -1.0 * torch.mean(logp + ldj)


def stack_backward(spec, sd, z):
    """Model.backward -> Compose.backward (glow.py:66-68, modules.py:336-339)."""
    y, ldj = z, torch.zeros(z.shape[0], dtype=z.dtype)
    for i in reversed(range(len(spec))):
        kind, opt = spec[i]
        y, ldj = _apply(kind, opt, sd, 'net.layers.%d.' % i, y, ldj, True)
    return y, ldj


# --------------------------------------------------------------------------------------
# likelihood (main.py:49-51, 83-85)
# --------------------------------------------------------------------------------------


def nll_rows(z, ldj):
    """-(log N(z; 0, I) + ldj) per sample, accumulated in fp64."""
    zf = z.reshape(z.shape[0], -1).double()
    D = zf.shape[1]
    return 0.5 * (zf * zf).sum(1) + 0.5 * D * math.log(2.0 * math.pi) - ldj.double()


def bits_per_dim(z, ldj):
    """mean NLL / (D ln 2); the reference itself never reports it."""
    D = z[0].numel()
    return float(nll_rows(z, ldj).mean() / (D * math.log(2.0)))


def to_dtype(sd, dtype):
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
