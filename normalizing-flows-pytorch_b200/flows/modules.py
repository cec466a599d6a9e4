"""Bijective layers of the reference's ``flows/modules.py`` on libnfb200 (CUDA, sm_100a).

Same class names, constructor signatures, ``forward(z, log_df_dz)`` / ``backward(y, log_df_dz)`` methods and
``state_dict`` keys as the reference, so reference checkpoints load unchanged.  ``inverse`` is an
alias of ``backward``.  ``forward`` is differentiable (``flows/autograd.py``: one gradient kernel per layer) whenever
gradients are enabled and an input or parameter requires them; the inverse direction is inference-only, like the
reference's sampling path (main.py:121-124 runs it under no_grad).
"""
import numpy as np
import torch
import torch.nn as nn

from .. import _lib as L
from . import autograd as G


def _world():
    """Number of ranks of the default process group (1 without one)."""
    import torch.distributed as dist
    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


# Cross-sample statistics (ActNorm's data-dependent init, the flow BatchNorm's train-mode batch statistics) are taken over
# the GLOBAL batch when a process group is initialised: every rank holds a shard, the per-channel moments are all-reduced
# (parallel.actnorm_init_sharded / batchnorm_stats_sharded), and all replicas end up with identical values -- as if the
# reference had seen the whole batch on one device.  Set to False for rank-local statistics.
SYNC_STATS = True


def _bchw(z):
    if z.dim() == 2:
        return z.size(0), z.size(1), 1
    return z.size(0), z.size(1), z[0, 0].numel()


class Logit(nn.Module):
    """modules.py:141-155.  Returns a NEW log-det tensor like the reference."""

    def __init__(self, eps=1.0e-5):
        super().__init__()
        self.eps = eps

    def forward(self, x, log_df_dz):
        # torch.clamp(x, eps, 1.0 - eps) rounds the python doubles to fp32 (modules.py:147)
        lo, hi = float(np.float32(self.eps)), float(np.float32(1.0 - self.eps))
        if G.needs_grad(x, log_df_dz):
            return G.LogitFn.apply(x, log_df_dz, lo, hi)
        x, log_df_dz = L.dev(x, 'x'), L.dev(log_df_dz, 'log_df_dz')
        out, ldj = torch.empty_like(x), torch.empty_like(log_df_dz)
        B = x.size(0)
        L.check(L.lib().nfb_logit_fwd(L.ptr(x), L.ptr(out), L.ptr(log_df_dz), L.ptr(ldj), lo, hi, B, x[0].numel(),
                                      L.stream()))
        return out, ldj

    def backward(self, x, log_df_dz):
        x, log_df_dz = L.dev(x, 'x'), L.dev(log_df_dz, 'log_df_dz')
        out, ldj = torch.empty_like(x), torch.empty_like(log_df_dz)
        L.check(L.lib().nfb_logit_inv(L.ptr(x), L.ptr(out), L.ptr(log_df_dz), L.ptr(ldj), x.size(0), x[0].numel(),
                                      L.stream()))
        return out, ldj

    inverse = backward


class MixLogCDF(nn.Module):
    """modules.py:186-212: the logistic-mixture CDF as a bijection of x given (log_pi, mu, s) of shape (B, K, *x.shape[1:]);
    `backward` is the reference's bisection.  Returns NEW log-det tensors like the reference.  (MixLogAttnCoupling runs the
    same arithmetic fused with its logit / affine stages, coupling.py:183-189.)"""

    def __init__(self):
        super().__init__()
        self._scratch = self._flag = None

    @staticmethod
    def _args(x, log_pi, mu, s, log_df_dz):
        x, log_df_dz = L.dev(x, 'x'), L.dev(log_df_dz, 'log_df_dz')
        log_pi, mu, s = L.dev(log_pi, 'log_pi'), L.dev(mu, 'mu'), L.dev(s, 's')
        B, K = log_pi.size(0), log_pi.size(1)
        n = x[0].numel()
        if not (log_pi.shape == mu.shape == s.shape) or log_pi[0, 0].numel() != n or x.size(0) != B:
            raise RuntimeError('nfb200: MixLogCDF expects x (B, ...) and log_pi / mu / s (B, K, ...) of matching shapes')
        return x, log_pi, mu, s, log_df_dz, B, n, K

    def forward(self, x, log_pi, mu, s, log_df_dz):
        x, log_pi, mu, s, log_df_dz, B, n, K = self._args(x, log_pi, mu, s, log_df_dz)
        y, ldj = torch.empty_like(x), torch.empty_like(log_df_dz)
        L.check(L.lib().nfb_mixlogcdf_fwd(L.ptr(x), L.ptr(y), L.ptr(log_pi), L.ptr(mu), L.ptr(s), L.ptr(log_df_dz),
                                          L.ptr(ldj), B, n, K, L.stream()))
        return y, ldj

    def backward(self, x, log_pi, mu, s, log_df_dz):
        x, log_pi, mu, s, log_df_dz, B, n, K = self._args(x, log_pi, mu, s, log_df_dz)
        if self._scratch is None or self._scratch.numel() < 2 * B * n or self._scratch.device != x.device:
            self._scratch = torch.empty(2 * B * n, device=x.device, dtype=torch.float32)
            self._flag = torch.zeros(1, device=x.device, dtype=torch.int32)
        out, ldj = torch.empty_like(x), torch.empty_like(log_df_dz)
        L.check(L.lib().nfb_mixlogcdf_inv(L.ptr(x), L.ptr(out), L.ptr(log_pi), L.ptr(mu), L.ptr(s), L.ptr(log_df_dz),
                                          L.ptr(ldj), L.ptr(self._scratch), L.ptr(self._flag), B, n, K, L.stream()))
        return out, ldj

    inverse = backward


class ActNorm(nn.Module):
    """modules.py:225-256 (data-dependent init on the first call; ``initialized`` is a plain attribute)."""

    def __init__(self, num_features, eps=1.0e-5):
        super().__init__()
        self.num_features = num_features
        self.eps = eps
        self.dimensions = [1] + [1 for _ in num_features]
        self.dimensions[1] = num_features[0]
        self.log_scale = nn.Parameter(torch.zeros(self.dimensions))
        self.bias = nn.Parameter(torch.zeros(self.dimensions))
        self.initialized = False

    def forward(self, z, log_df_dz):
        z, log_df_dz = L.dev(z, 'z'), L.dev(log_df_dz, 'log_df_dz')
        B, C, HW = _bchw(z)
        if not self.initialized:
            if SYNC_STATS and _world() > 1:
                from ..parallel import actnorm_init_sharded
                actnorm_init_sharded(self, z)  # moments of the global batch (one all-reduce of 2C+1 doubles)
            else:
                L.check(L.lib().nfb_actnorm_init(L.ptr(z), L.ptr(self.log_scale.data), L.ptr(self.bias.data), B, C, HW,
                                                 float(self.eps), L.stream()))
            self.initialized = True
        if G.needs_grad(z, log_df_dz, self.log_scale, self.bias):
            return G.ActNormFn.apply(z, log_df_dz, self.log_scale, self.bias)
        out = torch.empty_like(z)
        L.check(L.lib().nfb_actnorm_fwd(L.ptr(z), L.ptr(out), L.ptr(log_df_dz), L.ptr(log_df_dz),
                                        L.ptr(self.log_scale.data), L.ptr(self.bias.data), B, C, HW, L.stream()))
        return out, log_df_dz

    def backward(self, y, log_df_dz):
        y, log_df_dz = L.dev(y, 'y'), L.dev(log_df_dz, 'log_df_dz')
        B, C, HW = _bchw(y)
        out = torch.empty_like(y)
        L.check(L.lib().nfb_actnorm_inv(L.ptr(y), L.ptr(out), L.ptr(log_df_dz), L.ptr(log_df_dz),
                                        L.ptr(self.log_scale.data), L.ptr(self.bias.data), B, C, HW, L.stream()))
        return out, log_df_dz

    inverse = backward


class BatchNorm(nn.Module):
    """flow BatchNorm, modules.py:259-322 (train mode: batch statistics + running-stat update)."""

    def __init__(self, num_features, momentum=0.1, eps=1.0e-5, affine=True):
        super().__init__()
        self.num_features = num_features
        self.eps = eps
        self.momentum = momentum
        self.dimensions = [1] + [1 for _ in num_features]
        self.dimensions[1] = num_features[0]
        if affine:
            self.log_gamma = nn.Parameter(torch.zeros(self.dimensions))
            self.beta = nn.Parameter(torch.zeros(self.dimensions))
        else:
            self.register_buffer('log_gamma', torch.zeros(self.dimensions))
            self.register_buffer('beta', torch.zeros(self.dimensions))
        self.register_buffer('running_mean', torch.zeros(self.dimensions))
        self.register_buffer('running_var', torch.ones(self.dimensions))
        self.register_buffer('batch_mean', torch.zeros(self.dimensions))
        self.register_buffer('batch_var', torch.ones(self.dimensions))

    def _stats(self):
        if self.training:
            return self.batch_mean, self.batch_var
        return self.running_mean, self.running_var

    def forward(self, x, log_det_jacob):
        x, log_det_jacob = L.dev(x, 'x'), L.dev(log_det_jacob, 'log_det_jacob')
        B, C, HW = _bchw(x)
        if self.training and SYNC_STATS and _world() > 1:
            from ..parallel import batchnorm_stats_sharded
            batchnorm_stats_sharded(self, x)  # batch statistics of the global batch + running-statistic update
        elif self.training:
            L.check(L.lib().nfb_bnflow_batch_stats(L.ptr(x), L.ptr(self.batch_mean), L.ptr(self.batch_var), B, C, HW,
                                                   float(self.eps), L.stream()))
            with torch.no_grad():  # modules.py:291-294 (C-element bookkeeping)
                self.running_mean.mul_(1.0 - self.momentum).add_(self.batch_mean * self.momentum)
                self.running_var.mul_(1.0 - self.momentum).add_(self.batch_var * self.momentum)
        mean, var = self._stats()
        if G.needs_grad(x, log_det_jacob, self.log_gamma, self.beta):
            return G.BatchNormFlowFn.apply(x, log_det_jacob, mean, var, self.log_gamma, self.beta)
        out = torch.empty_like(x)
        L.check(L.lib().nfb_bnflow_fwd(L.ptr(x), L.ptr(out), L.ptr(log_det_jacob), L.ptr(log_det_jacob), L.ptr(mean),
                                       L.ptr(var), L.ptr(self.log_gamma.data), L.ptr(self.beta.data), B, C, HW,
                                       L.stream()))
        return out, log_det_jacob

    def backward(self, x, log_det_jacob):
        x, log_det_jacob = L.dev(x, 'x'), L.dev(log_det_jacob, 'log_det_jacob')
        B, C, HW = _bchw(x)
        mean, var = self._stats()
        out = torch.empty_like(x)
        L.check(L.lib().nfb_bnflow_inv(L.ptr(x), L.ptr(out), L.ptr(log_det_jacob), L.ptr(log_det_jacob), L.ptr(mean),
                                       L.ptr(var), L.ptr(self.log_gamma.data), L.ptr(self.beta.data), B, C, HW,
                                       L.stream()))
        return out, log_det_jacob

    inverse = backward


class InvertibleConv1x1(nn.Module):
    """modules.py:441-497.  W = P L U is assembled (and inverted, fp64 on the device) by one small kernel and
    cached until L / U / log_s change, instead of two matmuls per forward and an lu_solve per inverse."""

    def __init__(self, in_out_channels):
        super().__init__()
        C = in_out_channels
        W = torch.zeros((C, C), dtype=torch.float32)
        nn.init.orthogonal_(W)
        LU, pivots = torch.linalg.lu_factor(W)
        P, Lo, Up = torch.lu_unpack(LU, pivots)
        self.P = nn.Parameter(P, requires_grad=False)
        self.L = nn.Parameter(Lo, requires_grad=True)
        self.U = nn.Parameter(Up, requires_grad=True)
        self.I = nn.Parameter(torch.eye(C), requires_grad=False)
        self.pivots = nn.Parameter(pivots, requires_grad=False)
        self.L_mask = nn.Parameter(torch.tril(torch.ones(C, C), -1), requires_grad=False)
        self.U_mask = nn.Parameter(torch.triu(torch.ones(C, C), 1), requires_grad=False)
        s = torch.diag(Up)
        self.log_s = nn.Parameter(torch.log(torch.abs(s)), requires_grad=True)
        self.sign_s = nn.Parameter(torch.sign(s), requires_grad=False)
        self._cache_key = None
        self._W = None
        self._Winv = None

    def matrices(self):
        """(W, W^-1) device tensors, rebuilt only when a parameter changed."""
        ps = (self.P, self.L, self.U, self.log_s, self.sign_s)
        key = L.param_key(ps)
        if key != self._cache_key:
            C = self.L.size(0)
            dev = self.L.device
            if self._W is None or self._W.device != dev:
                self._W = torch.empty((C, C), device=dev, dtype=torch.float32)
                self._Winv = torch.empty((C, C), device=dev, dtype=torch.float32)
            for p in ps:
                L.dev(p.data, 'InvertibleConv1x1 parameter')
            L.check(L.lib().nfb_invconv1x1_weight(L.ptr(self.P.data), L.ptr(self.L.data), L.ptr(self.U.data),
                                                  L.ptr(self.log_s.data), L.ptr(self.sign_s.data), L.ptr(self._W),
                                                  L.ptr(self._Winv), C, L.stream()))
            self._cache_key = key
        return self._W, self._Winv

    def _apply_matrix(self, z, log_df_dz, M, sign):
        z, log_df_dz = L.dev(z, 'z'), L.dev(log_df_dz, 'log_df_dz')
        B, C, HW = _bchw(z)
        out = torch.empty_like(z)
        L.check(L.lib().nfb_invconv1x1_apply(L.ptr(z), L.ptr(out), L.ptr(log_df_dz), L.ptr(log_df_dz), L.ptr(M),
                                             L.ptr(self.log_s.data), float(sign), B, C, HW, L.stream()))
        return out, log_df_dz

    def forward(self, z, log_df_dz):
        if G.needs_grad(z, log_df_dz, self.L, self.U, self.log_s):
            return G.InvConv1x1Fn.apply(z, log_df_dz, self.L, self.U, self.log_s, self.P, self.sign_s)
        return self._apply_matrix(z, log_df_dz, self.matrices()[0], 1.0)

    def backward(self, y, log_df_dz):
        return self._apply_matrix(y, log_df_dz, self.matrices()[1], -1.0)

    inverse = backward


def _actnorm_invconv(an, conv, z, log_df_dz):
    """ActNorm.forward + InvertibleConv1x1.forward as one launch; None if the shape needs the separate kernels."""
    z, log_df_dz = L.dev(z, 'z'), L.dev(log_df_dz, 'log_df_dz')
    B, C, HW = _bchw(z)
    out = torch.empty_like(z)
    rc = L.lib().nfb_actnorm_invconv_fwd(L.ptr(z), L.ptr(out), L.ptr(log_df_dz), L.ptr(log_df_dz),
                                         L.ptr(an.log_scale.data), L.ptr(an.bias.data), L.ptr(conv.matrices()[0]),
                                         L.ptr(conv.log_s.data), B, C, HW, L.stream())
    if rc == L.ERR_UNSUPPORTED:
        return None
    L.check(rc)
    return out, log_df_dz


def _invconv_actnorm_inv(an, conv, y, log_df_dz):
    """InvertibleConv1x1.backward + ActNorm.backward as one launch; None if the shape needs the separate kernels."""
    y, log_df_dz = L.dev(y, 'y'), L.dev(log_df_dz, 'log_df_dz')
    B, C, HW = _bchw(y)
    out = torch.empty_like(y)
    rc = L.lib().nfb_invconv_actnorm_inv(L.ptr(y), L.ptr(out), L.ptr(log_df_dz), L.ptr(log_df_dz),
                                         L.ptr(conv.matrices()[1]), L.ptr(conv.log_s.data), L.ptr(an.log_scale.data),
                                         L.ptr(an.bias.data), B, C, HW, L.stream())
    if rc == L.ERR_UNSUPPORTED:
        return None
    L.check(rc)
    return out, log_df_dz


class Compose(nn.Module):
    """modules.py:325-339: sequential / reversed application."""

    # 1 (default): two launches per Glow step -- ActNorm + 1x1 conv, then the tensor-core conditioner with the affine coupling
    # as its epilogue, in place; 2: the conditioner kernel also applies the NEXT step's ActNorm + 1x1 conv (one launch per
    # flow step for every (map, channel) combination of the Glow stacks: 172 instead of 329 launches per Glow-32 step;
    # measured on B200: slower than 1, 67.3 k vs 75.0 k samples/s with 5 batches in flight and no better for one batch -- the
    # per-pixel matrix product runs as a serial tail on the 256 epilogue threads of the conditioner CTAs while their tensor
    # pipe idles (ncu SM-time: +0.7 ms on the conditioner kernels for -0.4 ms of separate launches, warm) -- so it is not the
    # default); 0: every layer through its own forward()
    fuse_steps = 1

    def __init__(self, layers):
        super().__init__()
        self.layers = nn.ModuleList(layers)

    def forward(self, z, log_df_dz):
        layers = self.layers
        n, i = len(layers), 0
        # the fused peepholes are inference kernels; with autograd recording every layer runs through its own Function
        fuse = self.fuse_steps if not (torch.is_grad_enabled() and (z.requires_grad or log_df_dz.requires_grad or any(
            p.requires_grad for p in self.parameters()))) else 0
        owned = False  # z is a tensor produced inside this loop (safe to update in place)
        while i < n:
            layer = layers[i]
            # peepholes (same arithmetic, fewer launches): an initialised ActNorm followed by the 1x1 convolution as one kernel
            if (i + 1 < n and type(layer) is ActNorm and layer.initialized and type(layers[i + 1]) is InvertibleConv1x1
                    and fuse):
                out = _actnorm_invconv(layer, layers[i + 1], z, log_df_dz)
                if out is not None:
                    z, log_df_dz = out
                    i += 2
                    owned = True
                    continue
            # an affine coupling whose input this loop produced itself: conditioner + transform in place, one kernel -- which
            # also applies the NEXT step's ActNorm + 1x1 conv when those follow (a whole Glow step per launch)
            if fuse and owned and hasattr(layer, 'forward_fused'):
                if (self.fuse_steps > 1 and i + 2 < n and type(layers[i + 1]) is ActNorm and layers[i + 1].initialized
                        and type(layers[i + 2]) is InvertibleConv1x1):
                    out = layer.forward_fused(L.dev(z, 'z'), L.dev(log_df_dz, 'log_df_dz'), inplace=True,
                                              post=(layers[i + 1], layers[i + 2]))
                    if out is not None:
                        z, log_df_dz = out
                        i += 3
                        continue
                out = layer.forward_fused(L.dev(z, 'z'), L.dev(log_df_dz, 'log_df_dz'), inplace=True)
                if out is not None:
                    z, log_df_dz = out
                    i += 1
                    continue
            z, log_df_dz = layer(z, log_df_dz)
            owned = True  # every layer returns a fresh z (the reference's layers never alias their input either)
            i += 1
        return z, log_df_dz

    def backward(self, z, log_df_dz):
        layers = self.layers
        i = len(layers) - 1
        while i >= 0:
            layer = layers[i]
            # peephole (mirror of forward): 1x1 conv inverse followed by the ActNorm inverse as one kernel
            if (i >= 1 and type(layer) is InvertibleConv1x1 and type(layers[i - 1]) is ActNorm and layers[i - 1].initialized
                    and self.fuse_steps):
                out = _invconv_actnorm_inv(layers[i - 1], layer, z, log_df_dz)
                if out is not None:
                    z, log_df_dz = out
                    i -= 2
                    continue
            z, log_df_dz = layer.backward(z, log_df_dz)
            i -= 1
        return z, log_df_dz

    inverse = backward
