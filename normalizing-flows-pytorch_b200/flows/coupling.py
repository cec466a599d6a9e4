"""Coupling layers of the reference's ``flows/coupling.py`` with the bijection, the split-gather / merge-scatter
and the per-sample log-det fused into one libnfb200 kernel each.

Only the conditioner input z1 is materialised (one gather); z0, (t, s), the merged output and the log-det never
exist as separate tensors.
"""
import ctypes

import torch
import torch.nn as nn

from .. import _lib as L
from . import autograd as G
from .conditioner import MLP, ConvNet, GatedAttn, GatedConv2d, GatedLinear
from .squeeze import coupling_split


class AbstractCoupling(nn.Module):
    """coupling.py:12-49: chooses the split from (len(dims), masking)."""

    def __init__(self, dims, masking='checkerboard', odd=False):
        super().__init__()
        self.dims = tuple(dims)
        self.odd = bool(odd)
        if len(dims) == 1:
            self.mode = L.SPLIT_1D
        elif len(dims) == 3 and masking == 'checkerboard':
            self.mode = L.SPLIT_CHECKER
        elif len(dims) == 3 and masking == 'channelwise':
            self.mode = L.SPLIT_CHANNEL
        else:
            raise Exception('unsupported combination of masking and dimension: %s, %s' % (masking, str(dims)))

    def _half_channels(self):
        """(in_chs of the conditioner, channels of the transformed half) -- coupling.py:92-101."""
        d = self.dims
        if len(d) == 1:
            in_chs = d[0] // 2 if not self.odd else (d[0] + 1) // 2
            return in_chs, d[0] - in_chs
        c = d[0] * 2 if self.mode == L.SPLIT_CHECKER else d[0] // 2
        return c, c

    def _geom(self, z):
        if z.dim() == 2:
            return z.size(0), z.size(1), 1, 1
        return tuple(z.shape)

    def _params(self, z, net=None):
        net = self.net if net is None else net
        if hasattr(net, 'forward_from_z'):  # fused conditioner: gathers z1 from z inside the kernel
            p = net.forward_from_z(z, self.mode, self.odd)
            if p is not None:
                return p
        # train mode / autograd: z1 is materialised (its gradient is scattered back by SplitHalfFn)
        z1 = self._z1(z)
        return L.dev(net(z1), 'conditioner output')

    def _z1(self, z):
        if G.needs_grad(z):
            return G.SplitHalfFn.apply(z, self.mode, self.odd)
        return coupling_split(z, self.mode, self.odd, want_z0=False)[1]

    def _recording(self, z, ldj):
        """True when this call has to be differentiable (any input or parameter of the layer requires grad)."""
        return torch.is_grad_enabled() and (z.requires_grad or ldj.requires_grad
                                            or any(p.requires_grad for p in self.parameters()))

    def forward(self, z, log_df_dz):
        return self._run(L.dev(z, 'z'), L.dev(log_df_dz, 'log_df_dz'), False)

    def backward(self, y, log_df_dz):
        """Inverse direction: inference kernels only (the reference samples under no_grad, main.py:121-124)."""
        return self._run(L.dev(y, 'y'), L.dev(log_df_dz, 'log_df_dz'), True)

    inverse = backward


class AdditiveCoupling(AbstractCoupling):
    """coupling.py:52-79 (NICE); no log-det."""

    def __init__(self, dims, masking='checkerboard', odd=False):
        super().__init__(dims, masking, odd)
        in_chs, out_chs = self._half_channels()
        self.net_t = MLP(in_chs, out_chs) if len(dims) == 1 else ConvNet(in_chs, out_chs)

    def _run(self, z, ldj, inverse):
        if not inverse and self._recording(z, ldj):
            raise NotImplementedError('nfb200: AdditiveCoupling has no gradient kernel (no model of the reference uses it)')
        B, C, H, W = self._geom(z)
        t = self._params(z, self.net_t)
        out = torch.empty_like(z)
        L.check(L.lib().nfb_additive_coupling(L.ptr(z), L.ptr(out), L.ptr(t), -1.0 if inverse else 1.0, B, C, H, W,
                                              self.mode, int(self.odd), L.stream()))
        return out, ldj


class AffineCoupling(AbstractCoupling):
    """coupling.py:82-122 (RealNVP / Glow)."""

    def __init__(self, dims, masking='checkerboard', odd=False):
        super().__init__(dims, masking, odd)
        self.s_log_scale = nn.Parameter(torch.randn(1) * 0.01)
        self.s_bias = nn.Parameter(torch.randn(1) * 0.01)
        in_chs, self.out_chs = self._half_channels()
        self.net = MLP(in_chs, self.out_chs * 2) if len(dims) == 1 else ConvNet(in_chs, self.out_chs * 2)

    # conditioner + bijection + log-det as ONE tensor-core kernel (nfb_convnet_affine_fwd); False = two kernels
    fused_conditioner = True

    def forward_fused(self, z, ldj, inplace=False, post=None):
        """ConvNet conditioner and the affine transform as one kernel; (t, s) never leave the SM.  The kernel works in
        place on z: `inplace=False` (the layer API) first copies z so that the caller's tensor is left alone, like the
        reference's merged output; `Compose` passes inplace=True for a z it owns.  post = (ActNorm, InvertibleConv1x1) of
        the NEXT flow step: applied by the same kernel (nfb_convnet_affine_step_fwd).  None: shape / mode not covered."""
        if (not self.fused_conditioner or z.dim() != 4 or self.net.training or type(self.net) is not ConvNet
                or self._recording(z, ldj)):
            return None
        B, C, H, W = z.shape
        h, w = (H // 2, W // 2) if self.mode == L.SPLIT_CHECKER else (H, W)
        if (h, w) not in ((16, 16), (8, 8), (4, 4)) or B == 0:
            return None
        out = z if inplace else z.clone()
        if post is not None:
            an, conv = post
            rc = L.lib().nfb_convnet_affine_step_fwd(L.ptr(out), L.ptr(ldj), L.ptr(self.net.packed()),
                                                     L.ptr(self.s_log_scale.data), L.ptr(self.s_bias.data),
                                                     L.ptr(an.log_scale.data), L.ptr(an.bias.data), L.ptr(conv.matrices()[0]),
                                                     L.ptr(conv.log_s.data), B, C, H, W, self.mode, int(self.odd),
                                                     int(self.net.kernel_flags), L.stream())
        else:
            rc = L.lib().nfb_convnet_affine_fwd(L.ptr(out), L.ptr(ldj), L.ptr(self.net.packed()),
                                                L.ptr(self.s_log_scale.data), L.ptr(self.s_bias.data), B, C, H, W, self.mode,
                                                int(self.odd), int(self.net.kernel_flags), L.stream())
        if rc == L.ERR_UNSUPPORTED:
            return None
        L.check(rc)
        return out, ldj

    def _run(self, z, ldj, inverse):
        B, C, H, W = self._geom(z)
        if not inverse:
            out = self.forward_fused(z, ldj)
            if out is not None:
                return out
        params = self._params(z)
        if not inverse and self._recording(z, ldj):
            return G.AffineCouplingFn.apply(z, params, ldj, self.s_log_scale, self.s_bias, self.mode, self.odd)
        out = torch.empty_like(z)
        fn = L.lib().nfb_affine_coupling_inv if inverse else L.lib().nfb_affine_coupling_fwd
        L.check(fn(L.ptr(z), L.ptr(out), L.ptr(params), L.ptr(ldj), L.ptr(ldj), L.ptr(self.s_log_scale.data),
                   L.ptr(self.s_bias.data), B, C, H, W, self.mode, int(self.odd), L.stream()))
        return out, ldj


class MixLogAttnCoupling(AbstractCoupling):
    """coupling.py:125-210 (Flow++): logistic-mixture CDF -> logit -> affine."""

    def __init__(self, dims, masking='checkerboard', odd=False, base_filters=32, n_mixtures=4):
        super().__init__(dims, masking, odd)
        self.n_mixtures = n_mixtures
        self.a_log_scale = nn.Parameter(torch.randn(1) * 0.01)
        self.a_bias = nn.Parameter(torch.randn(1) * 0.01)
        in_chs, out_chs = self._half_channels()
        self.sections = [out_chs] * 2 + [out_chs * n_mixtures] * 3
        if len(dims) == 1:
            mid_shape = (base_filters, )
            self.net = nn.Sequential(nn.Linear(in_chs, base_filters), GatedLinear(base_filters, base_filters),
                                     nn.LayerNorm(mid_shape), GatedAttn(mid_shape, base_filters),
                                     nn.LayerNorm(mid_shape), nn.Linear(base_filters, sum(self.sections)))
        else:
            sp = tuple(d // 2 for d in dims[1:]) if self.mode == L.SPLIT_CHECKER else tuple(dims[1:])
            mid_shape = (base_filters, ) + sp
            self.net = nn.Sequential(nn.Conv2d(in_chs, base_filters, 3, 1, 1), GatedConv2d(base_filters, base_filters),
                                     nn.LayerNorm(mid_shape), GatedAttn(mid_shape, base_filters),
                                     nn.LayerNorm(mid_shape), nn.Conv2d(base_filters, sum(self.sections), 3, 1, 1))
        self._scratch = None
        self._flag = None

    # -- fused conditioner kernel (image case): packed 3x3 weights cached until a parameter changes ------------------
    def _fpp_params(self):
        n = self.net
        return [n[0].weight, n[0].bias, n[1].op.weight, n[1].op.bias, n[2].weight, n[2].bias, n[3].pos_emb,
                n[3].conv1.weight, n[3].conv1.bias, n[3].conv2.weight, n[3].conv2.bias, n[4].weight, n[4].bias,
                n[5].weight, n[5].bias]

    def _fpp_tensors(self):
        """Host array of the 15 device pointers nfb_flowpp_cond_fwd takes (3x3 weights packed, cached until a parameter
        changes).  Only plain Python objects are stored on the module (deepcopy / pickle safe); the ctypes array is built
        per call."""
        ts = self._fpp_params()
        key = L.param_key(ts)
        if key != getattr(self, '_fpp_key', None):
            packed = {}
            for i in (0, 2, 13):
                w = L.dev(ts[i].data, 'conv weight')
                O, I = w.size(0), w.size(1)
                buf = torch.empty(((O + 31) // 32) * 32 * I * 9, device=w.device, dtype=torch.float32)
                L.check(L.lib().nfb_pack_conv3x3(L.ptr(w), L.ptr(buf), O, I, L.stream()))
                packed[i] = buf
            self._fpp_packed = packed
            self._fpp_ptrs = tuple(L.ptr(packed[i]) if i in packed else L.ptr(L.dev(t.data, 'conditioner parameter'))
                                   for i, t in enumerate(ts))
            self._fpp_key = key
        return (ctypes.c_void_p * 15)(*self._fpp_ptrs)

    def _fpp_tensors_1d(self):
        """Host array of the 15 raw parameter pointers (nothing is packed for the 1-D kernel), read per call."""
        return (ctypes.c_void_p * 15)(*[L.ptr(L.dev(t.data, 'conditioner parameter')) for t in self._fpp_params()])

    def _apply(self, fn, *a, **k):
        self._fpp_key = None
        return super()._apply(fn, *a, **k)

    def _params(self, z):
        if z.dim() == 4 and not self.net.training and self.fused_conditioner and not self._recording(z, z):
            B, C, H, W = z.shape
            h, w = (H // 2, W // 2) if self.mode == L.SPLIT_CHECKER else (H, W)
            n_out = sum(self.sections)
            out = torch.empty((B, n_out, h, w), device=z.device, dtype=torch.float32)
            in_chs = C * 2 if self.mode == L.SPLIT_CHECKER else C // 2
            rc = L.lib().nfb_flowpp_cond_fwd(self._fpp_tensors(), L.ptr(z), L.ptr(out), B, C, H, W, self.mode,
                                             int(self.odd), in_chs, n_out, L.stream())
            if rc != L.ERR_UNSUPPORTED:
                L.check(rc)
                return out
        if z.dim() == 2 and not self.net.training and self.fused_conditioner and not self._recording(z, z):
            # 1-D couplings: the whole conditioner is one per-sample kernel (a single attention token: A = Q)
            ts = self._fpp_tensors_1d()
            n_out = sum(self.sections)
            out = torch.empty((z.size(0), n_out), device=z.device, dtype=torch.float32)
            rc = L.lib().nfb_flowpp_mlp_fwd(ts, L.ptr(z), L.ptr(out), z.size(0), z.size(1), self.mode, int(self.odd),
                                            z.size(1) // 2, n_out, L.stream())
            if rc != L.ERR_UNSUPPORTED:
                L.check(rc)
                return out
        # library path (torch ops on the device): spatial sizes the kernels do not cover, train mode
        L.note_library_path('Flow++ conditioner %s at %s' % ('train' if self.net.training else 'eval', tuple(z.shape[1:])))
        z1 = self._z1(z)
        return L.dev(self.net(z1), 'conditioner output')

    fused_conditioner = True

    def _run(self, z, ldj, inverse):
        B, C, H, W = self._geom(z)
        params = self._params(z)
        if not inverse and self._recording(z, ldj):
            return G.MixLogCouplingFn.apply(z, params, ldj, self.a_log_scale, self.a_bias, self.mode, self.odd,
                                            self.n_mixtures)
        out = torch.empty_like(z)
        ldj_out = torch.empty_like(ldj)  # MixLogCDF / Logit return new log-det tensors (modules.py:194,150)
        if not inverse:
            L.check(L.lib().nfb_mixlog_coupling_fwd(L.ptr(z), L.ptr(out), L.ptr(params), L.ptr(ldj), L.ptr(ldj_out),
                                                    L.ptr(self.a_log_scale.data), L.ptr(self.a_bias.data), B, C, H, W,
                                                    self.mode, int(self.odd), self.n_mixtures, L.stream()))
            return out, ldj_out
        n = z.numel()  # 2 floats (lo, hi) per transformed element
        if self._scratch is None or self._scratch.numel() < n or self._scratch.device != z.device:
            self._scratch = torch.empty(n, device=z.device, dtype=torch.float32)
            self._flag = torch.zeros(1, device=z.device, dtype=torch.int32)
        L.check(L.lib().nfb_mixlog_coupling_inv(L.ptr(z), L.ptr(out), L.ptr(params), L.ptr(ldj), L.ptr(ldj_out),
                                                L.ptr(self.a_log_scale.data), L.ptr(self.a_bias.data),
                                                L.ptr(self._scratch), L.ptr(self._flag), B, C, H, W, self.mode,
                                                int(self.odd), self.n_mixtures, L.stream()))
        return out, ldj_out


class RQSplineCoupling(AbstractCoupling):
    """Rational-quadratic spline coupling (Durkan et al. 2019) inside the reference's coupling container
    (coupling.py:12-49).  Not present in the reference; conditioner = the reference's MLP / ConvNet
    emitting (3K-1) values per transformed feature, bin-major."""

    def __init__(self, dims, masking='checkerboard', odd=False, n_bins=8, tail_bound=3.0):
        super().__init__(dims, masking, odd)
        self.n_bins, self.tail_bound = n_bins, float(tail_bound)
        in_chs, out_chs = self._half_channels()
        n_out = out_chs * (3 * n_bins - 1)
        self.net = MLP(in_chs, n_out) if len(dims) == 1 else ConvNet(in_chs, n_out)

    def _run(self, z, ldj, inverse):
        B, C, H, W = self._geom(z)
        params = self._params(z)
        if not inverse and self._recording(z, ldj):
            return G.RQSCouplingFn.apply(z, params, ldj, self.mode, self.odd, self.n_bins, self.tail_bound)
        out = torch.empty_like(z)
        fn = L.lib().nfb_rqs_coupling_inv if inverse else L.lib().nfb_rqs_coupling_fwd
        L.check(fn(L.ptr(z), L.ptr(out), L.ptr(params), L.ptr(ldj), L.ptr(ldj), B, C, H, W, self.mode, int(self.odd),
                   self.n_bins, self.tail_bound, L.stream()))
        return out, ldj
