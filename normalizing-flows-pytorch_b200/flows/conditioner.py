"""Conditioner networks of the couplings (modules.py:342-438, 500-578), same state_dict keys as the reference.

``MLP`` / ``ConvNet`` (RealNVP, Glow): in_block.0 -> 2 residual blocks (BN, ReLU, WN-layer, BN, ReLU, WN-layer)
-> out_block (BN, ReLU, WN-layer).  Eval mode without autograd: ONE fused kernel over folded weights.  Train mode (batch
statistics, modules.py:349-352): ``ConvNet`` at 32x32 / 16x16 / 8x8 / 4x4 and ``MLP`` (batch a multiple of 16) run layer by layer
on libnfb200 kernels, forward and backward (``conditioner_train.py``); other shapes and eval-mode-with-gradients use
cuDNN / cuBLAS ops under torch autograd (``_forward_autograd``).
"""
import ctypes

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib as L
from .weight_norm import WeightNorm



class _ResBlock(nn.Module):
    """parameter layout of ResBlockLinear / ResBlock2d (modules.py:342-388); `bridge` is empty (in == out)."""

    def __init__(self, ch, conv):
        super().__init__()
        bn = nn.BatchNorm2d if conv else nn.BatchNorm1d
        mk = (lambda: nn.Conv2d(ch, ch, 3, 1, 1)) if conv else (lambda: nn.Linear(ch, ch))
        self.net = nn.Sequential(bn(ch), nn.ReLU(inplace=True), WeightNorm(mk()), bn(ch), nn.ReLU(inplace=True),
                                 WeightNorm(mk()))
        self.bridge = nn.Sequential()


class _ResNetConditioner(nn.Module):
    conv = False
    # kernel selection of the fused ConvNet (flags of nfb_convnet_fwd_ex; 0 = default: tensor-core kernel).  A per-module
    # attribute, not library state: tests and the profiling scripts set it on the instance they measure.
    kernel_flags = 0

    def __init__(self, in_channels, out_channels, base_filters=32, n_blocks=2, weight_norm=True):
        super().__init__()
        if not weight_norm:
            raise NotImplementedError('nfb200 conditioners are weight-normalised like every use in the reference')
        self.in_channels, self.out_channels, self.base_filters = in_channels, out_channels, base_filters
        if self.conv:
            first, last = nn.Conv2d(in_channels, base_filters, 3, 1, 1), nn.Conv2d(base_filters, out_channels, 1, 1, 0)
            bn = nn.BatchNorm2d
        else:
            first, last = nn.Linear(in_channels, base_filters), nn.Linear(base_filters, out_channels)
            bn = nn.BatchNorm1d
        self.in_block = nn.Sequential(WeightNorm(first))
        self.mid_block = nn.Sequential(*[_ResBlock(base_filters, self.conv) for _ in range(n_blocks)])
        self.out_block = nn.Sequential(bn(base_filters), nn.ReLU(inplace=True), WeightNorm(last))

    # -- packed (WeightNorm + BatchNorm folded) weights for the fused kernel, rebuilt when any tensor changes ----
    def _tensors(self):
        wn = [self.in_block[0]] + [blk.net[i] for blk in self.mid_block for i in (2, 5)] + [self.out_block[2]]
        bn = [blk.net[i] for blk in self.mid_block for i in (0, 3)] + [self.out_block[0]]
        ts = []
        for m in wn:
            ts += [m.module.weight_v, m.module.weight_g, m.module.bias]
        for m in bn:
            ts += [m.weight, m.bias, m.running_mean, m.running_var]
        return ts, wn[0].eps, bn[0].eps

    def _apply(self, fn, *a, **k):  # .to() / .cuda(): storage moves, drop every cache
        self._pack_key = self._ts = None
        return super()._apply(fn, *a, **k)

    def packed(self):
        if getattr(self, '_ts', None) is None:
            self._ts = self._tensors()
        ts, wn_eps, bn_eps = self._ts
        key = L.param_key(ts)  # in-place updates bump _version; graph-replayed optimizer steps bump the global epoch
        if key != getattr(self, '_pack_key', None):
            if len(self.mid_block) != 2 or self.base_filters != 32:
                raise NotImplementedError('fused conditioner kernel is built for base_filters=32, n_blocks=2 '
                                          '(every use in the reference)')
            dev = ts[0].device
            n = L.lib().nfb_resnet_pack_size(self.in_channels, self.out_channels, int(self.conv))
            if n <= 0:
                L.check(n)
            self._pack = torch.empty(n, device=dev, dtype=torch.float32)
            arr = (ctypes.c_void_p * 38)(*[L.ptr(L.dev(t.data, 'conditioner parameter')) for t in ts])
            L.check(L.lib().nfb_resnet_pack(arr, L.ptr(self._pack), self.in_channels, self.out_channels,
                                            int(self.conv), float(wn_eps), float(bn_eps), L.stream()))
            self._pack_key = key
        return self._pack

    def _fused_ok(self, x):
        """The fused kernel is the eval-mode inference path; train mode and autograd take `_forward_autograd` (libnfb200
        layer kernels where they apply, torch ops otherwise)."""
        if self.training:
            return False
        return not (torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())))

    def forward(self, x):
        """params = net(x) for an explicit conditioner input x: (B, in, h, w) or (B, in)."""
        if not self._fused_ok(x):
            return self._forward_autograd(x)
        x = L.dev(x, 'conditioner input')
        B = x.size(0)
        if self.conv:
            h, w = x.size(2), x.size(3)
            out = torch.empty((B, self.out_channels, h, w), device=x.device, dtype=torch.float32)
            rc = L.lib().nfb_convnet_fwd_ex(L.ptr(x), L.ptr(out), L.ptr(self.packed()), B, 0, h, w, -1, 0,
                                            self.in_channels, self.out_channels, int(self.kernel_flags), L.stream())
            if rc == L.ERR_UNSUPPORTED:
                return self._forward_library(x)
        else:
            out = torch.empty((B, self.out_channels), device=x.device, dtype=torch.float32)
            rc = L.lib().nfb_mlp_fwd(L.ptr(x), L.ptr(out), L.ptr(self.packed()), B, 0, -1, 0, self.in_channels,
                                     self.out_channels, L.stream())
        L.check(rc)
        return out

    def forward_from_z(self, z, mode, odd):
        """params = net(z1) with z1 (the pass-through half of the coupling split) gathered inside the kernel."""
        if not self._fused_ok(z):
            return None
        B = z.size(0)
        if self.conv:
            _, C, H, W = z.shape
            h, w = (H // 2, W // 2) if mode == L.SPLIT_CHECKER else (H, W)
            out = torch.empty((B, self.out_channels, h, w), device=z.device, dtype=torch.float32)
            rc = L.lib().nfb_convnet_fwd_ex(L.ptr(z), L.ptr(out), L.ptr(self.packed()), B, C, H, W, mode, int(odd),
                                            self.in_channels, self.out_channels, int(self.kernel_flags), L.stream())
            if rc == L.ERR_UNSUPPORTED:
                return None
        else:
            out = torch.empty((B, self.out_channels), device=z.device, dtype=torch.float32)
            rc = L.lib().nfb_mlp_fwd(L.ptr(z), L.ptr(out), L.ptr(self.packed()), B, z.size(1), mode, int(odd),
                                     self.in_channels, self.out_channels, L.stream())
        L.check(rc)
        return out

    # -- library path (cuDNN / cuBLAS through torch, TF32 off): only for eval-mode conditioner inputs whose spatial size the
    #    fused kernel does not cover yet (anything but 16x16 / 8x8 / 4x4, e.g. the 32x32 level of a 64x64 Glow), and
    #    as an on-device cross-check in the tests -------------------------------------------------------------
    def _layer(self, wn, x):
        w = wn.weight()
        if self.conv:
            return F.conv2d(x, w, wn.bias, 1, (w.size(2) - 1) // 2)
        return F.linear(x, w, wn.bias)

    @staticmethod
    def _bn_relu(bn, x):
        return F.relu(F.batch_norm(x, bn.running_mean, bn.running_var, bn.weight, bn.bias, False, 0.0, bn.eps))

    def _forward_library(self, x):
        L.note_library_path('%s eval forward at %s' % (type(self).__name__, tuple(x.shape[2:])))
        x = L.dev(x, 'conditioner input')
        with torch.no_grad():
            x = self._layer(self.in_block[0], x)
            for blk in self.mid_block:
                y = self._bn_relu(blk.net[0], x)
                y = self._layer(blk.net[2], y)
                y = self._bn_relu(blk.net[3], y)
                y = self._layer(blk.net[5], y)
                x = x + y
            x = self._bn_relu(self.out_block[0], x)
            return self._layer(self.out_block[2], x)


    # -- train mode / autograd: torch ops with the reference's semantics (batch statistics + running-stat update by the
    #    nn.BatchNorm modules, WeightNorm recomputed differentiably as in weight_norm.py:40) ---------------------------
    def _wn_apply(self, wn, x):
        v, g = wn.module.weight_v, wn.module.weight_g
        w = v * (g / (torch.norm(v, dim=0) + wn.eps)).expand_as(v)
        if self.conv:
            return F.conv2d(x, w, wn.module.bias, 1, (w.size(2) - 1) // 2)
        return F.linear(x, w, wn.module.bias)

    # train mode on libnfb200 kernels (ConvNet at 32x32 / 16x16 / 8x8 / 4x4, MLP); False = cuDNN / cuBLAS ops under torch autograd
    native_train = True

    def _forward_native_train(self, x):
        from .conditioner_train import SUPPORTED_HW, ConvNetTrainFn, RowsPlanesFn
        if not (self.training and self.native_train and len(self.mid_block) == 2 and self.base_filters == 32):
            return None
        hw = None
        if self.conv:
            if x.dim() != 4 or tuple(x.shape[2:]) not in SUPPORTED_HW:
                return None
        else:  # MLP: rows -> planes of hw*hw rows, then every Linear is a 1x1 convolution of the layer kernels
            if x.dim() != 2 or x.size(0) < 16:
                return None
            hw = next((s for s in (16, 8, 4) if x.size(0) % (s * s) == 0 and x.size(0) // (s * s) <= 65535), None)
            if hw is None:
                return None
            x = RowsPlanesFn.apply(x, hw, True)
        wn = [self.in_block[0]] + [blk.net[i] for blk in self.mid_block for i in (2, 5)] + [self.out_block[2]]
        bn = [blk.net[i] for blk in self.mid_block for i in (0, 3)] + [self.out_block[0]]
        ts = []
        for m in wn:
            ts += [m.module.weight_v, m.module.weight_g, m.module.bias]
        for m in bn:
            ts += [m.weight, m.bias]
        for m in bn:
            ts += [m.running_mean, m.running_var]
        out = ConvNetTrainFn.apply(x, (wn[0].eps, bn[0].eps, bn[0].momentum), *ts)
        for m in bn:
            m.num_batches_tracked += 1  # nn.BatchNorm bookkeeping (momentum is fixed, so it does not enter the update)
        return out if hw is None else RowsPlanesFn.apply(out, hw, False)

    def _forward_autograd(self, x):
        x = L.dev(x, 'conditioner input')
        out = self._forward_native_train(x)
        if out is not None:
            return out
        L.note_library_path('%s %s forward under autograd at %s' % (type(self).__name__, 'train' if self.training else 'eval',
                                                                   tuple(x.shape[1:])))
        x = self._wn_apply(self.in_block[0], x)
        for blk in self.mid_block:
            y = F.relu(blk.net[0](x))
            y = self._wn_apply(blk.net[2], y)
            y = F.relu(blk.net[3](y))
            y = self._wn_apply(blk.net[5], y)
            x = x + y
        x = F.relu(self.out_block[0](x))
        return self._wn_apply(self.out_block[2], x)


def set_throughput_mode(on=True):
    """Default kernel selection of every ConvNet conditioner that has no `kernel_flags` of its own.  on: several batches are
    in flight (streams, one CUDA graph per stream), so the cost of a kernel is the SM-time it occupies: two units per CTA
    on every map size (NFB_CONV_PAIR: more work per SM-second on fewer SMs).  off (default): the lowest latency of a single
    batch.  Results are the same to rounding.  Captured CUDA graphs keep the choice they were captured with.
    (One launch per flow step -- Compose.fuse_steps = 2 -- is NOT part of it: measured 67.3 k vs 75.0 k samples/s on Glow-32,
    the post-op's serial tail costs the conditioner CTAs more SM-time than the separate 1x1-conv launches take.)"""
    _ResNetConditioner.kernel_flags = L.CONV_PAIR if on else 0


class MLP(_ResNetConditioner):
    """modules.py:391-413."""
    conv = False


class ConvNet(_ResNetConditioner):
    """modules.py:416-438."""
    conv = True


# ---- Flow++ conditioner pieces (modules.py:500-578): parameter holders with the reference's keys.  Inference runs them as
#      one kernel (nfb_flowpp_cond_fwd / nfb_flowpp_mlp_fwd, see coupling.py); these torch-op forwards serve train mode,
#      autograd and the spatial sizes the kernels do not cover ----


class GatedLinear(nn.Module):
    def __init__(self, in_features, out_features):
        super().__init__()
        self.op = nn.Linear(in_features * 2, out_features)

    def forward(self, x):
        C = x.size(1)
        y = self.op(F.elu(torch.cat([x, -x], dim=1)))
        y = F.elu(torch.cat([y, -y], dim=1))
        return x + y[:, :C] * torch.sigmoid(y[:, C:])


class GatedConv2d(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.op = nn.Conv2d(in_channels * 2, out_channels, 3, 1, 1)

    def forward(self, x):
        C = x.size(1)
        y = self.op(F.elu(torch.cat([x, -x], dim=1)))
        y = F.elu(torch.cat([y, -y], dim=1))
        return x + y[:, :C] * torch.sigmoid(y[:, C:])


class GatedAttn(nn.Module):
    def __init__(self, in_out_shape, filters=8, heads=4):
        super().__init__()
        assert filters % heads == 0
        self.channels, self.filters, self.heads = in_out_shape[0], filters, heads
        self.conv1 = nn.Conv1d(self.channels, filters * 3, 1, 1, 0)
        self.conv2 = nn.Conv1d(filters, self.channels * 2, 1, 1, 0)
        self.pos_emb = nn.Parameter(torch.randn(1, *in_out_shape) * 0.01)

    def forward(self, x):
        shape = x.size()
        B, C = shape[0], shape[1]
        D = self.filters // self.heads
        p = self.conv1((x + self.pos_emb).view(B, C, -1)).view(B, 3 * self.heads, D, -1)
        V, K, Q = torch.split(p, self.heads, dim=1)  # order of modules.py:566
        Wm = F.softmax(torch.matmul(V.permute(0, 1, 3, 2), K) / np.sqrt(D), dim=2)
        A = torch.matmul(Q, Wm).view(B, C, -1)
        y = self.conv2(A)
        y = (y[:, :C] * torch.sigmoid(y[:, C:])).view(shape)
        return x + y
