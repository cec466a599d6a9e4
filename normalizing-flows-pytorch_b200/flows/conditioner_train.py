"""Train-mode ``ConvNet`` conditioner (modules.py:416-438 under ``net.train()``) on libnfb200 kernels, forward and backward.

In train mode the conditioner's BatchNorm layers use batch statistics, so the one-kernel eval conditioner does not apply:
the network runs layer by layer (``csrc/conditioner_train.cu``) -- WeightNorm + packing, convolution (+ bias, + residual,
+ channel moments), BatchNorm+ReLU -- and the backward pass mirrors it (weight gradient, data gradient = the same
convolution kernel over flipped / transposed packed weights, BatchNorm+ReLU backward in two passes).  Every intermediate
activation is kept for the backward pass (11 tensors of (B, 32, h, w) per conditioner call).

``ConvNetTrainFn.apply(x, cfg, *tensors)`` with ``tensors`` =
  6 x (weight_v, weight_g, bias)   of in_block.0, mid_block.{0,1}.net.{2,5}, out_block.2          (differentiable)
  5 x (weight, bias)               of mid_block.{0,1}.net.{0,3}, out_block.0 (BatchNorm gamma / beta) (differentiable)
  5 x (running_mean, running_var)  of the same BatchNorms                                      (updated in place)
"""
import ctypes

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import _lib as L

SUPPORTED_HW = ((32, 32), (16, 16), (8, 8), (4, 4))  # 32x32 (first level of a 64x64 Glow) runs as four 8 x 32 bands
F32 = 32  # base_filters


def _conv(x, w_packed, bias, skip, cin, cout, ks, want_stats, stats=None):
    """`stats`: a pre-zeroed double[2*cout] slice of the caller's accumulator arena (one fill for the whole network)."""
    B, _, h, w = x.shape
    out = torch.empty((B, cout, h, w), device=x.device, dtype=torch.float32)
    zeroed = stats is not None
    if want_stats and stats is None:
        stats = torch.empty(2 * cout, device=x.device, dtype=torch.float64)
    L.check(L.lib().nfb_conv_train(L.ptr(x), L.ptr(w_packed), L.ptr(bias) if bias is not None else None,
                                   L.ptr(skip) if skip is not None else None, L.ptr(out),
                                   stats.data_ptr() if want_stats else None, int(zeroed), B, cin, cout, h, w, ks, L.stream()))
    return out, stats


def _bn_relu(x, stats, gamma, beta, rmean, rvar, momentum, eps):
    B, C, h, w = x.shape
    a = torch.empty_like(x)
    mr = torch.empty(2 * C, device=x.device, dtype=torch.float32)
    L.check(L.lib().nfb_bn_relu_fwd(L.ptr(x), stats.data_ptr(), L.ptr(gamma), L.ptr(beta), L.ptr(rmean), L.ptr(rvar),
                                    float(momentum), float(eps), L.ptr(a), L.ptr(mr), B, C, h * w, L.stream()))
    return a, mr


def _wgrad(gy, a, cin, cout, ks, side=None, keep=None):
    """Weight / bias gradient.  With `side` (a CUDA stream) the kernels are enqueued there, after everything issued so far
    on the current stream, so they overlap with the data-gradient chain that follows; the caller joins the streams before
    it reads the results and holds `keep` (the scratch buffers) until then."""
    B, _, h, w = gy.shape
    gw = torch.empty((cout, cin, ks, ks), device=gy.device, dtype=torch.float32)
    gb = torch.empty(cout, device=gy.device, dtype=torch.float32)
    n = int(L.lib().nfb_conv_train_wgrad_scratch(B, cin, cout, h, w, ks))
    if n < 0:
        L.check(n)
    scratch = torch.empty(n, device=gy.device, dtype=torch.float32)  # per-sample-group partial sums (no zero-fill needed)
    if side is None:
        L.check(L.lib().nfb_conv_train_wgrad(L.ptr(gy), L.ptr(a), L.ptr(gw), L.ptr(gb), L.ptr(scratch), B, cin, cout, h, w,
                                             ks, L.stream()))
    else:
        keep.extend([scratch, gy, a])  # allocated on the main stream: must outlive the side-stream kernels that use them
        side.wait_stream(torch.cuda.current_stream())
        L.check(L.lib().nfb_conv_train_wgrad(L.ptr(gy), L.ptr(a), L.ptr(gw), L.ptr(gb), L.ptr(scratch), B, cin, cout, h, w,
                                             ks, side.cuda_stream))
    return gw, gb


_SIDE = {}


def _side_stream(device):
    key = (device.type, device.index)
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=device)
    return _SIDE[key]


def _bn_relu_bwd(ga, a, x, mr, gamma, add, sums=None):
    """-> (gx, g_gamma, g_beta): ReLU mask, the two batch sums, then the BatchNorm input gradient (+ residual branch)."""
    B, C, h, w = x.shape
    U = torch.empty_like(x)
    zeroed = sums is not None
    if sums is None:
        sums = torch.empty(2 * C, device=x.device, dtype=torch.float64)
    L.check(L.lib().nfb_bn_relu_bwd_reduce(L.ptr(ga), L.ptr(a), L.ptr(x), L.ptr(mr), L.ptr(U), sums.data_ptr(), int(zeroed),
                                           B, C, h * w, L.stream()))
    gx = torch.empty_like(x)
    gg, gb = torch.empty_like(gamma), torch.empty_like(gamma)
    L.check(L.lib().nfb_bn_bwd_apply(L.ptr(U), L.ptr(x), L.ptr(mr), L.ptr(gamma), sums.data_ptr(),
                                     L.ptr(add) if add is not None else None, L.ptr(gx), L.ptr(gg), L.ptr(gb), B, C, h * w,
                                     L.stream()))
    return gx, gg, gb


def _dgrad_bn_relu(gy, w_bwd, cin, ks, a, x, mr, gamma, add, sums):
    """Data gradient of a layer (C_in = cin channels of gy -> 32) fused with the ReLU mask and the two BatchNorm sums of
    the layer below, then the BatchNorm input gradient (+ residual branch) -> (gx, g_gamma, g_beta)."""
    B, C, h, w = x.shape
    U = torch.empty_like(x)
    L.check(L.lib().nfb_conv_train_dgrad_bnrelu(L.ptr(gy), L.ptr(w_bwd), L.ptr(a), L.ptr(x), L.ptr(mr), L.ptr(U),
                                                sums.data_ptr(), 1, B, cin, C, h, w, ks, L.stream()))
    gx = torch.empty_like(x)
    gg, gb = torch.empty_like(gamma), torch.empty_like(gamma)
    L.check(L.lib().nfb_bn_bwd_apply(L.ptr(U), L.ptr(x), L.ptr(mr), L.ptr(gamma), sums.data_ptr(),
                                     L.ptr(add) if add is not None else None, L.ptr(gx), L.ptr(gg), L.ptr(gb), B, C, h * w,
                                     L.stream()))
    return gx, gg, gb


class RowsPlanesFn(Function):
    """(B, C) rows -> (B/HW, C, h, w) planes (to_planes) or back; the gradient is the opposite transposition."""

    @staticmethod
    def forward(ctx, x, hw, to_planes):
        x = L.dev(x, 'conditioner tensor')
        ctx.meta = (hw, to_planes)
        return _rows_planes(x, hw, to_planes)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        hw, to_planes = ctx.meta
        return _rows_planes(L.dev(g, 'grad'), hw, not to_planes), None, None


def _rows_planes(x, hw, to_planes):
    if to_planes:
        B, C = x.shape
        out = torch.empty((B // (hw * hw), C, hw, hw), device=x.device, dtype=torch.float32)
        L.check(L.lib().nfb_rows_to_planes(L.ptr(x), L.ptr(out), B, C, hw * hw, L.stream()))
    else:
        G, C = x.shape[:2]
        B = G * hw * hw
        out = torch.empty((B, C), device=x.device, dtype=torch.float32)
        L.check(L.lib().nfb_planes_to_rows(L.ptr(x), L.ptr(out), B, C, hw * hw, L.stream()))
    return out


class ConvNetTrainFn(Function):
    """Also serves the MLP conditioner: with 2-D weights every layer is a 1x1 convolution over planes of rows."""

    # weight gradients on a second stream: they only feed the WeightNorm backward at the very end, so they run next to the
    # data-gradient / BatchNorm-backward chain and fill the SMs its 256-CTA launches leave idle (also inside graph capture)
    overlap_wgrad = True

    @staticmethod
    def forward(ctx, x, cfg, *T):
        x = L.dev(x, 'conditioner input')
        wn, bn, rs = T[:18], T[18:28], T[28:38]
        wn_eps, bn_eps, momentum = cfg
        cin, cout = wn[0].size(1), wn[15].size(0)
        k3 = 3 if wn[0].dim() == 4 and wn[0].size(2) == 3 else 1  # ConvNet: 3x3 body; MLP (2-D weights): all 1x1
        packed, ptrs, dims, keep = [], [], [], []
        for i in range(6):  # WeightNorm + packing of all six layers: one launch
            v, g = wn[3 * i], wn[3 * i + 1]
            O, I, KK = v.size(0), v.size(1), v[0, 0].numel()
            n = ((O + 31) // 32) * ((I + 31) // 32) * 32 * KK * 32
            w_nat = torch.empty_like(v)
            w_fwd = torch.empty(n, device=v.device, dtype=torch.float32)  # padding is written by the kernel
            w_bwd = torch.empty(n, device=v.device, dtype=torch.float32)
            ptrs += [L.ptr(v), L.ptr(g), L.ptr(w_nat), L.ptr(w_fwd), L.ptr(w_bwd)]
            dims += [O, I, KK]
            keep.append(w_nat)
            packed.append((w_fwd, w_bwd))
        L.check(L.lib().nfb_wn_pack_train_multi((ctypes.c_void_p * 30)(*ptrs), (ctypes.c_int * 18)(*dims), 6,
                                                float(wn_eps), L.stream()))
        b = [wn[3 * i + 2] for i in range(6)]
        gam, bet = [bn[2 * i] for i in range(5)], [bn[2 * i + 1] for i in range(5)]

        def bnr(t, st, i):
            return _bn_relu(t, st, gam[i], bet[i], rs[2 * i], rs[2 * i + 1], momentum, bn_eps)

        arena = torch.zeros(5 * 2 * F32, device=x.device, dtype=torch.float64)  # the five moment accumulators, one fill
        acc = [arena[i * 2 * F32:(i + 1) * 2 * F32] for i in range(5)]
        h0, st = _conv(x, packed[0][0], b[0], None, cin, F32, k3, True, acc[0])
        a1, mr1 = bnr(h0, st, 0)
        y1, st = _conv(a1, packed[1][0], b[1], None, F32, F32, k3, True, acc[1])
        a2, mr2 = bnr(y1, st, 1)
        h1, st = _conv(a2, packed[2][0], b[2], h0, F32, F32, k3, True, acc[2])
        a3, mr3 = bnr(h1, st, 2)
        y2, st = _conv(a3, packed[3][0], b[3], None, F32, F32, k3, True, acc[3])
        a4, mr4 = bnr(y2, st, 3)
        h2, st = _conv(a4, packed[4][0], b[4], h1, F32, F32, k3, True, acc[4])
        a5, mr5 = bnr(h2, st, 4)
        out, _ = _conv(a5, packed[5][0], b[5], None, F32, cout, 1, False)
        ctx.save_for_backward(x, h0, a1, y1, a2, h1, a3, y2, a4, h2, a5, mr1, mr2, mr3, mr4, mr5,
                              *[p[1] for p in packed], *[wn[3 * i] for i in range(6)], *[wn[3 * i + 1] for i in range(6)],
                              *gam)
        ctx.meta = (cin, cout, float(wn_eps), k3)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        S = ctx.saved_tensors
        x, h0, a1, y1, a2, h1, a3, y2, a4, h2, a5, mr1, mr2, mr3, mr4, mr5 = S[:16]
        wb, vs, gs, gam = S[16:22], S[22:28], S[28:34], S[34:39]
        cin, cout, wn_eps, k3 = ctx.meta
        gout = L.dev(gout, 'grad params')
        gw, gb, ggam, gbet = [None] * 6, [None] * 6, [None] * 5, [None] * 5
        side = _side_stream(gout.device) if ConvNetTrainFn.overlap_wgrad else None
        keep = []
        # out block
        gw[5], gb[5] = _wgrad(gout, a5, F32, cout, 1, side, keep)
        arena = torch.zeros(5 * 2 * F32, device=gout.device, dtype=torch.float64)  # the five BatchNorm-backward sum pairs
        acc = [arena[i * 2 * F32:(i + 1) * 2 * F32] for i in range(5)]
        G, ggam[4], gbet[4] = _dgrad_bn_relu(gout, wb[5], cout, 1, a5, h2, mr5, gam[4], None, acc[4])
        # residual blocks, last first: (layer indices, BatchNorm indices, activations)
        for (l2, l1, bB, bA, aB, yB, mrB, aA, hA, mrA) in ((4, 3, 3, 2, a4, y2, mr4, a3, h1, mr3),
                                                          (2, 1, 1, 0, a2, y1, mr2, a1, h0, mr1)):
            gw[l2], gb[l2] = _wgrad(G, aB, F32, F32, k3, side, keep)
            gy, ggam[bB], gbet[bB] = _dgrad_bn_relu(G, wb[l2], F32, k3, aB, yB, mrB, gam[bB], None, acc[bB])
            gw[l1], gb[l1] = _wgrad(gy, aA, F32, F32, k3, side, keep)
            G, ggam[bA], gbet[bA] = _dgrad_bn_relu(gy, wb[l1], F32, k3, aA, hA, mrA, gam[bA], G, acc[bA])  # + the skip branch
        gw[0], gb[0] = _wgrad(G, x, cin, F32, k3, side, keep)
        gx = None
        if ctx.needs_input_grad[0]:
            gx, _ = _conv(G, wb[0], None, None, F32, cin, k3, False)
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)  # every weight gradient has landed
        grads, ptrs, dims = [], [], []
        for i in range(6):  # WeightNorm backward of all six layers: one launch
            v, g = vs[i], gs[i]
            gv, gg = torch.empty_like(v), torch.empty_like(g)
            ptrs += [L.ptr(v), L.ptr(g), L.ptr(gw[i]), L.ptr(gv), L.ptr(gg)]
            dims += [v.size(0), v.size(1), v[0, 0].numel()]
            grads += [gv, gg, gb[i]]
        L.check(L.lib().nfb_wn_bwd_multi((ctypes.c_void_p * 30)(*ptrs), (ctypes.c_int * 18)(*dims), 6, wn_eps, L.stream()))
        for i in range(5):
            grads += [ggam[i], gbet[i]]
        return (gx, None) + tuple(grads) + (None, ) * 10
