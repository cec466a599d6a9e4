"""``torch.autograd.Function`` wrappers that make the libnfb200 layers trainable.

The reference gets its gradients from autograd over ~10-45 eager ops per layer (main.py:85-90); here every bijective
layer is one forward kernel plus one backward kernel (``csrc/backward.cu``), and autograd only strings them together.
A layer takes this path when gradients are enabled and one of its inputs or parameters requires them
(:func:`needs_grad`); under ``torch.no_grad()`` the inference kernels (in-place log-det, fused peepholes) run as before.

Conventions: ``forward`` returns fresh ``(z_out, ldj_out)`` tensors; ``backward`` receives ``(gy, gl)`` and passes
``gl`` through to the log-det input unchanged.
"""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import _lib as L


def needs_grad(*tensors):
    if not torch.is_grad_enabled():
        return False
    return any(t is not None and t.requires_grad for t in tensors)


def _geom(z):
    if z.dim() == 2:
        return z.size(0), z.size(1), 1, 1
    return tuple(z.shape)


def _bchw(z):
    if z.dim() == 2:
        return z.size(0), z.size(1), 1
    return z.size(0), z.size(1), z[0, 0].numel()


def _scratch(z, n):
    return torch.empty(n, device=z.device, dtype=torch.float64)


class AffineCouplingFn(Function):
    """coupling.py:104-112 on the full tensor z (split / merge are index arithmetic inside the kernels)."""

    @staticmethod
    def forward(ctx, z, params, ldj, a, b, mode, odd):
        z, params, ldj = L.dev(z, 'z'), L.dev(params, 'conditioner output'), L.dev(ldj, 'log_df_dz')
        B, C, H, W = _geom(z)
        out, ldj_out = torch.empty_like(z), torch.empty_like(ldj)
        L.check(L.lib().nfb_affine_coupling_fwd(L.ptr(z), L.ptr(out), L.ptr(params), L.ptr(ldj), L.ptr(ldj_out),
                                                L.ptr(a), L.ptr(b), B, C, H, W, mode, int(odd), L.stream()))
        ctx.save_for_backward(z, params, a, b)
        ctx.meta = (B, C, H, W, mode, int(odd))
        return out, ldj_out

    @staticmethod
    @once_differentiable
    def backward(ctx, gy, gl):
        z, params, a, b = ctx.saved_tensors
        B, C, H, W, mode, odd = ctx.meta
        gy, gl = L.dev(gy, 'grad z'), L.dev(gl, 'grad log_df_dz')
        gz, gp = torch.empty_like(z), torch.empty_like(params)
        ga, gb = torch.empty_like(a), torch.empty_like(b)
        L.check(L.lib().nfb_affine_coupling_bwd(L.ptr(z), L.ptr(params), L.ptr(gy), L.ptr(gl), L.ptr(gz), L.ptr(gp),
                                                L.ptr(ga), L.ptr(gb), _scratch(z, 2).data_ptr(), L.ptr(a), L.ptr(b), B,
                                                C, H, W, mode, odd, L.stream()))
        return gz, gp, gl, ga, gb, None, None


class MixLogCouplingFn(Function):
    """coupling.py:172-190 (mixture CDF -> logit -> affine), forward direction."""

    @staticmethod
    def forward(ctx, z, params, ldj, a, b, mode, odd, K):
        z, params, ldj = L.dev(z, 'z'), L.dev(params, 'conditioner output'), L.dev(ldj, 'log_df_dz')
        B, C, H, W = _geom(z)
        out, ldj_out = torch.empty_like(z), torch.empty_like(ldj)
        L.check(L.lib().nfb_mixlog_coupling_fwd(L.ptr(z), L.ptr(out), L.ptr(params), L.ptr(ldj), L.ptr(ldj_out),
                                                L.ptr(a), L.ptr(b), B, C, H, W, mode, int(odd), K, L.stream()))
        ctx.save_for_backward(z, params, a, b)
        ctx.meta = (B, C, H, W, mode, int(odd), K)
        return out, ldj_out

    @staticmethod
    @once_differentiable
    def backward(ctx, gy, gl):
        z, params, a, b = ctx.saved_tensors
        B, C, H, W, mode, odd, K = ctx.meta
        gy, gl = L.dev(gy, 'grad z'), L.dev(gl, 'grad log_df_dz')
        gz, gp = torch.empty_like(z), torch.empty_like(params)
        ga, gb = torch.empty_like(a), torch.empty_like(b)
        L.check(L.lib().nfb_mixlog_coupling_bwd(L.ptr(z), L.ptr(params), L.ptr(gy), L.ptr(gl), L.ptr(gz), L.ptr(gp),
                                                L.ptr(ga), L.ptr(gb), _scratch(z, 2).data_ptr(), L.ptr(a), L.ptr(b), B,
                                                C, H, W, mode, odd, K, L.stream()))
        return gz, gp, gl, ga, gb, None, None, None


class RQSCouplingFn(Function):
    """Rational-quadratic spline coupling (Durkan et al. 2019), forward direction."""

    @staticmethod
    def forward(ctx, z, params, ldj, mode, odd, K, bound):
        z, params, ldj = L.dev(z, 'z'), L.dev(params, 'conditioner output'), L.dev(ldj, 'log_df_dz')
        B, C, H, W = _geom(z)
        out, ldj_out = torch.empty_like(z), torch.empty_like(ldj)
        L.check(L.lib().nfb_rqs_coupling_fwd(L.ptr(z), L.ptr(out), L.ptr(params), L.ptr(ldj), L.ptr(ldj_out), B, C, H,
                                             W, mode, int(odd), K, bound, L.stream()))
        ctx.save_for_backward(z, params)
        ctx.meta = (B, C, H, W, mode, int(odd), K, bound)
        return out, ldj_out

    @staticmethod
    @once_differentiable
    def backward(ctx, gy, gl):
        z, params = ctx.saved_tensors
        B, C, H, W, mode, odd, K, bound = ctx.meta
        gy, gl = L.dev(gy, 'grad z'), L.dev(gl, 'grad log_df_dz')
        gz, gp = torch.empty_like(z), torch.empty_like(params)
        L.check(L.lib().nfb_rqs_coupling_bwd(L.ptr(z), L.ptr(params), L.ptr(gy), L.ptr(gl), L.ptr(gz), L.ptr(gp), B, C,
                                             H, W, mode, odd, K, bound, L.stream()))
        return gz, gp, gl, None, None, None, None


class SplitHalfFn(Function):
    """The conditioner input z1 = pass-through half of the coupling split (coupling.py:33); its gradient is scattered
    back into the layout of z (zeros on the transformed half)."""

    @staticmethod
    def forward(ctx, z, mode, odd):
        from .squeeze import coupling_split
        ctx.meta = _geom(z) + (mode, int(bool(odd)))
        return coupling_split(z, mode, odd, want_z0=False)[1]

    @staticmethod
    @once_differentiable
    def backward(ctx, g1):
        B, C, H, W, mode, odd = ctx.meta
        g1 = L.dev(g1, 'grad z1')
        shape = (B, C) if mode == L.SPLIT_1D else (B, C, H, W)
        gz = torch.empty(shape, device=g1.device, dtype=g1.dtype)
        L.check(L.lib().nfb_coupling_merge(None, L.ptr(g1), L.ptr(gz), B, C, H, W, mode, odd, L.stream()))
        return gz, None, None


class ActNormFn(Function):
    """modules.py:246-250."""

    @staticmethod
    def forward(ctx, z, ldj, log_scale, bias):
        z, ldj = L.dev(z, 'z'), L.dev(ldj, 'log_df_dz')
        B, C, HW = _bchw(z)
        out, ldj_out = torch.empty_like(z), torch.empty_like(ldj)
        L.check(L.lib().nfb_actnorm_fwd(L.ptr(z), L.ptr(out), L.ptr(ldj), L.ptr(ldj_out), L.ptr(log_scale), L.ptr(bias),
                                        B, C, HW, L.stream()))
        ctx.save_for_backward(z, log_scale, bias)
        return out, ldj_out

    @staticmethod
    @once_differentiable
    def backward(ctx, gy, gl):
        z, log_scale, bias = ctx.saved_tensors
        B, C, HW = _bchw(z)
        gy, gl = L.dev(gy, 'grad z'), L.dev(gl, 'grad log_df_dz')
        gz, gls, gb = torch.empty_like(z), torch.empty_like(log_scale), torch.empty_like(bias)
        L.check(L.lib().nfb_actnorm_bwd(L.ptr(gy), L.ptr(z), L.ptr(gl), L.ptr(log_scale), L.ptr(bias), L.ptr(gz),
                                        L.ptr(gls), L.ptr(gb), _scratch(z, 2 * C).data_ptr(), B, C, HW, L.stream()))
        return gz, gl, gls, gb


class BatchNormFlowFn(Function):
    """modules.py:300-305; mean / var are buffers (no gradient, like the reference)."""

    @staticmethod
    def forward(ctx, x, ldj, mean, var, log_gamma, beta):
        x, ldj = L.dev(x, 'x'), L.dev(ldj, 'log_det_jacob')
        B, C, HW = _bchw(x)
        out, ldj_out = torch.empty_like(x), torch.empty_like(ldj)
        L.check(L.lib().nfb_bnflow_fwd(L.ptr(x), L.ptr(out), L.ptr(ldj), L.ptr(ldj_out), L.ptr(mean), L.ptr(var),
                                       L.ptr(log_gamma), L.ptr(beta), B, C, HW, L.stream()))
        # the statistics buffers are overwritten by the next training forward: keep this step's values
        ctx.save_for_backward(x, mean.clone(), var.clone(), log_gamma)
        return out, ldj_out

    @staticmethod
    @once_differentiable
    def backward(ctx, gy, gl):
        x, mean, var, log_gamma = ctx.saved_tensors
        B, C, HW = _bchw(x)
        gy, gl = L.dev(gy, 'grad x'), L.dev(gl, 'grad log_det_jacob')
        gx = torch.empty_like(x)
        want = ctx.needs_input_grad[4] or ctx.needs_input_grad[5]
        glg = torch.empty_like(log_gamma) if want else None
        gbt = torch.empty_like(log_gamma) if want else None
        L.check(L.lib().nfb_bnflow_bwd(L.ptr(gy), L.ptr(x), L.ptr(gl), L.ptr(mean), L.ptr(var), L.ptr(log_gamma),
                                       L.ptr(gx), L.ptr(glg) if want else None, L.ptr(gbt) if want else None,
                                       _scratch(x, 2 * C).data_ptr(), B, C, HW, L.stream()))
        return gx, gl, None, None, glg, gbt


class InvConv1x1Fn(Function):
    """modules.py:470-482: W is assembled from (P, L, U, log_s, sign_s) by nfb_invconv1x1_weight inside the function, so
    the gradients come out directly for the trainable factors L, U, log_s."""

    @staticmethod
    def forward(ctx, z, ldj, Lp, Up, log_s, P, sign_s):
        z, ldj = L.dev(z, 'z'), L.dev(ldj, 'log_df_dz')
        B, C, HW = _bchw(z)
        Wm = torch.empty((C, C), device=z.device, dtype=torch.float32)
        L.check(L.lib().nfb_invconv1x1_weight(L.ptr(P), L.ptr(Lp), L.ptr(Up), L.ptr(log_s), L.ptr(sign_s), L.ptr(Wm),
                                              None, C, L.stream()))
        out, ldj_out = torch.empty_like(z), torch.empty_like(ldj)
        L.check(L.lib().nfb_invconv1x1_apply(L.ptr(z), L.ptr(out), L.ptr(ldj), L.ptr(ldj_out), L.ptr(Wm), L.ptr(log_s),
                                             1.0, B, C, HW, L.stream()))
        ctx.save_for_backward(z, Wm, Lp, Up, log_s, P, sign_s)
        return out, ldj_out

    @staticmethod
    @once_differentiable
    def backward(ctx, gy, gl):
        z, Wm, Lp, Up, log_s, P, sign_s = ctx.saved_tensors
        B, C, HW = _bchw(z)
        gy, gl = L.dev(gy, 'grad z'), L.dev(gl, 'grad log_df_dz')
        gz = torch.empty_like(z)
        Wt = Wm.t().contiguous()
        L.check(L.lib().nfb_invconv1x1_apply(L.ptr(gy), L.ptr(gz), None, None, L.ptr(Wt), None, 0.0, B, C, HW,
                                             L.stream()))
        gW = torch.empty_like(Wm)
        scratch = torch.empty(int(L.lib().nfb_invconv1x1_wgrad_scratch(B, C, HW)), device=z.device, dtype=torch.float32)
        L.check(L.lib().nfb_invconv1x1_wgrad(L.ptr(gy), L.ptr(z), L.ptr(gW), L.ptr(scratch), B, C, HW, L.stream()))
        gL, gU, gls = torch.empty_like(Lp), torch.empty_like(Up), torch.empty_like(log_s)
        L.check(L.lib().nfb_invconv1x1_weight_bwd(L.ptr(gW), L.ptr(P), L.ptr(Lp), L.ptr(Up), L.ptr(log_s), L.ptr(sign_s),
                                                  L.ptr(gl), L.ptr(gL), L.ptr(gU), L.ptr(gls), B, C, HW, L.stream()))
        return gz, gl, gL, gU, gls, None, None


class LogitFn(Function):
    """modules.py:146-150."""

    @staticmethod
    def forward(ctx, x, ldj, lo, hi):
        x, ldj = L.dev(x, 'x'), L.dev(ldj, 'log_df_dz')
        out, ldj_out = torch.empty_like(x), torch.empty_like(ldj)
        L.check(L.lib().nfb_logit_fwd(L.ptr(x), L.ptr(out), L.ptr(ldj), L.ptr(ldj_out), lo, hi, x.size(0), x[0].numel(),
                                      L.stream()))
        ctx.save_for_backward(x)
        ctx.meta = (lo, hi)
        return out, ldj_out

    @staticmethod
    @once_differentiable
    def backward(ctx, gy, gl):
        x, = ctx.saved_tensors
        gx = None
        if ctx.needs_input_grad[0]:  # the data itself rarely needs a gradient
            gy, gl = L.dev(gy, 'grad x'), L.dev(gl, 'grad log_df_dz')
            gx = torch.empty_like(x)
            L.check(L.lib().nfb_logit_bwd(L.ptr(x), L.ptr(gy), L.ptr(gl), L.ptr(gx), ctx.meta[0], ctx.meta[1], x.size(0),
                                          x[0].numel(), L.stream()))
        return gx, gl, None, None


class Squeeze2dFn(Function):
    """squeeze.py:153-189: a permutation; the gradient is the inverse permutation of gy."""

    @staticmethod
    def forward(ctx, z, odd, unsqueeze):
        from .squeeze import squeeze2d_tensor, unsqueeze2d_tensor
        ctx.meta = (odd, unsqueeze)
        return unsqueeze2d_tensor(z, odd) if unsqueeze else squeeze2d_tensor(z, odd)

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        from .squeeze import squeeze2d_tensor, unsqueeze2d_tensor
        odd, unsqueeze = ctx.meta
        return (squeeze2d_tensor(gy, odd) if unsqueeze else unsqueeze2d_tensor(gy, odd)), None, None


class GaussNLLFn(Function):
    """main.py:85: per-sample -(log N(z; 0, I) + ldj); the (sum, count) fp64 payload is not differentiable."""

    @staticmethod
    def forward(ctx, z, ldj):
        z, ldj = L.dev(z, 'z'), L.dev(ldj, 'log_df_dz')
        B = z.size(0)
        rows = torch.empty(B, device=z.device, dtype=torch.float32)
        total = torch.empty(2, device=z.device, dtype=torch.float64)
        L.check(L.lib().nfb_gauss_nll(L.ptr(z), L.ptr(ldj), L.ptr(rows), total.data_ptr(), B, z[0].numel(), L.stream()))
        ctx.save_for_backward(z)
        ctx.mark_non_differentiable(total)
        return rows, total

    @staticmethod
    @once_differentiable
    def backward(ctx, grows, _gtotal):
        z, = ctx.saved_tensors
        grows = L.dev(grows, 'grad nll')
        gz, gl = torch.empty_like(z), torch.empty_like(grows)
        L.check(L.lib().nfb_gauss_nll_bwd(L.ptr(z), L.ptr(grows), L.ptr(gz), L.ptr(gl), z.size(0), z[0].numel(),
                                          L.stream()))
        return gz, gl
