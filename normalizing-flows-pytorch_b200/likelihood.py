"""Gaussian NLL of the flow output (main.py:49-51, 83-85) as one reduction on the device."""
import math

import torch

from . import _lib as L

LN2 = math.log(2.0)


def gauss_nll(z, ldj):
    """-> (nll_rows float[B], total double[2] = (sum_b nll_b, B)); nll_b = -(log N(z_b; 0, I) + ldj_b)."""
    if torch.is_grad_enabled() and (z.requires_grad or ldj.requires_grad):
        from .flows.autograd import GaussNLLFn
        return GaussNLLFn.apply(z, ldj)  # differentiable rows (main.py:85: loss = rows.mean())
    z, ldj = L.dev(z, 'z'), L.dev(ldj, 'log_df_dz')
    B = z.size(0)
    rows = torch.empty(B, device=z.device, dtype=torch.float32)
    total = torch.empty(2, device=z.device, dtype=torch.float64)
    L.check(L.lib().nfb_gauss_nll(L.ptr(z), L.ptr(ldj), L.ptr(rows), L.ptr(total), B, z[0].numel(), L.stream()))
    return rows, total


def bits_per_dim_from_total(total, D):
    """total = (sum NLL, count) as returned by gauss_nll (after any cross-rank all-reduce)."""
    s, n = (float(v) for v in total.tolist())
    return s / n / (D * LN2)
