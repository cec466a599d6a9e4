"""Index permutations of the reference's ``flows/squeeze.py`` on the CUDA path (bit-exact).

``Squeeze2d`` / ``Unsqueeze2d`` are flow layers (``forward(z, log_df_dz)`` / ``backward``); the functional
split/merge pairs are what ``AbstractCoupling`` uses in the reference (coupling.py:19-27).  The fused coupling
kernels never materialise the halves; these functions exist for API parity and for the conditioner input.
"""
import torch
import torch.nn as nn

from .. import _lib as L
from . import autograd as G


def _dims4(z):
    if z.dim() == 2:
        return z.size(0), z.size(1), 1, 1
    assert z.dim() == 4
    return tuple(z.shape)


def coupling_split(z, mode, odd=False, want_z0=True, want_z1=True):
    """(z0, z1) of the coupling split `mode`; either half may be skipped (returned as None)."""
    z = L.dev(z, 'z')
    B, C, H, W = _dims4(z)
    if mode == L.SPLIT_CHECKER:
        shape = (B, 2 * C, H // 2, W // 2)
    elif mode == L.SPLIT_CHANNEL:
        shape = (B, C // 2, H, W)
    else:
        shape = (B, C // 2)
    z0 = torch.empty(shape, device=z.device, dtype=z.dtype) if want_z0 else None
    z1 = torch.empty(shape, device=z.device, dtype=z.dtype) if want_z1 else None
    L.check(L.lib().nfb_coupling_split(L.ptr(z), L.ptr(z0) if want_z0 else None, L.ptr(z1) if want_z1 else None, B, C,
                                       H, W, mode, int(bool(odd)), L.stream()))
    return z0, z1


def coupling_merge(z0, z1, mode, odd=False):
    z0, z1 = L.dev(z0, 'z0'), L.dev(z1, 'z1')
    if mode == L.SPLIT_CHECKER:
        B, C2, h, w = z0.shape
        C, H, W = C2 // 2, 2 * h, 2 * w
        out = torch.empty((B, C, H, W), device=z0.device, dtype=z0.dtype)
    elif mode == L.SPLIT_CHANNEL:
        B, Ch, H, W = z0.shape
        C = 2 * Ch
        out = torch.empty((B, C, H, W), device=z0.device, dtype=z0.dtype)
    else:
        B, Ch = z0.shape
        C, H, W = 2 * Ch, 1, 1
        out = torch.empty((B, C), device=z0.device, dtype=z0.dtype)
    L.check(L.lib().nfb_coupling_merge(L.ptr(z0), L.ptr(z1), L.ptr(out), B, C, H, W, mode, int(bool(odd)), L.stream()))
    return out


# functional names of the reference (squeeze.py:5-83)
def channel_split(z, dim=1, odd=False):
    assert dim == 1
    return coupling_split(z, L.SPLIT_CHANNEL, odd)


def channel_merge(z0, z1, dim=1, odd=False):
    assert dim == 1
    return coupling_merge(z0, z1, L.SPLIT_CHANNEL, odd)


def checker_split(z, odd=False):
    return coupling_split(z, L.SPLIT_CHECKER, odd)


def checker_merge(z0, z1, odd=False):
    return coupling_merge(z0, z1, L.SPLIT_CHECKER, odd)


def squeeze1d(z, odd=False):
    return coupling_split(z, L.SPLIT_1D, odd)


def unsqueeze1d(z0, z1, odd=False):
    return coupling_merge(z0, z1, L.SPLIT_1D, odd)


def squeeze2d_tensor(z, odd=False):
    """(B,C,H,W) -> (B,4C,H/2,W/2), channel k = 4c + 2dy + dx (squeeze.py:86-97 + cat of Squeeze2d.forward)."""
    z = L.dev(z, 'z')
    B, C, H, W = z.shape
    out = torch.empty((B, 4 * C, H // 2, W // 2), device=z.device, dtype=z.dtype)
    L.check(L.lib().nfb_squeeze2d(L.ptr(z), L.ptr(out), B, C, H, W, int(bool(odd)), L.stream()))
    return out


def unsqueeze2d_tensor(z, odd=False):
    z = L.dev(z, 'z')
    B, C4, h, w = z.shape
    C, H, W = C4 // 4, 2 * h, 2 * w
    out = torch.empty((B, C, H, W), device=z.device, dtype=z.dtype)
    L.check(L.lib().nfb_unsqueeze2d(L.ptr(z), L.ptr(out), B, C, H, W, int(bool(odd)), L.stream()))
    return out


class Squeeze2d(nn.Module):
    """squeeze.py:153-170."""

    def __init__(self, odd=False):
        super().__init__()
        self.odd = odd

    def forward(self, z, log_df_dz):
        if G.needs_grad(z):
            return G.Squeeze2dFn.apply(z, self.odd, False), log_df_dz
        return squeeze2d_tensor(z, self.odd), log_df_dz

    def backward(self, z, log_df_dz):
        return unsqueeze2d_tensor(z, self.odd), log_df_dz

    inverse = backward


class Unsqueeze2d(nn.Module):
    """squeeze.py:173-189."""

    def __init__(self, odd=False):
        super().__init__()
        self.odd = odd

    def forward(self, z, log_df_dz):
        if G.needs_grad(z):
            return G.Squeeze2dFn.apply(z, self.odd, True), log_df_dz
        return unsqueeze2d_tensor(z, self.odd), log_df_dz

    def backward(self, z, log_df_dz):
        return squeeze2d_tensor(z, self.odd), log_df_dz

    inverse = backward
