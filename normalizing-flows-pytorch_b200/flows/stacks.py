"""Flow stacks: the layer recipes of the reference's ``glow.py:17-60``, ``flowpp.py:17-70`` and
``realnvp.py:17-55`` replayed from one table, over the libnfb200 layers.

``Glow`` / ``Flowpp`` / ``RealNVP`` keep the reference's constructor ``(dims, datatype=None, cfg=None)`` (cfg needs
``.layers`` and, for Flow++, ``.mixtures``), ``forward(z) -> (z, log_df_dz)`` and ``backward(z)``; ``state_dict`` keys
are identical (``net.layers.<i>....``).  Additions: ``inverse`` alias, ``nll`` / ``bits_per_dim`` (main.py:85).
"""
import torch
import torch.nn as nn

from .. import _lib as L
from ..likelihood import gauss_nll
from .coupling import AffineCoupling, MixLogAttnCoupling, RQSplineCoupling
from .modules import ActNorm, BatchNorm, Compose, InvertibleConv1x1, Logit
from .squeeze import Squeeze2d, Unsqueeze2d


def _multiscale_plan(dims, n_layers):
    """Yield (dims, masking, odd) / 'squeeze' / 'unsqueeze' in the order of glow.py:22-51."""
    d = tuple(dims)
    n_sq = 0
    while max(d[1], d[2]) > 8:
        for i in range(n_layers):
            yield d, 'checkerboard', i % 2 != 0
        yield 'squeeze'
        n_sq += 1
        d = (d[0] * 4, d[1] // 2, d[2] // 2)
        for i in range(n_layers):
            yield d, 'channelwise', i % 2 != 0
    for i in range(n_layers + 1):
        yield d, 'checkerboard', i % 2 != 0
    for _ in range(n_sq):
        yield 'unsqueeze'


class _FlowStack(nn.Module):
    def __init__(self, dims, datatype=None, cfg=None):
        super().__init__()
        self.dims = tuple(dims)
        self.n_layers = cfg.layers
        self._cfg = cfg
        layers = []
        if datatype == 'image':
            layers.append(Logit(eps=0.01))
            for item in _multiscale_plan(self.dims, self.n_layers):
                if item == 'squeeze':
                    layers.append(Squeeze2d(odd=False))
                elif item == 'unsqueeze':
                    layers.append(Unsqueeze2d(odd=False))
                else:
                    layers.extend(self._step(*item, image=True))
        else:
            for i in range(self.n_layers):
                layers.extend(self._step(self.dims, 'checkerboard', i % 2 != 0, image=False))
        self.net = Compose(layers)

    def _step(self, dims, masking, odd, image):
        raise NotImplementedError

    def forward(self, z):
        z = L.dev(z, 'z')
        log_df_dz = torch.zeros(z.size(0), device=z.device, dtype=z.dtype)
        return self.net(z, log_df_dz)

    def backward(self, z):
        z = L.dev(z, 'z')
        log_df_dz = torch.zeros(z.size(0), device=z.device, dtype=z.dtype)
        return self.net.backward(z, log_df_dz)

    inverse = backward

    # ---- likelihood helpers (main.py:83-85, 121-124) ----
    def nll(self, y):
        """-> (per-sample NLL float[B], device double[2] = (sum NLL, B))."""
        z, ldj = self.forward(y)
        return gauss_nll(z, ldj)

    def bits_per_dim(self, y):
        _, tot = self.nll(y)
        s, n = tot.tolist()
        return s / n / (float(torch.tensor(self.dims).prod()) * 0.6931471805599453)

    def mark_initialized(self, flag=True):
        """ActNorm.initialized is not part of the state dict: call after load_state_dict."""
        for m in self.modules():
            if isinstance(m, ActNorm):
                m.initialized = flag
        return self


class Glow(_FlowStack):
    """glow.py:10-68."""

    def _step(self, dims, masking, odd, image):
        return [ActNorm(dims), InvertibleConv1x1(dims[0]), AffineCoupling(dims, masking=masking, odd=odd)]


class Flowpp(_FlowStack):
    """flowpp.py:9-78 (no 1x1 conv in the density-sample branch, flowpp.py:64-66)."""

    def _step(self, dims, masking, odd, image):
        cpl = MixLogAttnCoupling(dims, masking=masking, odd=odd, n_mixtures=self._cfg.mixtures)
        if image:
            return [ActNorm(dims), InvertibleConv1x1(dims[0]), cpl]
        return [ActNorm(dims), cpl]


class RealNVP(_FlowStack):
    """realnvp.py:9-63.  ``cfg.coupling = 'rqs'`` swaps the affine bijection for the RQ-spline one
    (BASELINE.json config 4; not in the reference)."""

    def _step(self, dims, masking, odd, image):
        kind = getattr(self._cfg, 'coupling', 'affine')
        if kind == 'rqs':
            cpl = RQSplineCoupling(dims, masking=masking, odd=odd, n_bins=getattr(self._cfg, 'bins', 8),
                                   tail_bound=getattr(self._cfg, 'tail_bound', 3.0))
        else:
            cpl = AffineCoupling(dims, masking=masking, odd=odd)
        return [BatchNorm(dims, affine=False), cpl]
