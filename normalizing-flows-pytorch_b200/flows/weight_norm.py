"""WeightNorm parameter holder with the reference's key layout (weight_norm.py:5-45).

``<name>.module.{weight_g, weight_v, bias}``; the norm is taken over dim 0 (output channels) with eps in the
denominator.  The folded weight is produced by ``nfb_weight_norm`` and cached until g or v change -- the reference
recomputes it on every forward (weight_norm.py:43-45), which is constant work in eval mode.
"""
import torch
import torch.nn as nn

from .. import _lib as L


class WeightNorm(nn.Module):
    def __init__(self, module, eps=1.0e-5):
        super().__init__()
        w = module.weight.detach()
        g = torch.norm(w, dim=0)  # weight_norm.py:21
        v = w / (g.expand_as(w) + eps)  # weight_norm.py:22
        self.eps = eps
        self.kernel_size = tuple(w.shape[2:])
        holder = nn.Module()
        holder.bias = nn.Parameter(module.bias.detach().clone())
        holder.weight_g = nn.Parameter(g)
        holder.weight_v = nn.Parameter(v)
        self.module = holder
        self._key = None
        self._w = None

    @property
    def bias(self):
        return self.module.bias

    def weight(self):
        """Folded weight w = v * g / (||v|| + eps) as a device tensor shaped like v."""
        g, v = self.module.weight_g, self.module.weight_v
        key = L.param_key((g, v))
        if key != self._key:
            L.dev(v.data, 'weight_v')
            if self._w is None or self._w.device != v.device:
                self._w = torch.empty_like(v.data)
            O = v.size(0)
            L.check(L.lib().nfb_weight_norm(L.ptr(v.data), L.ptr(g.data), L.ptr(self._w), O, v[0].numel(),
                                            float(self.eps), L.stream()))
            self._key = key
        return self._w
