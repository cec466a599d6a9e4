"""Drop-in mirror of the reference's ``flows`` package for the coupling / Glow / Flow++ / RealNVP hot path."""
from .coupling import (AbstractCoupling, AdditiveCoupling, AffineCoupling, MixLogAttnCoupling, RQSplineCoupling)
from .conditioner import MLP, ConvNet, GatedAttn, GatedConv2d, GatedLinear, set_throughput_mode
from .modules import ActNorm, BatchNorm, Compose, InvertibleConv1x1, Logit, MixLogCDF
from .squeeze import (Squeeze2d, Unsqueeze2d, channel_merge, channel_split, checker_merge, checker_split, squeeze1d,
                      unsqueeze1d)
from .stacks import Flowpp, Glow, RealNVP
from .weight_norm import WeightNorm

__all__ = [
    'RealNVP', 'Glow', 'Flowpp', 'AbstractCoupling', 'AdditiveCoupling', 'AffineCoupling', 'MixLogAttnCoupling',
    'RQSplineCoupling', 'MLP', 'ConvNet', 'GatedAttn', 'GatedConv2d', 'GatedLinear', 'ActNorm', 'BatchNorm', 'Compose',
    'InvertibleConv1x1', 'Logit', 'MixLogCDF', 'Squeeze2d', 'Unsqueeze2d', 'WeightNorm', 'channel_split', 'channel_merge',
    'checker_split', 'checker_merge', 'squeeze1d', 'unsqueeze1d', 'set_throughput_mode'
]
