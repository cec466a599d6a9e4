"""nfb200 -- B200-native (sm_100a) flow layers: forward / inverse + log|det J| of the couplings, ActNorm,
invertible 1x1 convolution, Logit and the squeeze permutations of tatsy/normalizing-flows-pytorch, as fused CUDA
kernels behind a C ABI (include/nfb200.h), wrapped in nn.Modules that keep the reference's layer API.

The directory is named ``normalizing-flows-pytorch_b200`` (not an importable identifier); ``import nfb200`` (the
shim package at the repo root) is the supported import name.
"""
from . import _lib, flows, likelihood, parallel  # noqa: F401
from .flows import Flowpp, Glow, RealNVP, set_throughput_mode  # noqa: F401
from .likelihood import bits_per_dim_from_total, gauss_nll  # noqa: F401
from ._lib import library_path_calls, library_path_log  # noqa: F401

__version__ = '0.1.0'
