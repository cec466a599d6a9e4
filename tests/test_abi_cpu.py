"""CPU tests of the drop-in boundary: libnfb200.so loads without a GPU and exports every symbol that
include/nfb200.h declares; the Python binding covers each of them; argument errors come back as negative codes;
the nn.Module mirror keeps the reference's state_dict layout; and there is no CPU fallback."""
import ctypes
import types

import pytest
import torch

from tests import _golden


def test_library_exports_every_header_symbol():
    import nfb200._lib as L
    lib = L.lib()
    names = L.header_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), n
        assert n in L._SIGNATURES, 'no python signature for ' + n
    assert lib.nfb_version() >= 100


def test_argument_errors_are_negative_codes_without_touching_the_gpu():
    import nfb200._lib as L
    lib = L.lib()
    assert lib.nfb_logit_fwd(None, None, None, None, 0.01, 0.99, 4, 8, None) == L.ERR_NULL
    fake = ctypes.c_void_p(256)  # never dereferenced: the shape check fails first
    assert lib.nfb_affine_coupling_fwd(fake, fake, fake, fake, fake, fake, fake, 4, 3, 5, 5, L.SPLIT_CHECKER, 0, None) == L.ERR_SPLIT
    assert lib.nfb_affine_coupling_fwd(fake, fake, fake, fake, fake, fake, fake, 4, 3, 4, 4, L.SPLIT_CHANNEL, 0, None) == L.ERR_SPLIT
    assert lib.nfb_affine_coupling_fwd(fake, fake, fake, fake, fake, fake, fake, 0, 3, 4, 4, L.SPLIT_CHECKER, 0, None) == L.ERR_SHAPE
    assert lib.nfb_mixlog_coupling_fwd(fake, fake, fake, fake, fake, fake, fake, 4, 4, 4, 4, L.SPLIT_CHANNEL, 0, 99, None) == L.ERR_UNSUPPORTED
    assert lib.nfb_squeeze2d(fake, fake, 2, 3, 3, 4, 0, None) == L.ERR_SPLIT
    assert b'split' in lib.nfb_error_string(L.ERR_SPLIT)
    with pytest.raises(RuntimeError):
        L.check(L.ERR_SHAPE)


def test_training_entry_points_validate_arguments_on_the_host():
    """The gradient / train-conditioner entry points reject bad arguments before any launch, and the scratch-size helpers
    are plain host arithmetic (no GPU needed)."""
    import nfb200._lib as L
    lib = L.lib()
    fake = ctypes.c_void_p(256)
    assert lib.nfb_affine_coupling_bwd(None, fake, fake, fake, fake, fake, fake, fake, fake, fake, fake, 4, 3, 4, 4,
                                       L.SPLIT_CHECKER, 0, None) == L.ERR_NULL
    assert lib.nfb_affine_coupling_bwd(fake, fake, fake, fake, fake, fake, fake, fake, fake, fake, fake, 4, 3, 5, 5,
                                       L.SPLIT_CHECKER, 0, None) == L.ERR_SPLIT
    assert lib.nfb_rqs_coupling_bwd(fake, fake, fake, fake, fake, fake, 4, 4, 1, 1, L.SPLIT_1D, 0, 1, 3.0, None) == L.ERR_SHAPE
    assert lib.nfb_mixlog_coupling_bwd(fake, fake, fake, fake, fake, fake, fake, fake, fake, fake, fake, 4, 4, 1, 1,
                                       L.SPLIT_1D, 0, 99, None) == L.ERR_UNSUPPORTED
    assert lib.nfb_conv_train(fake, fake, None, None, fake, None, 0, 4, 32, 32, 5, 5, 3, None) == L.ERR_UNSUPPORTED
    assert lib.nfb_conv_train(fake, fake, None, None, fake, None, 0, 4, 32, 32, 16, 16, 2, None) == L.ERR_SHAPE
    assert lib.nfb_rows_to_planes(fake, fake, 100, 8, 64, None) == L.ERR_SHAPE  # 100 rows are not whole planes of 64
    assert lib.nfb_wn_pack_train_multi(fake, fake, 0, 1e-5, None) == L.ERR_SHAPE
    # sample groups x (|gw| + |gb|): B = 256 at 16x16 -> 2 samples per CTA -> 128 groups
    assert lib.nfb_conv_train_wgrad_scratch(256, 32, 32, 16, 16, 3) == 128 * (32 * 32 * 9 + 32)
    # 32x32: 4 bands per sample = 1024 units, 4 per CTA (about two CTAs per SM in total) -> 256 groups
    assert lib.nfb_conv_train_wgrad_scratch(256, 32, 32, 32, 32, 3) == 256 * (32 * 32 * 9 + 32)
    assert lib.nfb_invconv1x1_wgrad_scratch(256, 48, 64) == 128 * 48 * 48
    ptrs = (ctypes.c_void_p * 15)(*[256] * 15)
    assert lib.nfb_flowpp_mlp_fwd(ptrs, fake, fake, 8, 4, L.SPLIT_1D, 0, 2, 100000, None) == L.ERR_UNSUPPORTED
    assert lib.nfb_flowpp_mlp_fwd(ptrs, fake, fake, 8, 5, L.SPLIT_1D, 0, 2, 14, None) == L.ERR_SHAPE


def test_no_cpu_fallback():
    import nfb200
    with pytest.raises(RuntimeError, match='CUDA'):
        nfb200.flows.ActNorm((4, )).forward(torch.randn(2, 4), torch.zeros(2))
    net = nfb200.RealNVP((2, ), None, types.SimpleNamespace(layers=2))
    with pytest.raises(RuntimeError, match='CUDA'):
        net(torch.randn(3, 2))


@pytest.mark.parametrize('name', _golden.names('model_'))
def test_state_dict_layout_matches_reference(name):
    """Keys, shapes and dtypes equal the reference's (goldens hold its state_dict) -> its checkpoints load unchanged."""
    import nfb200
    meta, _, sd = _golden.load(name)
    cls = {'RealNVP': nfb200.RealNVP, 'Glow': nfb200.Glow, 'Flowpp': nfb200.Flowpp}[meta['kind']]
    net = cls(tuple(meta['dims']), meta['datatype'], types.SimpleNamespace(layers=meta['layers'], mixtures=meta['mixtures']))
    mine = net.state_dict()
    assert list(mine.keys()) == list(sd.keys())  # same order too
    for k in sd:
        assert tuple(mine[k].shape) == tuple(sd[k].shape), k
        assert mine[k].dtype == sd[k].dtype, k
    net.load_state_dict(sd)  # strict


def test_layer_signatures_match_reference():
    import inspect
    import nfb200.flows as F
    sig = lambda c: list(inspect.signature(c.__init__).parameters)[1:]  # noqa: E731
    assert sig(F.AffineCoupling) == ['dims', 'masking', 'odd']
    assert sig(F.MixLogAttnCoupling) == ['dims', 'masking', 'odd', 'base_filters', 'n_mixtures']
    assert sig(F.ActNorm) == ['num_features', 'eps']
    assert sig(F.InvertibleConv1x1) == ['in_out_channels']
    assert sig(F.Logit) == ['eps']
    assert sig(F.BatchNorm) == ['num_features', 'momentum', 'eps', 'affine']
    assert sig(F.Glow) == ['dims', 'datatype', 'cfg']
    for cls in (F.AffineCoupling, F.ActNorm, F.InvertibleConv1x1, F.Logit, F.BatchNorm, F.Squeeze2d, F.Compose):
        assert list(inspect.signature(cls.forward).parameters)[1:] in (['z', 'log_df_dz'], ['x', 'log_df_dz'], ['x', 'log_det_jacob'])
        assert hasattr(cls, 'backward') and hasattr(cls, 'inverse')
    with pytest.raises(Exception, match='unsupported combination'):
        F.AffineCoupling((3, 4), masking='checkerboard')  # coupling.py:29-30
