"""Host-side consistency of the kernel-selection flags: include/nfb200.h (the C ABI) and nfb200._lib (the ctypes binding)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_defines():
    txt = open(os.path.join(ROOT, 'include', 'nfb200.h')).read()
    out = {}
    for name, val in re.findall(r'^#define\s+(NFB_CONV_[A-Z0-9_]+)\s+(0x[0-9a-fA-F]+|\d+)\b', txt, re.M):
        out[name] = int(val, 0)
    return out


def test_flag_values_match_the_binding():
    import nfb200._lib as L
    d = header_defines()
    assert d['NFB_CONV_FFMA'] == L.CONV_FFMA
    assert d['NFB_CONV_PAIR'] == L.CONV_PAIR
    assert d['NFB_CONV_TF32'] == L.CONV_TF32
    assert d['NFB_CONV_SINGLE'] == L.CONV_SINGLE
    assert L.conv_groups(3) == 3 << d['NFB_CONV_GROUPS_SHIFT']
    assert L.conv_debug(5) == 5 << d['NFB_CONV_DEBUG_SHIFT']
    assert L.conv_iters(2) == 2 << d['NFB_CONV_ITERS_SHIFT']
    assert L.conv_variant(4) == 4 and d['NFB_CONV_VARIANT_MASK'] == 7


def test_flags_do_not_overlap():
    d = header_defines()
    fields = [d['NFB_CONV_VARIANT_MASK'], d['NFB_CONV_FFMA'], 7 << d['NFB_CONV_GROUPS_SHIFT'], d['NFB_CONV_PAIR'],
              0xff << d['NFB_CONV_DEBUG_SHIFT'], d['NFB_CONV_TF32'], d['NFB_CONV_SINGLE'], 3 << d['NFB_CONV_ITERS_SHIFT']]
    seen = 0
    for f in fields:
        assert seen & f == 0, hex(f)
        seen |= f


def test_pack_size_covers_both_tensor_core_sections():
    """nfb_resnet_pack_size = FFMA section + 3xTF32 section + FP16-split section (host arithmetic only, no GPU needed)."""
    import nfb200._lib as L
    lib = L.lib()
    for cin, cout in ((6, 12), (24, 48), (96, 192), (3, 7), (40, 100)):
        n_conv = lib.nfb_resnet_pack_size(cin, cout, 1)
        n_mlp = lib.nfb_resnet_pack_size(cin, cout, 0)
        assert n_conv > 0 and n_mlp > 0
        # four 32->32 3x3 layers alone: 4 * 9 * 32 * 32 weights, kept once (FFMA) + twice as hi/lo (TF32, 4 B) + twice as
        # hi/lo halves (FP16, 2 B) = at least 1 + 2 + 1 copies
        assert n_conv >= 4 * 4 * 9 * 32 * 32
