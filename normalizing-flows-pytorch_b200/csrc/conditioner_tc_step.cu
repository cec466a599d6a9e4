// conditioner_tc_step.cu -- the tensor-core conditioner + coupling kernel with the NEXT flow step's ActNorm + invertible 1x1
// convolution as a post-op (glow.py:27-29: one launch per flow step).  Kernel: conditioner_tc.cuh.
#include "conditioner_tc.cuh"

namespace nfb {

// The (conditioner map, channel count of z, split) combinations of the Glow stacks (glow.py:17-60 on 32x32 and 64x64 images);
// anything else returns NFB_ERR_UNSUPPORTED and the caller launches nfb_actnorm_invconv_fwd separately.  Why it pays although
// the per-pixel matrix product runs on the 256 epilogue threads of a few CTAs instead of the whole machine: with several
// batches in flight the cost of a kernel is the SM-time it occupies, and the separate ActNorm + 1x1-conv launches (161 per
// Glow-32 step, ~4-10 us each on ALL SMs for ~1 us of work) were a quarter of the step's SM-time (ncu sm__cycles_active).
template <int MODE>
static int tc_step_by_size(float* z, float* ldj, const float* pk_tc, const SplitGeom& g, int Cin, int Cout, int B, const float* sa,
                           const float* sb, int flags, cudaStream_t st, const PostOp& post) {
    const int h = g.h, w = g.w, C = g.C;
    if (flags & NFB_CONV_TF32) return NFB_ERR_UNSUPPORTED;  // the step kernels exist for the default operand format only
    const long long tiles = (static_cast<long long>(B) * h * w + 127) / 128;
    const bool pair = (flags & NFB_CONV_PAIR) ? tiles >= 2 : tiles >= 2 * kSMs;
    const bool single = (flags & NFB_CONV_SINGLE) != 0;
    int rc;
#define NFB_STEP(H_, W_, PAIR_, CP_, DUAL_)                                                                                          \
    {                                                                                                                              \
        rc = launch_tc<H_, W_, MODE, true, PAIR_, CP_, true, DUAL_>(z, z, ldj, pk_tc, g, Cin, Cout, B, sa, sb, flags, st, post);     \
        if (rc != NFB_ERR_UNSUPPORTED) return rc;                                                                                  \
    }
    if (h == 16 && w == 16) {
        const bool dual = ((flags & NFB_CONV_PAIR) ? B >= 2 : B > kSMs) && !single;
        if (MODE == NFB_SPLIT_CHECKER && C == 3) {  // first level of a 32x32 image
            if (dual) NFB_STEP(16, 16, false, 3, true);
            NFB_STEP(16, 16, false, 3, false);
        }
        if (MODE == NFB_SPLIT_CHANNEL && C == 12) {
            if (dual) NFB_STEP(16, 16, false, 12, true);
            NFB_STEP(16, 16, false, 12, false);
        }
    } else if (h == 8 && w == 8) {
        const bool dual = pair && tiles >= 4 && !single;
        if (MODE == NFB_SPLIT_CHECKER && C == 12) {
            if (dual) NFB_STEP(8, 8, true, 12, true);
            if (pair) NFB_STEP(8, 8, true, 12, false);
            NFB_STEP(8, 8, false, 12, false);
        }
        if (MODE == NFB_SPLIT_CHANNEL && C == 48) {
            if (dual) NFB_STEP(8, 8, true, 48, true);
            if (pair) NFB_STEP(8, 8, true, 48, false);
            NFB_STEP(8, 8, false, 48, false);
        }
    } else if (h == 4 && w == 4) {
        if (MODE == NFB_SPLIT_CHECKER && C == 48) {
            if (pair) NFB_STEP(4, 4, true, 48, false);
            NFB_STEP(4, 4, false, 48, false);
        }
    }
#undef NFB_STEP
    return NFB_ERR_UNSUPPORTED;
}

int convnet_affine_step_tc_dispatch(float* z, float* ldj, const float* pk_tc, const SplitGeom& g, int mode, int Cin, int Cout,
                                    int B, const float* sa, const float* sb, const float* an_ls, const float* an_b,
                                    const float* Wm, const float* log_s, int flags, cudaStream_t st) {
    if (Cout != 2 * g.c0) return NFB_ERR_SHAPE;
    const PostOp post{an_ls, an_b, Wm, log_s};
    if (mode == NFB_SPLIT_CHECKER) return tc_step_by_size<NFB_SPLIT_CHECKER>(z, ldj, pk_tc, g, Cin, Cout, B, sa, sb, flags, st, post);
    if (mode == NFB_SPLIT_CHANNEL) return tc_step_by_size<NFB_SPLIT_CHANNEL>(z, ldj, pk_tc, g, Cin, Cout, B, sa, sb, flags, st, post);
    return NFB_ERR_UNSUPPORTED;
}


}  // namespace nfb
