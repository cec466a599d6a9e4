// conditioner_tc.cu -- packing of the tensor-core weight sections and the dispatch of the conditioner (+ coupling) kernels of
// conditioner_tc.cuh.  The variants that also apply the next flow step's ActNorm + 1x1 convolution are instantiated in
// conditioner_tc_step.cu (a second translation unit: the two compile in parallel).
#include "conditioner_tc.cuh"

namespace nfb {

// =====================================================================================================================
// packing of the tensor-core section from the FFMA section (once per weight update)
// =====================================================================================================================
__global__ void __launch_bounds__(256) pack_tc_kernel(const float* __restrict__ pk, float* __restrict__ tc, int Cin, int Cout,
                                                      int f16) {
    const PackLayout L = pack_layout(Cin, Cout, 9);
    const TcPlan P = tc_plan(Cin, Cout, f16);
    const int c0 = Cout / 2;
    const int NJ = f16 ? 2 : 4, KS = f16 ? 16 : 8;   // k-steps per tap of a 32-channel layer, channels per k-step
    const int CPW = f16 ? 2 : 1;                     // input channels per 4-byte word
    const int stage = 9 * NJ * 512;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < P.total; idx += gridDim.x * blockDim.x) {
        float val[2] = {0.f, 0.f};
        bool split = true, lo = false;
        if (idx < P.obias_g) {
            // constants: b0 | blk0: sA tA b1' b2 | blk1: sA tA b1' b2 | sO tO
            const int k = idx - P.consts, c = k & 31;
            if (k < 32) val[0] = pk[L.b0 + c];
            else if (k < 288) {
                const int blk = (k - 32) / 128, which = ((k - 32) % 128) / 32;
                val[0] = which == 0 ? pk[L.bnA[blk] + c] : which == 1 ? pk[L.bnA[blk] + kF + c]
                         : which == 2 ? pk[L.b1[blk] + c] : pk[L.b2[blk] + c];
            } else if (k < 320) val[0] = pk[L.bnO + c];
            else if (k < 352) val[0] = pk[L.bnO + kF + c];
            split = false;
        } else if (idx < P.in0) {
            // output bias by column: generic order, then fused order
            const bool fused = idx >= P.obias_f;
            const int k = idx - (fused ? P.obias_f : P.obias_g);
            const int NW = fused ? P.NWf : P.NWg;
            const int oc = tc_out_channel(fused, k / NW, k % NW, NW, c0, Cout);
            val[0] = oc >= 0 ? pk[L.bout + oc] : 0.f;
            split = false;
        } else if (idx < P.outg) {
            // 3x3 stages: [tap][j][k-chunk 2][n (64: w_hi | w_lo)][16 bytes: 4 TF32 / 8 FP16 input channels]
            int e = idx - P.in0, cig_base, base, nj;
            if (idx < P.mid0) {
                const int pass = e / stage;
                e -= pass * stage;
                nj = (pass == P.n_in - 1) ? P.nj_last : NJ;
                cig_base = pass * kF;
                base = -1;
            } else {
                e = idx - P.mid0;
                const int i = e / stage;
                e -= i * stage;
                nj = NJ;
                cig_base = 0;
                base = (i & 1) ? L.w2[i >> 1] : L.w1[i >> 1];
            }
            const int ks = e / 512, r = e % 512;           // k-step (tap*nj + j), 512 words each
            const int tap = ks / nj, j = ks % nj;
            const int kch = r / 256, n = (r % 256) / 4, q = r % 4;
            const int co = n & 31;
            lo = n >= 32;
            for (int h = 0; h < CPW; ++h) {
                const int ci = cig_base + KS * j + (KS / 2) * kch + CPW * q + h;
                if (base < 0) val[h] = ci < Cin ? pk[L.w0 + (ci * 9 + tap) * kF + co] : 0.f;
                else val[h] = pk[base + (ci * 9 + tap) * kF + co];
            }
        } else {
            // output chunks: [j][k-chunk 2][n 2NW: hi | lo][16 bytes]
            const bool fused = idx >= P.outf;
            const int NW = fused ? P.NWf : P.NWg;
            const int e = idx - (fused ? P.outf : P.outg);
            const int q_chunk = e / (16 * NJ * NW), r = e % (16 * NJ * NW);
            const int blkk = r / (8 * NW), r2 = r % (8 * NW);  // block (j*2 + k-chunk) of 2NW rows x 4 words
            const int n = r2 / 4, q = r2 % 4;
            lo = n >= NW;
            const int oc = tc_out_channel(fused, q_chunk, lo ? n - NW : n, NW, c0, Cout);
            for (int h = 0; h < CPW; ++h) {
                const int ci = (KS / 2) * blkk + CPW * q + h;
                val[h] = oc >= 0 ? pk[L.wout + ((oc >> 5) * kF + ci) * kF + (oc & 31)] : 0.f;  // pack_wn layout, J = 32
            }
        }
        float out = val[0];
        if (split && f16) {
            uint32_t hi, l;
            // a weight outside the fp16 range cannot be split: NaN makes every output of the network NaN (loud), see above
            if (!(fabsf(val[0]) < 65504.f)) val[0] = __int_as_float(0x7fc00000);
            if (!(fabsf(val[1]) < 65504.f)) val[1] = __int_as_float(0x7fc00000);
            split_f16_pair(val[0], val[1], hi, l);
            out = __uint_as_float(lo ? l : hi);
        } else if (split) {
            float hi, l;
            split_tf32(val[0], hi, l);
            out = lo ? l : hi;
        }
        tc[idx] = out;
    }
}

int pack_tc_launch(const float* pk_ffma, float* pk_tc_section, int Cin, int Cout, int f16, cudaStream_t st) {
    const TcPlan P = tc_plan(Cin, Cout, f16);
    int blocks = (P.total + 255) / 256;
    if (blocks > kSMs * 8) blocks = kSMs * 8;
    pack_tc_kernel<<<blocks, 256, 0, st>>>(pk_ffma, pk_tc_section, Cin, Cout, f16);
    return launch_status();
}

template <int MODE, bool FUSED, bool F16>
static int tc_by_size_p(const float* zsrc, float* zdst, float* ldj, const float* pk_tc, const SplitGeom& g, int Cin, int Cout, int B,
                        int h, int w, const float* sa, const float* sb, int flags, cudaStream_t st) {
    if (h == 16 && w == 16) {
        // two samples in flight per CTA (FP16 split only) when asked for (NFB_CONV_PAIR: throughput mode) or when the batch
        // gives every SM more than one sample anyway; output layers wider than 64 columns per chunk stay on the one-unit kernel
        const bool dual = F16 && ((flags & NFB_CONV_PAIR) ? B >= 2 : B > kSMs) && !(flags & NFB_CONV_SINGLE);
        if (dual) {
            const int rc = launch_tc<16, 16, MODE, FUSED, false, 0, F16, F16>(zsrc, zdst, ldj, pk_tc, g, Cin, Cout, B, sa, sb, flags, st);
            if (rc != NFB_ERR_UNSUPPORTED) return rc;
        }
        return launch_tc<16, 16, MODE, FUSED, false, 0, F16>(zsrc, zdst, ldj, pk_tc, g, Cin, Cout, B, sa, sb, flags, st);
    }
    // maps of <= 128 pixels: two tiles per CTA when asked for (NFB_CONV_PAIR: several batches in flight) or when the batch
    // alone gives every SM at least two single-tile units
    const long long tiles = (static_cast<long long>(B) * h * w + 127) / 128;
    const bool pair = (flags & NFB_CONV_PAIR) ? tiles >= 2 : tiles >= 2 * kSMs;
    if (h == 8 && w == 8) {
        if (pair && F16 && tiles >= 4 && !(flags & NFB_CONV_SINGLE)) {  // two two-tile units in flight per CTA
            const int rc = launch_tc<8, 8, MODE, FUSED, true, 0, F16, F16>(zsrc, zdst, ldj, pk_tc, g, Cin, Cout, B, sa, sb, flags, st);
            if (rc != NFB_ERR_UNSUPPORTED) return rc;
        }
        if (pair) return launch_tc<8, 8, MODE, FUSED, true, 0, F16>(zsrc, zdst, ldj, pk_tc, g, Cin, Cout, B, sa, sb, flags, st);
        return launch_tc<8, 8, MODE, FUSED, false, 0, F16>(zsrc, zdst, ldj, pk_tc, g, Cin, Cout, B, sa, sb, flags, st);
    }
    if (h == 4 && w == 4) {
        if (pair) return launch_tc<4, 4, MODE, FUSED, true, 0, F16>(zsrc, zdst, ldj, pk_tc, g, Cin, Cout, B, sa, sb, flags, st);
        return launch_tc<4, 4, MODE, FUSED, false, 0, F16>(zsrc, zdst, ldj, pk_tc, g, Cin, Cout, B, sa, sb, flags, st);
    }
    return NFB_ERR_UNSUPPORTED;
}
template <int MODE, bool FUSED>
static int tc_by_size(const float* zsrc, float* zdst, float* ldj, const float* pk_tc, const SplitGeom& g, int Cin, int Cout, int B,
                      int h, int w, const float* sa, const float* sb, int flags, cudaStream_t st) {
    if (!(flags & NFB_CONV_TF32)) return tc_by_size_p<MODE, FUSED, true>(zsrc, zdst, ldj, pk_tc, g, Cin, Cout, B, h, w, sa, sb, flags, st);
    return tc_by_size_p<MODE, FUSED, false>(zsrc, zdst, ldj, pk_tc, g, Cin, Cout, B, h, w, sa, sb, flags, st);
}

// conditioner + coupling + the next step's ActNorm / 1x1 conv: the (map size, channel count) pairs of the Glow stacks
int convnet_tc_dispatch(const float* zsrc, float* out, const float* pk_tc, const SplitGeom& g, int mode, int Cin, int Cout,
                        int B, int h, int w, int flags, cudaStream_t st) {
    if (mode == NFB_SPLIT_CHECKER)
        return tc_by_size<NFB_SPLIT_CHECKER, false>(zsrc, out, nullptr, pk_tc, g, Cin, Cout, B, h, w, nullptr, nullptr, flags, st);
    if (mode == NFB_SPLIT_CHANNEL)
        return tc_by_size<NFB_SPLIT_CHANNEL, false>(zsrc, out, nullptr, pk_tc, g, Cin, Cout, B, h, w, nullptr, nullptr, flags, st);
    return tc_by_size<-1, false>(zsrc, out, nullptr, pk_tc, g, Cin, Cout, B, h, w, nullptr, nullptr, flags, st);
}

int convnet_affine_tc_dispatch(float* z, float* ldj, const float* pk_tc, const SplitGeom& g, int mode, int Cin, int Cout, int B,
                               const float* sa, const float* sb, int flags, cudaStream_t st) {
    if (Cout != 2 * g.c0) return NFB_ERR_SHAPE;
    if (mode == NFB_SPLIT_CHECKER)
        return tc_by_size<NFB_SPLIT_CHECKER, true>(z, z, ldj, pk_tc, g, Cin, Cout, B, g.h, g.w, sa, sb, flags, st);
    if (mode == NFB_SPLIT_CHANNEL)
        return tc_by_size<NFB_SPLIT_CHANNEL, true>(z, z, ldj, pk_tc, g, Cin, Cout, B, g.h, g.w, sa, sb, flags, st);
    return NFB_ERR_UNSUPPORTED;
}

}  // namespace nfb

// developer entry point (not part of include/nfb200.h): device buffer of 3 * 512 uint64 for CTA 0's timeline, or NULL
extern "C" int nfb_debug_timeline(unsigned long long* buf) {
    return static_cast<int>(cudaMemcpyToSymbol(nfb::g_tl_buf, &buf, sizeof(buf)));
}

