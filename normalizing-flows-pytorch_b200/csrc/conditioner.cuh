// conditioner.cuh -- declarations shared by the FFMA (conditioner.cu) and tensor-core (conditioner_tc.cu) conditioners.
#pragma once
#include "common.cuh"

namespace nfb {

constexpr int kF = 32;            // base_filters of every conditioner in the reference (modules.py:392,417)
constexpr int kWStage = 9 * kF * kF;  // floats of one 32->32 3x3 layer

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
    const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// packed buffer offsets (in floats); every section is a multiple of 32 floats (128 B)
struct PackLayout {
    int w0, b0;
    int bnA[2], w1[2], b1[2], w2[2], b2[2];
    int bnO, wout, bout;
    int total;
};
__host__ __device__ inline PackLayout pack_layout(int Cin, int Cout, int kk) {
    PackLayout L;
    const int CoutPad = (Cout + 31) & ~31;
    int o = 0;
    L.w0 = o; o += Cin * kk * kF;
    L.b0 = o; o += kF;
    for (int i = 0; i < 2; ++i) {
        L.bnA[i] = o; o += 2 * kF;
        L.w1[i] = o; o += kF * kk * kF;
        L.b1[i] = o; o += kF;
        L.w2[i] = o; o += kF * kk * kF;
        L.b2[i] = o; o += kF;
    }
    L.bnO = o; o += 2 * kF;
    L.wout = o; o += kF * CoutPad;
    L.bout = o; o += CoutPad;
    L.total = o;
    return L;
}


// OCT output channels x 4 pixels per thread, one 3x3 layer over CI input channels held in shared memory.
template <int OCT>
__device__ __forceinline__ void load_w(float (&wv)[OCT], const float* __restrict__ p) {
    if (OCT == 8) {
        const float4 a = ld4(p), b = ld4(p + 4);
        wv[0] = a.x; wv[1] = a.y; wv[2] = a.z; wv[3] = a.w;
        wv[4 % OCT] = b.x; wv[5 % OCT] = b.y; wv[6 % OCT] = b.z; wv[7 % OCT] = b.w;
    } else if (OCT == 4) {
        const float4 a = ld4(p);
        wv[0] = a.x; wv[1] = a.y; wv[2 % OCT] = a.z; wv[3 % OCT] = a.w;
    } else if (OCT == 2) {
        const float2 a = *reinterpret_cast<const float2*>(p);
        wv[0] = a.x; wv[1 % OCT] = a.y;
    } else {
#pragma unroll
        for (int o = 0; o < OCT; ++o) wv[o] = p[o];
    }
}

template <int H, int W, int OCT>
__device__ __forceinline__ void conv3x3_acc(float (&acc)[OCT][4], const float* __restrict__ a_base /* bufA + sample base */,
                                            int chs, const float* __restrict__ wst /* [ci][9][32] */, int CI, int og,
                                            int y, int x0) {
    const bool left_edge = (x0 == 0), right_edge = (x0 + 4 == W);
#pragma unroll 2
    for (int ci = 0; ci < CI; ++ci) {
        const float* a = a_base + ci * chs + y * W + x0;  // padded row index y+ky, ky = 0..2
        const float* wrow = wst + ci * 9 * kF + og * OCT;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const float4 c = ld4(a + ky * W);
            float l = 0.f, r = 0.f;
            if (W > 4) {
                l = __shfl_up_sync(0xffffffffu, c.w, 1);
                r = __shfl_down_sync(0xffffffffu, c.x, 1);
                if (left_edge) l = 0.f;
                if (right_edge) r = 0.f;
            }
            const float av[6] = {l, c.x, c.y, c.z, c.w, r};
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                float wv[OCT];
                load_w<OCT>(wv, wrow + (ky * 3 + kx) * kF);
#pragma unroll
                for (int o = 0; o < OCT; ++o)
#pragma unroll
                    for (int p = 0; p < 4; ++p) acc[o][p] = fmaf(wv[o], av[p + kx], acc[o][p]);
            }
        }
    }
}

// ---- tensor-core (tcgen05) sections of the packed buffer, appended after the FFMA section: 3xTF32, then FP16 split ---
// All offsets in floats relative to `base`; every block is the exact shared-memory image the kernel's TMA copies fetch.
//   consts   b0 | blk0: sA tA b1' b2 | blk1: sA tA b1' b2 | sO tO   (11 x 32 floats; b1' has the second BatchNorm of the
//            block folded in, like the weights of its conv)
//   obias_*  output-layer bias by column, generic channel order / fused (t | s_raw) order
//   in0      in-conv passes of <= 32 input channels, mid0: the four 32->32 layers.  One 3x3 stage =
//            [tap 9][k-step j][k4 2][n 64: w_hi(co) | w_lo(co)][4 ci], ci = 32*pass + 8j + 4*k4 + q: the UMMA K-major
//            no-swizzle layout of a (64 x 8) B tile per k-step (LBO 1024 B, SBO 128 B), 2048 B each
//   out*     output layer (1x1) in chunks of NW columns: [j 4][k4 2][n 2NW: hi | lo][4 ci]; fused order: columns
//            [0, NW/2) = t, [NW/2, NW) = s_raw of the same NW/2 transformed channels (coupling.py:106-107)
struct TcPlan {
    int base;
    int n_in, nj_last;       // in-conv passes; k-steps per tap of the last pass (the others have 4)
    int NWg, nqg, NWf, nqf;  // output layer: columns per chunk (multiple of 16, <= 128) and chunks, generic / fused
    int consts, obias_g, obias_f, in0, mid0, outg, outf, total;
};
constexpr int kTcStageFloats = 9 * 4 * 512;  // one full 3x3 stage of the TF32 section (72 KB)
// f16 = 1: the FP16-split section (appended after the TF32 section).  Same shared-memory images with 8 half-precision
// channels per 16-byte row instead of 4 TF32 ones: a k-step covers 16 input channels, a full stage has 2 k-steps per tap
// (36 KB), the lo parts are stored scaled by 2^kTcLoShift (conditioner_tc.cu).
__host__ __device__ inline TcPlan tc_plan(int Cin, int Cout, int f16 = 0) {
    TcPlan P;
    const int KS = f16 ? 16 : 8, NJ = 32 / KS;  // channels per k-step, k-steps per tap of a 32-channel layer
    const int stage = 9 * NJ * 512;
    P.base = (pack_layout(Cin, Cout, 9).total + 31) & ~31;
    P.n_in = (Cin + kF - 1) / kF;
    const int last = Cin - (P.n_in - 1) * kF;
    P.nj_last = (last + KS - 1) / KS;
    P.nqg = (Cout + 127) / 128;
    P.NWg = (((Cout + P.nqg - 1) / P.nqg) + 15) & ~15;
    if (Cout % 2 == 0) {
        const int c0 = Cout / 2;
        P.nqf = (((c0 + 7) & ~7) + 63) / 64;
        P.NWf = 2 * ((((c0 + P.nqf - 1) / P.nqf) + 7) & ~7);
    } else {
        P.nqf = 0;
        P.NWf = 0;
    }
    int o = 0;
    P.consts = o; o += 352;
    P.obias_g = o; o += P.nqg * P.NWg;
    P.obias_f = o; o += P.nqf * P.NWf;
    P.in0 = o; o += (P.n_in - 1) * stage + 9 * P.nj_last * 512;
    P.mid0 = o; o += 4 * stage;
    P.outg = o; o += P.nqg * 16 * NJ * P.NWg;
    P.outf = o; o += P.nqf * 16 * NJ * P.NWf;
    P.total = o;
    if (f16) P.base = tc_plan(Cin, Cout, 0).base + ((tc_plan(Cin, Cout, 0).total + 31) & ~31);
    return P;
}
// output channel behind column `col` of chunk `q` (-1: padding)
__host__ __device__ inline int tc_out_channel(bool fused, int q, int col, int NW, int c0, int Cout) {
    if (!fused) {
        const int oc = q * NW + col;
        return oc < Cout ? oc : -1;
    }
    const int PC = NW / 2;
    const int m = q * PC + (col < PC ? col : col - PC);
    if (m >= c0) return -1;
    return col < PC ? m : c0 + m;
}

int convnet_tc_dispatch(const float* zsrc, float* out, const float* pk_tc, const SplitGeom& g, int mode, int Cin, int Cout,
                        int B, int h, int w, int flags, cudaStream_t st);
int convnet_affine_tc_dispatch(float* z, float* ldj, const float* pk_tc, const SplitGeom& g, int mode, int Cin, int Cout, int B,
                               const float* sa, const float* sb, int flags, cudaStream_t st);
int convnet_affine_step_tc_dispatch(float* z, float* ldj, const float* pk_tc, const SplitGeom& g, int mode, int Cin, int Cout,
                                    int B, const float* sa, const float* sb, const float* an_ls, const float* an_b,
                                    const float* Wm, const float* log_s, int flags, cudaStream_t st);
int pack_tc_launch(const float* pk_ffma, float* pk_tc_section, int Cin, int Cout, int f16, cudaStream_t st);

}  // namespace nfb
