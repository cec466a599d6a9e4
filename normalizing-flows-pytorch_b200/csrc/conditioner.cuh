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

// ---- tensor-core (tcgen05, 3xTF32) section of the packed buffer, appended after the FFMA section -----------------
// consts: b0 | blk0: sA tA b1' b2 | blk1: sA tA b1' b2 | sO tO  (11 x 32 floats) + bout (CoutPad); b1' has the second
// BatchNorm of the block folded in (like the weights of its conv).  stage i (n_in in-conv passes, then 4 mid layers): hi[9216],
// lo[9216] floats, B operand K-major no-swizzle: element (n, tap, ci) at ((tap*8 + ci/4)*32 + n)*4 + ci%4.
// out: per chunk of 32 output channels 1024 floats, (n, ci) at ((ci/4)*32 + n)*4 + ci%4; hi section then lo section.
constexpr int kTcStage = 9 * 8 * 32 * 4;  // 9216 floats per hi (or lo) half of a 3x3 stage
struct TcLayout {
    int base;     // offset of the TC section inside the packed buffer
    int consts, stage0, out_hi, out_lo, total;  // offsets relative to `base`
    int n_in, n_chunks, cout_pad;
};
__host__ __device__ inline TcLayout tc_layout(int Cin, int Cout) {
    TcLayout T;
    T.base = (pack_layout(Cin, Cout, 9).total + 31) & ~31;
    T.n_in = (Cin + kF - 1) / kF;
    T.cout_pad = (Cout + 31) & ~31;
    T.n_chunks = T.cout_pad / 32;
    int o = 0;
    T.consts = o; o += 352 + T.cout_pad;
    T.stage0 = o; o += (T.n_in + 4) * 2 * kTcStage;
    T.out_hi = o; o += T.n_chunks * 1024;
    T.out_lo = o; o += T.n_chunks * 1024;
    T.total = o;
    return T;
}

extern int g_tune[8];
int convnet_tc_dispatch(const float* zsrc, float* out, const float* pk_tc, const SplitGeom& g, int mode, int Cin, int Cout,
                        int B, int h, int w, cudaStream_t st);
int pack_tc_launch(const float* pk_ffma, float* pk_tc_section, int Cin, int Cout, cudaStream_t st);

}  // namespace nfb
