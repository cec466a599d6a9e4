// flowpp_cond.cu -- the Flow++ coupling conditioner (coupling.py:160-167, modules.py:519-578) as ONE kernel per call:
//   Conv2d(in,32,3) -> GatedConv2d(32,32) -> LayerNorm(32,h,w) -> GatedAttn(4 heads x 8) -> LayerNorm -> Conv2d(32,out,3)
// One CTA owns whole samples.  The 32-channel feature map lives in registers in the "conv layout" (thread = OCT channels
// x 4 pixels, as in conditioner.cu); shared memory holds the zero-haloed conv input (bufA) and two weight buffers that
// double as the token-major V / Q matrices during attention.  Attention runs in a "column layout" (thread = one token j
// and 4/TPC heads): K_j stays in registers, the softmax over i (dim=2 of modules.py:568) is an online softmax in blocks
// of 4 tokens, A = Q @ softmax is accumulated on the fly -- the N x N score matrix is never materialised (the
// reference builds a B x 4 x N x N tensor: 268 MB at B = 256, N = 256).
#include "conditioner.cuh"

namespace nfb {

struct FppArgs {
    const float* w0p;   // in conv, packed [ci][tap][32]
    const float* b0;
    const float* w1p;   // gated conv 64 -> 32, packed [ci (64)][tap][32]
    const float* b1;
    const float* ln1w;  // (32,h,w)
    const float* ln1b;
    const float* pos;   // GatedAttn.pos_emb (32,h,w)
    const float* c1w;   // conv1 (96,32), rows: V heads 0-3, K heads 0-3, Q heads 0-3 (modules.py:565-566)
    const float* c1b;
    const float* c2w;   // conv2 (64,32)
    const float* c2b;
    const float* ln2w;
    const float* ln2b;
    const float* w5p;   // out conv, packed [chunk][ci][tap][32]
    const float* b5;
};

__device__ __forceinline__ float elu_f(float v) { return v > 0.f ? v : expm1f(v); }          // F.elu, alpha = 1
__device__ __forceinline__ float sigmoid_f(float v) { return __fdiv_rn(1.f, 1.f + expf(-v)); }

// plain Conv2d weight (O, I, 3, 3) -> [o/32][i][tap][o%32], zero-padded to a multiple of 32 output channels
__global__ void __launch_bounds__(256) pack_conv3x3_kernel(const float* __restrict__ w, float* __restrict__ out, int O, int I) {
    const int Opad = (O + 31) & ~31;
    const long long total = static_cast<long long>(Opad) * I * 9;
    for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int lane = static_cast<int>(idx & 31);
        const long long r = idx >> 5;
        const int tap = static_cast<int>(r % 9);
        const long long r2 = r / 9;
        const int i = static_cast<int>(r2 % I);
        const int chunk = static_cast<int>(r2 / I);
        const int o = chunk * 32 + lane;
        out[idx] = o < O ? w[(static_cast<size_t>(o) * I + i) * 9 + tap] : 0.f;
    }
}

template <int H, int W, int NT, int OCT, int MODE>
__global__ void __launch_bounds__(NT) flowpp_cond_kernel(const float* __restrict__ zsrc, float* __restrict__ out, FppArgs A,
                                                        SplitGeom g, int Cin, int Cout, int B) {
    constexpr int N = H * W;               // tokens per sample
    constexpr int PGS = N / 4;
    constexpr int NOG = kF / OCT;
    constexpr int NPG = NT / NOG;
    constexpr int S = NPG / PGS;           // samples per CTA
    static_assert(S >= 1 && S * PGS == NPG, "tile must hold whole samples");
    constexpr int TPC = NT / (S * N);      // threads per token in the attention (column) layout
    static_assert(TPC * S * N == NT && (TPC == 1 || TPC == 2 || TPC == 4), "attention layout");
    constexpr int HPT = 4 / TPC;           // heads per thread
    constexpr int CHS = S * (H + 2) * W;
    constexpr int TS = kF + 4;             // token stride of V / Q: 36 floats -> conflict-free 16-byte stores per quarter-warp
    static_assert(S * N * TS <= kWStage, "V / Q must fit a weight buffer");
    extern __shared__ __align__(16) float smem[];
    float* bufA = smem;                    // [32][S][H+2][W]
    float* wbuf = smem + kF * CHS;         // [2][kWStage]; V / Q (token-major) and the gate output during attention
    __shared__ float red[33];

    const int t = threadIdx.x;
    // conv layout
    const int pg = t % NPG, og = t / NPG;
    const int s = pg / PGS, r = pg % PGS;
    const int y = r / (W / 4), x0 = 4 * (r % (W / 4));
    const int b = blockIdx.x * S + s;
    const bool valid = b < B;
    const int sbase = s * (H + 2) * W;
    // column (attention) layout: consecutive threads -> consecutive tokens
    const int sa = t / (N * TPC), ra = t % (N * TPC);
    const int hgrp = ra / N, j = ra % N;
    const int jy = j / W, jx = j % W;

    int pf_cnt = 0, use_cnt = 0;
    auto prefetch = [&](const float* src, int nfloats) {
        float* dst = wbuf + (pf_cnt & 1) * kWStage;
        for (int i = t * 4; i < nfloats; i += NT * 4) cp_async16(dst + i, src + i);
        cp_async_commit();
        ++pf_cnt;
    };
    auto acquire = [&]() -> const float* {
        cp_async_wait_all();
        __syncthreads();
        const float* w = wbuf + (use_cnt & 1) * kWStage;
        ++use_cnt;
        return w;
    };
    auto zero_acc = [&](float (&a)[OCT][4]) {
#pragma unroll
        for (int o = 0; o < OCT; ++o)
#pragma unroll
            for (int p = 0; p < 4; ++p) a[o][p] = 0.f;
    };
    // f(value) of this thread's tile -> interior of bufA (after everyone finished reading it)
    auto store_tile = [&](const float (&v)[OCT][4], int kind) {
        __syncthreads();
#pragma unroll
        for (int o = 0; o < OCT; ++o) {
            const int ch = og * OCT + o;
            float4 q;
            if (kind == 0) q = make_float4(elu_f(v[o][0]), elu_f(v[o][1]), elu_f(v[o][2]), elu_f(v[o][3]));
            else if (kind == 1) q = make_float4(elu_f(-v[o][0]), elu_f(-v[o][1]), elu_f(-v[o][2]), elu_f(-v[o][3]));
            else q = make_float4(v[o][0], v[o][1], v[o][2], v[o][3]);
            st4(bufA + ch * CHS + sbase + (y + 1) * W + x0, q);
        }
    };
    // LayerNorm over (32, H, W) of each sample (biased variance, eps 1e-5), elementwise affine
    auto layer_norm = [&](float (&x)[OCT][4], const float* __restrict__ lw, const float* __restrict__ lb) {
        float mean[S], rstd[S];
#pragma unroll
        for (int ss = 0; ss < S; ++ss) {
            float part = 0.f;
            if (s == ss) {
#pragma unroll
                for (int o = 0; o < OCT; ++o) part += (x[o][0] + x[o][1]) + (x[o][2] + x[o][3]);
            }
            mean[ss] = block_sum(part, red) * (1.f / (kF * N));
        }
#pragma unroll
        for (int ss = 0; ss < S; ++ss) {
            float part = 0.f;
            if (s == ss) {
#pragma unroll
                for (int o = 0; o < OCT; ++o)
#pragma unroll
                    for (int p = 0; p < 4; ++p) { const float d = x[o][p] - mean[ss]; part = fmaf(d, d, part); }
            }
            rstd[ss] = rsqrtf(block_sum(part, red) * (1.f / (kF * N)) + 1.0e-5f);
        }
        float m = mean[0], rs = rstd[0];
#pragma unroll
        for (int ss = 1; ss < S; ++ss) if (s == ss) { m = mean[ss]; rs = rstd[ss]; }
#pragma unroll
        for (int o = 0; o < OCT; ++o) {
            const int off = (og * OCT + o) * N + y * W + x0;
            const float4 wv = ldg4(lw + off), bv = ldg4(lb + off);
            x[o][0] = fmaf((x[o][0] - m) * rs, wv.x, bv.x);
            x[o][1] = fmaf((x[o][1] - m) * rs, wv.y, bv.y);
            x[o][2] = fmaf((x[o][2] - m) * rs, wv.z, bv.z);
            x[o][3] = fmaf((x[o][3] - m) * rs, wv.w, bv.w);
        }
    };

    // ---- 0: in conv, Cin -> 32 ------------------------------------------------------------------------------------
    const int n_in = (Cin + kF - 1) / kF;
    auto in_stage = [&](int c, int& n) -> const float* {
        const int ci = (Cin - c * kF) < kF ? (Cin - c * kF) : kF;
        n = ci * 9 * kF;
        return A.w0p + static_cast<size_t>(c) * kF * 9 * kF;
    };
    { int n; const float* src = in_stage(0, n); prefetch(src, n); }
    for (int i = t; i < kF * CHS; i += NT) bufA[i] = 0.f;
    __syncthreads();
    float x[OCT][4], acc[OCT][4];
    zero_acc(acc);
    for (int c = 0; c < n_in; ++c) {
        const int CI = (Cin - c * kF) < kF ? (Cin - c * kF) : kF;
        if (c > 0) __syncthreads();
        for (int i = t; i < CI * S * N; i += NT) {
            const int ci = i / (S * N);
            int rem = i - ci * (S * N);
            const int ss = rem / N;
            rem -= ss * N;
            const int bb = blockIdx.x * S + ss;
            float v = 0.f;
            if (bb < B) {
                const int jj = (c * kF + ci) * N + rem;
                if (MODE < 0) v = __ldg(zsrc + static_cast<size_t>(bb) * Cin * N + jj);
                else v = __ldg(zsrc + static_cast<size_t>(bb) * g.D + half_offset<(MODE < 0 ? NFB_SPLIT_CHANNEL : MODE)>(g, jj, 1));
            }
            bufA[ci * CHS + ss * (H + 2) * W + W + rem] = v;
        }
        const float* w = acquire();
        if (c + 1 < n_in) { int n; const float* src = in_stage(c + 1, n); prefetch(src, n); }
        else prefetch(A.w1p, kWStage);  // first half of the gated conv (input channels 0..31 = elu(x))
        conv3x3_acc<H, W, OCT>(acc, bufA + sbase, CHS, w, CI, og, y, x0);
    }
#pragma unroll
    for (int o = 0; o < OCT; ++o) {
        const float bias = __ldg(A.b0 + og * OCT + o);
#pragma unroll
        for (int p = 0; p < 4; ++p) x[o][p] = acc[o][p] + bias;
    }

    // ---- 1: GatedConv2d (modules.py:519-535): conv(elu(cat[x,-x])) -> elu(cat[y,-y]) -> y * sigmoid(a), residual -----------
    store_tile(x, 0);
    {
        const float* w = acquire();
        prefetch(A.w1p + kWStage, kWStage);  // second half: input channels 32..63 = elu(-x)
        zero_acc(acc);
        conv3x3_acc<H, W, OCT>(acc, bufA + sbase, CHS, w, kF, og, y, x0);
        store_tile(x, 1);
        w = acquire();
        conv3x3_acc<H, W, OCT>(acc, bufA + sbase, CHS, w, kF, og, y, x0);
#pragma unroll
        for (int o = 0; o < OCT; ++o) {
            const float bias = __ldg(A.b1 + og * OCT + o);
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const float v = acc[o][p] + bias;
                x[o][p] += elu_f(v) * sigmoid_f(elu_f(-v));
            }
        }
    }

    // ---- 2: LayerNorm ---------------------------------------------------------------------------------------------
    layer_norm(x, A.ln1w, A.ln1b);

    // ---- 3: GatedAttn (modules.py:556-578) --------------------------------------------------------------------------
    {
        float* Vt = wbuf;              // [S*N][32] token-major
        float* Qt = wbuf + kWStage;
        // (x + pos_emb) -> bufA, channel-major
        __syncthreads();
#pragma unroll
        for (int o = 0; o < OCT; ++o) {
            const int ch = og * OCT + o;
            const float4 pe = ldg4(A.pos + ch * N + y * W + x0);
            st4(bufA + ch * CHS + sbase + (y + 1) * W + x0,
                make_float4(x[o][0] + pe.x, x[o][1] + pe.y, x[o][2] + pe.z, x[o][3] + pe.w));
        }
        __syncthreads();
        // column layout: conv1 (1x1, 32 -> 96) for this token and this thread's heads
        float xin[kF];
        const float* col = bufA + sa * (H + 2) * W + (jy + 1) * W + jx;
#pragma unroll
        for (int c = 0; c < kF; ++c) xin[c] = col[c * CHS];
        float kreg[HPT][8];
#pragma unroll
        for (int hh = 0; hh < HPT; ++hh) {
            const int head = hgrp * HPT + hh;
#pragma unroll
            for (int part = 0; part < 3; ++part) {  // 0: V, 1: K, 2: Q
                float r8[8];
#pragma unroll
                for (int d = 0; d < 8; ++d) {
                    const int row = (part * 4 + head) * 8 + d;
                    const float* wr = A.c1w + row * kF;
                    float a = __ldg(A.c1b + row);
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4) {
                        const float4 wv = ldg4(wr + 4 * c4);
                        a = fmaf(wv.x, xin[4 * c4], a);
                        a = fmaf(wv.y, xin[4 * c4 + 1], a);
                        a = fmaf(wv.z, xin[4 * c4 + 2], a);
                        a = fmaf(wv.w, xin[4 * c4 + 3], a);
                    }
                    r8[d] = a;
                }
                if (part == 1) {
#pragma unroll
                    for (int d = 0; d < 8; ++d) kreg[hh][d] = r8[d];
                } else {
                    float* dst = (part == 0 ? Vt : Qt) + (sa * N + j) * TS + head * 8;
                    st4(dst, make_float4(r8[0], r8[1], r8[2], r8[3]));
                    st4(dst + 4, make_float4(r8[4], r8[5], r8[6], r8[7]));
                }
            }
        }
        __syncthreads();
        // A[:, j] = Q @ softmax_i(V^T K / sqrt(8))[:, j]  -- online softmax over i in blocks of 4 tokens
        float aout[HPT][8];
        const float scale = 0.35355339059327373f;  // 1/sqrt(8)
#pragma unroll
        for (int hh = 0; hh < HPT; ++hh) {
            const int head = hgrp * HPT + hh;
            float m = -INFINITY, l = 0.f, av[8];
#pragma unroll
            for (int d = 0; d < 8; ++d) av[d] = 0.f;
            const float* vb = Vt + sa * N * TS + head * 8;
            const float* qb = Qt + sa * N * TS + head * 8;
#pragma unroll 1
            for (int i0 = 0; i0 < N; i0 += 4) {
                float sc[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float4 v0 = ld4(vb + (i0 + k) * TS), v1 = ld4(vb + (i0 + k) * TS + 4);
                    float d0 = v0.x * kreg[hh][0];
                    d0 = fmaf(v0.y, kreg[hh][1], d0); d0 = fmaf(v0.z, kreg[hh][2], d0); d0 = fmaf(v0.w, kreg[hh][3], d0);
                    d0 = fmaf(v1.x, kreg[hh][4], d0); d0 = fmaf(v1.y, kreg[hh][5], d0); d0 = fmaf(v1.z, kreg[hh][6], d0);
                    d0 = fmaf(v1.w, kreg[hh][7], d0);
                    sc[k] = d0 * scale;
                }
                const float bm = fmaxf(fmaxf(sc[0], sc[1]), fmaxf(sc[2], sc[3]));
                if (bm > m) {
                    const float corr = expf(m - bm);  // exp(-inf) = 0 on the first block
                    l *= corr;
#pragma unroll
                    for (int d = 0; d < 8; ++d) av[d] *= corr;
                    m = bm;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float p = expf(sc[k] - m);
                    l += p;
                    const float4 q0 = ld4(qb + (i0 + k) * TS), q1 = ld4(qb + (i0 + k) * TS + 4);
                    av[0] = fmaf(p, q0.x, av[0]); av[1] = fmaf(p, q0.y, av[1]); av[2] = fmaf(p, q0.z, av[2]); av[3] = fmaf(p, q0.w, av[3]);
                    av[4] = fmaf(p, q1.x, av[4]); av[5] = fmaf(p, q1.y, av[5]); av[6] = fmaf(p, q1.z, av[6]); av[7] = fmaf(p, q1.w, av[7]);
                }
            }
            const float inv = __fdiv_rn(1.f, l);
#pragma unroll
            for (int d = 0; d < 8; ++d) aout[hh][d] = av[d] * inv;
        }
        // A -> bufA (channel-major; bufA is free: everyone passed the barrier after reading xin)
#pragma unroll
        for (int hh = 0; hh < HPT; ++hh)
#pragma unroll
            for (int d = 0; d < 8; ++d)
                bufA[((hgrp * HPT + hh) * 8 + d) * CHS + sa * (H + 2) * W + (jy + 1) * W + jx] = aout[hh][d];
        __syncthreads();
        // conv2 (1x1, 32 -> 64) + gate for this thread's 8*HPT channels; result -> Vt region as [c][S*N]
#pragma unroll
        for (int c = 0; c < kF; ++c) xin[c] = col[c * CHS];
        float* gate = wbuf;  // V is dead (all threads finished the softmax loop before the barrier above)
#pragma unroll
        for (int oo = 0; oo < 8 * HPT; ++oo) {
            const int o = hgrp * 8 * HPT + oo;
            float p = __ldg(A.c2b + o), a = __ldg(A.c2b + kF + o);
            const float* wp = A.c2w + o * kF;
            const float* wa = A.c2w + (kF + o) * kF;
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
                const float4 u = ldg4(wp + 4 * c4), v = ldg4(wa + 4 * c4);
                p = fmaf(u.x, xin[4 * c4], p); p = fmaf(u.y, xin[4 * c4 + 1], p); p = fmaf(u.z, xin[4 * c4 + 2], p); p = fmaf(u.w, xin[4 * c4 + 3], p);
                a = fmaf(v.x, xin[4 * c4], a); a = fmaf(v.y, xin[4 * c4 + 1], a); a = fmaf(v.z, xin[4 * c4 + 2], a); a = fmaf(v.w, xin[4 * c4 + 3], a);
            }
            gate[o * (S * N) + sa * N + j] = p * sigmoid_f(a);
        }
        __syncthreads();
#pragma unroll
        for (int o = 0; o < OCT; ++o) {
            const float4 gv = ld4(gate + (og * OCT + o) * (S * N) + s * N + y * W + x0);
            x[o][0] += gv.x; x[o][1] += gv.y; x[o][2] += gv.z; x[o][3] += gv.w;
        }
    }

    // ---- 4: LayerNorm ---------------------------------------------------------------------------------------------
    layer_norm(x, A.ln2w, A.ln2b);

    // ---- 5: out conv 32 -> Cout in chunks of 32 output channels ----------------------------------------------------------
    const int n_out = (Cout + 31) / 32;
    pf_cnt = use_cnt = 0;
    store_tile(x, 2);  // leading barrier: every thread is done with the gate buffer before the prefetch below overwrites it
    prefetch(A.w5p, kWStage);
    for (int c = 0; c < n_out; ++c) {
        const float* w = acquire();
        if (c + 1 < n_out) prefetch(A.w5p + static_cast<size_t>(c + 1) * kWStage, kWStage);
        zero_acc(acc);
        conv3x3_acc<H, W, OCT>(acc, bufA + sbase, CHS, w, kF, og, y, x0);
        if (valid) {
#pragma unroll
            for (int o = 0; o < OCT; ++o) {
                const int oc = c * kF + og * OCT + o;
                if (oc < Cout) {
                    const float bias = __ldg(A.b5 + oc);
                    st4(out + ((static_cast<size_t>(b) * Cout + oc) * H + y) * W + x0,
                        make_float4(acc[o][0] + bias, acc[o][1] + bias, acc[o][2] + bias, acc[o][3] + bias));
                }
            }
        }
    }
}

template <int H, int W, int NT, int OCT, int MODE>
static int launch_fpp(const float* zsrc, float* out, const FppArgs& A, const SplitGeom& g, int Cin, int Cout, int B,
                      cudaStream_t st) {
    constexpr int S = (NT / (kF / OCT)) / (H * W / 4);
    constexpr size_t smem = (static_cast<size_t>(kF) * S * (H + 2) * W + 2 * kWStage) * sizeof(float);
    auto kern = flowpp_cond_kernel<H, W, NT, OCT, MODE>;
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    kern<<<(B + S - 1) / S, NT, smem, st>>>(zsrc, out, A, g, Cin, Cout, B);
    return launch_status();
}

template <int MODE>
static int fpp_by_size(const float* zsrc, float* out, const FppArgs& A, const SplitGeom& g, int Cin, int Cout, int B, int h,
                       int w, cudaStream_t st) {
    if (h == 16 && w == 16) return launch_fpp<16, 16, 256, 8, MODE>(zsrc, out, A, g, Cin, Cout, B, st);
    if (h == 8 && w == 8) return launch_fpp<8, 8, 128, 4, MODE>(zsrc, out, A, g, Cin, Cout, B, st);
    if (h == 4 && w == 4) return launch_fpp<4, 4, 128, 2, MODE>(zsrc, out, A, g, Cin, Cout, B, st);
    return NFB_ERR_UNSUPPORTED;
}

}  // namespace nfb

using namespace nfb;

extern "C" int nfb_pack_conv3x3(const float* w, float* out, int O, int I, nfb_stream_t stream) {
    if (!w || !out) return NFB_ERR_NULL;
    if (O <= 0 || I <= 0) return NFB_ERR_SHAPE;
    const long long total = static_cast<long long>((O + 31) & ~31) * I * 9;
    long long blocks = (total + 255) / 256;
    if (blocks > kSMs * 8) blocks = kSMs * 8;
    pack_conv3x3_kernel<<<static_cast<int>(blocks), 256, 0, as_stream(stream)>>>(w, out, O, I);
    return launch_status();
}

extern "C" int nfb_flowpp_cond_fwd(const float* const* tensors, const float* src, float* params_out, int B, int C, int H,
                                   int W, int mode, int odd, int in_ch, int out_ch, nfb_stream_t stream) {
    if (!tensors || !src || !params_out) return NFB_ERR_NULL;
    for (int i = 0; i < 15; ++i)
        if (!tensors[i]) return NFB_ERR_NULL;
    if (in_ch <= 0 || out_ch <= 0) return NFB_ERR_SHAPE;
    const FppArgs A{tensors[0], tensors[1], tensors[2],  tensors[3],  tensors[4],  tensors[5],  tensors[6], tensors[7],
                    tensors[8], tensors[9], tensors[10], tensors[11], tensors[12], tensors[13], tensors[14]};
    cudaStream_t st = as_stream(stream);
    SplitGeom g{};
    if (mode < 0) {
        if (B <= 0 || H <= 0 || W <= 0) return NFB_ERR_SHAPE;
        return fpp_by_size<-1>(src, params_out, A, g, in_ch, out_ch, B, H, W, st);
    }
    const int rc = make_geom(g, B, C, H, W, mode, odd);
    if (rc != NFB_OK) return rc;
    if (g.c0 != in_ch) return NFB_ERR_SHAPE;
    if (mode == NFB_SPLIT_CHECKER) return fpp_by_size<NFB_SPLIT_CHECKER>(src, params_out, A, g, in_ch, out_ch, B, g.h, g.w, st);
    if (mode == NFB_SPLIT_CHANNEL) return fpp_by_size<NFB_SPLIT_CHANNEL>(src, params_out, A, g, in_ch, out_ch, B, g.h, g.w, st);
    return NFB_ERR_SHAPE;
}
