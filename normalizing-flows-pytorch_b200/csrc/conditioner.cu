// conditioner.cu -- the couplings' conditioner networks (modules.py:342-438, weight_norm.py) as ONE kernel per call.
//
// ConvNet: conv3x3(in->32) -> 2 x [BN, ReLU, conv3x3, BN, ReLU, conv3x3] + skip -> BN, ReLU, conv1x1(32->out);
// MLP: the same with Linear layers.  Eval mode: WeightNorm and the BatchNorm that directly follows a conv are folded
// into packed weights once per weight update (nfb_pack_*), the remaining BatchNorms are per-channel scale/shift.
//
// Design (B200): the 32-channel activations of a whole sample never leave the SM -- the residual stream lives in
// registers (each thread owns 8 output channels x 4 consecutive pixels for every layer), the ReLU'd conv input lives
// in shared memory with zero rows above/below (x borders by predication, x neighbours by warp shuffle), and the
// 36.9 KB of weights per 3x3 layer are streamed L2 -> shared memory with cp.async, double-buffered so the next
// layer's weights land while the current layer computes.  The conditioner input z1 is gathered straight from z with
// the coupling's split addressing (no materialised split).  Arithmetic is FP32 FFMA: bits/dim parity at 1e-5 with a
// reference noise floor of 2e-7 rules out single-pass TF32/BF16 tensor-core math; an error-compensated
// tcgen05 path is the next step.
#include <cooperative_groups.h>

#include "conditioner.cuh"

namespace nfb {

// ---------------------------------------------------------------------------------------------------------
// packing (once per weight update)
// ---------------------------------------------------------------------------------------------------------
// One weight-normalised layer: v (O, I, kk), g (I, kk), bias (O)  ->  w_out laid out [(i*kk+tap)][chunk][32] with
// o = chunk*32 + lane, i.e. index ((o/32) * I*kk + i*kk + tap) * 32 + o%32  (rows of 32 output channels; for O <= 32
// a single chunk).  Optionally folds a BatchNorm (eval) that directly follows the layer:
//   y = ((conv + b) - rm) / sqrt(rv + eps) * gamma + beta.
__global__ void __launch_bounds__(128) pack_wn_kernel(const float* __restrict__ v, const float* __restrict__ gw,
                                                     const float* __restrict__ bias, const float* __restrict__ bn_w,
                                                     const float* __restrict__ bn_b, const float* __restrict__ bn_rm,
                                                     const float* __restrict__ bn_rv, float* __restrict__ w_out,
                                                     float* __restrict__ b_out, int O, int J /* I*kk */, int O_pad,
                                                     float wn_eps, float bn_eps) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < J) {
        float ss = 0.f;
        for (int o = 0; o < O; ++o) { const float x = v[static_cast<size_t>(o) * J + j]; ss = fmaf(x, x, ss); }
        const float scale = __fdiv_rn(gw[j], __fadd_rn(sqrtf(ss), wn_eps));  // weight_norm.py:40
        for (int o = 0; o < O_pad; ++o) {
            float w = 0.f;
            if (o < O) {
                w = __fmul_rn(v[static_cast<size_t>(o) * J + j], scale);
                if (bn_w) w *= bn_w[o] / sqrtf(bn_rv[o] + bn_eps);
            }
            w_out[(static_cast<size_t>(o >> 5) * J + j) * 32 + (o & 31)] = w;
        }
    }
    if (blockIdx.x == 0) {
        for (int o = threadIdx.x; o < O_pad; o += blockDim.x) {
            float b = 0.f;
            if (o < O) {
                b = bias[o];
                if (bn_w) {
                    const float sc = bn_w[o] / sqrtf(bn_rv[o] + bn_eps);
                    b = (b - bn_rm[o]) * sc + bn_b[o];
                }
            }
            b_out[o] = b;
        }
    }
}

// BatchNorm (eval) as scale/shift: y = x*scale + shift
__global__ void pack_bn_kernel(const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ rm,
                               const float* __restrict__ rv, float* __restrict__ scale, float* __restrict__ shift, int C,
                               float eps) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) {
        const float sc = w[c] / sqrtf(rv[c] + eps);
        scale[c] = sc;
        shift[c] = b[c] - rm[c] * sc;
    }
}

// ---------------------------------------------------------------------------------------------------------
// fused ConvNet
// ---------------------------------------------------------------------------------------------------------
// MODE: NFB_SPLIT_CHECKER / NFB_SPLIT_CHANNEL gather z1 from the coupling's z; MODE < 0: x is the (B,Cin,H,W) input.
template <int H, int W, int NT, int OCT, int MODE>
__global__ void __launch_bounds__(NT) convnet_fused_kernel(const float* __restrict__ zsrc, float* __restrict__ out,
                                                          const float* __restrict__ pk, SplitGeom g, int Cin, int Cout,
                                                          int B) {
    constexpr int PGS = H * W / 4;         // pixel groups per sample
    constexpr int NOG = kF / OCT;          // groups of output channels
    constexpr int NPG = NT / NOG;          // pixel groups per CTA
    constexpr int S = NPG / PGS;           // samples per CTA
    static_assert(S >= 1 && S * PGS == NPG, "tile must hold whole samples");
    constexpr int CHS = S * (H + 2) * W;   // channel stride of bufA (rows 0 and H+1 of every sample stay zero)
    extern __shared__ __align__(16) float smem[];
    float* bufA = smem;                    // [32][S][H+2][W]
    float* wbuf = smem + kF * CHS;         // [2][kWStage]

    const PackLayout L = pack_layout(Cin, Cout, 9);
    const int CoutPad = (Cout + 31) & ~31;
    const int t = threadIdx.x;
    const int pg = t % NPG, og = t / NPG;
    const int s = pg / PGS, r = pg % PGS;
    const int y = r / (W / 4), x0 = 4 * (r % (W / 4));
    const int b = blockIdx.x * S + s;
    const bool valid = b < B;
    const int sbase = s * (H + 2) * W;

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

    // weight stage schedule: in-conv chunks, 4 mid layers, out chunks
    const int n_in = (Cin + kF - 1) / kF;
    const int n_out = CoutPad / kF;
    const int n_stage = n_in + 4 + n_out;
    auto stage_src = [&](int i, int& n) -> const float* {
        if (i < n_in) {
            const int ci = (Cin - i * kF) < kF ? (Cin - i * kF) : kF;
            n = ci * 9 * kF;
            return pk + L.w0 + i * kF * 9 * kF;
        }
        i -= n_in;
        if (i < 4) { n = kWStage; return pk + ((i & 1) ? L.w2[i >> 1] : L.w1[i >> 1]); }
        i -= 4;
        n = kF * kF;
        return pk + L.wout + i * kF * kF;
    };
    int stage = 0;
    auto prefetch_stage = [&](int i) {
        if (i < n_stage) { int n; const float* src = stage_src(i, n); prefetch(src, n); }
    };

    prefetch_stage(0);
    for (int i = t; i < kF * CHS; i += NT) bufA[i] = 0.f;  // includes the zero halo rows
    __syncthreads();
    const float* src = zsrc;

    float xres[OCT][4];  // residual stream: this thread's OCT channels x 4 pixels
    float acc[OCT][4];
#pragma unroll
    for (int o = 0; o < OCT; ++o)
#pragma unroll
        for (int p = 0; p < 4; ++p) acc[o][p] = 0.f;

    // ---- in conv: Cin -> 32, input channels in chunks of 32 ---------------------------------------------
    for (int c = 0; c < n_in; ++c) {
        const int CI = (Cin - c * kF) < kF ? (Cin - c * kF) : kF;
        if (c > 0) __syncthreads();  // everyone finished reading the previous chunk
        for (int i = t; i < CI * S * H * W; i += NT) {
            const int ci = i / (S * H * W);
            int rem = i - ci * (S * H * W);
            const int ss = rem / (H * W);
            rem -= ss * (H * W);
            const int bb = blockIdx.x * S + ss;
            float v = 0.f;
            if (bb < B) {
                const int j = (c * kF + ci) * (H * W) + rem;  // index inside the (Cin, H, W) conditioner input
                if (MODE < 0) v = __ldg(src + static_cast<size_t>(bb) * Cin * (H * W) + j);
                else v = __ldg(src + static_cast<size_t>(bb) * g.D + half_offset<(MODE < 0 ? NFB_SPLIT_CHANNEL : MODE)>(g, j, 1));
            }
            bufA[ci * CHS + ss * (H + 2) * W + W + rem] = v;  // +W: skip the zero row above
        }
        const float* w = acquire();
        prefetch_stage(++stage);
        conv3x3_acc<H, W, OCT>(acc, bufA + sbase, CHS, w, CI, og, y, x0);
    }
#pragma unroll
    for (int o = 0; o < OCT; ++o) {
        const float bias = __ldg(pk + L.b0 + og * OCT + o);
#pragma unroll
        for (int p = 0; p < 4; ++p) xres[o][p] = acc[o][p] + bias;
    }

    // write relu(scale*v + shift) of this thread's tile into bufA (after everyone finished reading it)
    auto store_act = [&](const float (&v)[OCT][4], const float* sc_sh) {
        __syncthreads();
#pragma unroll
        for (int o = 0; o < OCT; ++o) {
            const int ch = og * OCT + o;
            const float sc = sc_sh ? __ldg(sc_sh + ch) : 1.f, sh = sc_sh ? __ldg(sc_sh + kF + ch) : 0.f;
            float4 q;
            q.x = fmaxf(fmaf(v[o][0], sc, sh), 0.f);
            q.y = fmaxf(fmaf(v[o][1], sc, sh), 0.f);
            q.z = fmaxf(fmaf(v[o][2], sc, sh), 0.f);
            q.w = fmaxf(fmaf(v[o][3], sc, sh), 0.f);
            st4(bufA + ch * CHS + sbase + (y + 1) * W + x0, q);
        }
    };

    // ---- two residual blocks ------------------------------------------------------------------------------
#pragma unroll 1
    for (int blk = 0; blk < 2; ++blk) {
        store_act(xres, pk + L.bnA[blk]);
        const float* w = acquire();
        prefetch_stage(++stage);
#pragma unroll
        for (int o = 0; o < OCT; ++o)
#pragma unroll
            for (int p = 0; p < 4; ++p) acc[o][p] = 0.f;
        conv3x3_acc<H, W, OCT>(acc, bufA + sbase, CHS, w, kF, og, y, x0);
#pragma unroll
        for (int o = 0; o < OCT; ++o) {
            const float bias = __ldg(pk + L.b1[blk] + og * OCT + o);  // second BN folded into (w1, b1)
#pragma unroll
            for (int p = 0; p < 4; ++p) acc[o][p] += bias;
        }
        store_act(acc, nullptr);  // relu only
        w = acquire();
        prefetch_stage(++stage);
#pragma unroll
        for (int o = 0; o < OCT; ++o)
#pragma unroll
            for (int p = 0; p < 4; ++p) acc[o][p] = 0.f;
        conv3x3_acc<H, W, OCT>(acc, bufA + sbase, CHS, w, kF, og, y, x0);
#pragma unroll
        for (int o = 0; o < OCT; ++o) {
            const float bias = __ldg(pk + L.b2[blk] + og * OCT + o);
#pragma unroll
            for (int p = 0; p < 4; ++p) xres[o][p] += acc[o][p] + bias;
        }
    }

    // ---- out block: BN, ReLU, conv1x1 32 -> Cout in chunks of 32 output channels ---------------------------------
    store_act(xres, pk + L.bnO);
    for (int c = 0; c < n_out; ++c) {
        const float* w = acquire();  // [32 ic][32 oc]
        prefetch_stage(++stage);
#pragma unroll
        for (int o = 0; o < OCT; ++o)
#pragma unroll
            for (int p = 0; p < 4; ++p) acc[o][p] = 0.f;
        const float* a = bufA + sbase + (y + 1) * W + x0;
#pragma unroll 4
        for (int ci = 0; ci < kF; ++ci) {
            const float4 v = ld4(a + ci * CHS);
            float wv[OCT];
            load_w<OCT>(wv, w + ci * kF + og * OCT);
#pragma unroll
            for (int o = 0; o < OCT; ++o) {
                acc[o][0] = fmaf(wv[o], v.x, acc[o][0]);
                acc[o][1] = fmaf(wv[o], v.y, acc[o][1]);
                acc[o][2] = fmaf(wv[o], v.z, acc[o][2]);
                acc[o][3] = fmaf(wv[o], v.w, acc[o][3]);
            }
        }
        if (valid) {
#pragma unroll
            for (int o = 0; o < OCT; ++o) {
                const int oc = c * kF + og * OCT + o;
                if (oc < Cout) {
                    const float bias = __ldg(pk + L.bout + oc);
                    const float4 r = make_float4(acc[o][0] + bias, acc[o][1] + bias, acc[o][2] + bias, acc[o][3] + bias);
                    st4(out + ((static_cast<size_t>(b) * Cout + oc) * H + y) * W + x0, r);
                }
            }
        }
    }
}

template <int H, int W, int NT, int OCT, int MODE>
static int launch_convnet(const float* zsrc, float* out, const float* pk, const SplitGeom& g, int Cin, int Cout, int B,
                          cudaStream_t st) {
    constexpr int S = (NT / (kF / OCT)) / (H * W / 4);
    constexpr size_t smem = (static_cast<size_t>(kF) * S * (H + 2) * W + 2 * kWStage) * sizeof(float);
    auto kern = convnet_fused_kernel<H, W, NT, OCT, MODE>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    kern<<<(B + S - 1) / S, NT, smem, st>>>(zsrc, out, pk, g, Cin, Cout, B);
    return launch_status();
}

// ---------------------------------------------------------------------------------------------------------
// 32x32 conditioner inputs (first level of a 64x64 Glow): one sample = a CLUSTER of two CTAs.
// Each CTA owns 16 image rows (512 pixels, 512 threads, same 8 channels x 4 pixels thread tile); its padded buffer has
// one halo row above and below.  After every layer the boundary row is written straight into the peer CTA's halo row
// through distributed shared memory (cluster.map_shared_rank) and a cluster barrier replaces the block barrier.
// ---------------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(512, 1)
    convnet_cluster32_kernel(const float* __restrict__ zsrc, float* __restrict__ out, const float* __restrict__ pk, SplitGeom g,
                             int Cin, int Cout, int B) {
    namespace cg = cooperative_groups;
    constexpr int HL = 16, W = 32, NT = 512, OCT = 8, HF = 32;  // local rows, width, threads, tile, full height
    constexpr int NPG = NT / (kF / OCT);                         // 128 pixel groups = 16 rows x 8
    constexpr int CHS = (HL + 2) * W;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = static_cast<int>(cluster.block_rank());     // 0: rows 0..15, 1: rows 16..31
    extern __shared__ __align__(16) float smem[];
    float* bufA = smem;                                          // [32][HL+2][W]
    float* wbuf = smem + kF * CHS;                               // [2][kWStage]
    float* peerA = cluster.map_shared_rank(bufA, rank ^ 1);

    const PackLayout L = pack_layout(Cin, Cout, 9);
    const int CoutPad = (Cout + 31) & ~31;
    const int t = threadIdx.x;
    const int pg = t % NPG, og = t / NPG;
    const int y = pg / (W / 4), x0 = 4 * (pg % (W / 4));
    const int b = blockIdx.x >> 1;
    const int gy = rank * HL + y;                                // row in the full image

    int pf_cnt = 0, use_cnt = 0;
    auto prefetch = [&](const float* src, int nfloats) {
        float* dst = wbuf + (pf_cnt & 1) * kWStage;
        for (int i = t * 4; i < nfloats; i += NT * 4) cp_async16(dst + i, src + i);
        cp_async_commit();
        ++pf_cnt;
    };
    const int n_in = (Cin + kF - 1) / kF;
    const int n_out = CoutPad / kF;
    const int n_stage = n_in + 4 + n_out;
    auto stage_src = [&](int i, int& n) -> const float* {
        if (i < n_in) {
            const int ci = (Cin - i * kF) < kF ? (Cin - i * kF) : kF;
            n = ci * 9 * kF;
            return pk + L.w0 + i * kF * 9 * kF;
        }
        i -= n_in;
        if (i < 4) { n = kWStage; return pk + ((i & 1) ? L.w2[i >> 1] : L.w1[i >> 1]); }
        i -= 4;
        n = kF * kF;
        return pk + L.wout + i * kF * kF;
    };
    int stage = 0;
    auto prefetch_stage = [&](int i) {
        if (i < n_stage) { int n; const float* src = stage_src(i, n); prefetch(src, n); }
    };
    // weights of the current stage landed + every activation write (own and the peer's halo row) is visible
    auto acquire = [&]() -> const float* {
        cp_async_wait_all();
        cluster.sync();
        const float* w = wbuf + (use_cnt & 1) * kWStage;
        ++use_cnt;
        return w;
    };

    prefetch_stage(0);
    for (int i = t; i < kF * CHS; i += NT) bufA[i] = 0.f;
    __syncthreads();

    float xres[OCT][4], acc[OCT][4];
#pragma unroll
    for (int o = 0; o < OCT; ++o)
#pragma unroll
        for (int p = 0; p < 4; ++p) acc[o][p] = 0.f;

    // ---- in conv: both the interior rows and the two halo rows come straight from global memory -------------------
    for (int c = 0; c < n_in; ++c) {
        const int CI = (Cin - c * kF) < kF ? (Cin - c * kF) : kF;
        if (c > 0) __syncthreads();
        for (int i = t; i < CI * (HL + 2) * W; i += NT) {
            const int ci = i / ((HL + 2) * W);
            const int rem = i - ci * ((HL + 2) * W);
            const int pr = rem / W, xx = rem - pr * W;
            const int yy = rank * HL + pr - 1;
            float v = 0.f;
            if (yy >= 0 && yy < HF) {
                const int j = (c * kF + ci) * (HF * W) + yy * W + xx;
                if (MODE < 0) v = __ldg(zsrc + static_cast<size_t>(b) * Cin * (HF * W) + j);
                else v = __ldg(zsrc + static_cast<size_t>(b) * g.D + half_offset<(MODE < 0 ? NFB_SPLIT_CHANNEL : MODE)>(g, j, 1));
            }
            bufA[ci * CHS + rem] = v;
        }
        const float* w = acquire();
        prefetch_stage(++stage);
        conv3x3_acc<HL, W, OCT>(acc, bufA, CHS, w, CI, og, y, x0);
    }
#pragma unroll
    for (int o = 0; o < OCT; ++o) {
        const float bias = __ldg(pk + L.b0 + og * OCT + o);
#pragma unroll
        for (int p = 0; p < 4; ++p) xres[o][p] = acc[o][p] + bias;
    }

    // relu(scale*v + shift) -> own interior row, and into the peer's halo row when this is a boundary row
    auto store_act = [&](const float (&v)[OCT][4], const float* sc_sh) {
        cluster.sync();  // both CTAs finished reading the previous activations (including each other's halo rows)
        const bool to_peer = (rank == 0 && y == HL - 1) || (rank == 1 && y == 0);
        const int peer_row = rank == 0 ? 0 : HL + 1;  // my last row is the peer's top halo; my first row its bottom halo
#pragma unroll
        for (int o = 0; o < OCT; ++o) {
            const int ch = og * OCT + o;
            const float sc = sc_sh ? __ldg(sc_sh + ch) : 1.f, sh = sc_sh ? __ldg(sc_sh + kF + ch) : 0.f;
            float4 q;
            q.x = fmaxf(fmaf(v[o][0], sc, sh), 0.f);
            q.y = fmaxf(fmaf(v[o][1], sc, sh), 0.f);
            q.z = fmaxf(fmaf(v[o][2], sc, sh), 0.f);
            q.w = fmaxf(fmaf(v[o][3], sc, sh), 0.f);
            st4(bufA + ch * CHS + (y + 1) * W + x0, q);
            if (to_peer) st4(peerA + ch * CHS + peer_row * W + x0, q);
        }
    };

#pragma unroll 1
    for (int blk = 0; blk < 2; ++blk) {
        store_act(xres, pk + L.bnA[blk]);
        const float* w = acquire();
        prefetch_stage(++stage);
#pragma unroll
        for (int o = 0; o < OCT; ++o)
#pragma unroll
            for (int p = 0; p < 4; ++p) acc[o][p] = 0.f;
        conv3x3_acc<HL, W, OCT>(acc, bufA, CHS, w, kF, og, y, x0);
#pragma unroll
        for (int o = 0; o < OCT; ++o) {
            const float bias = __ldg(pk + L.b1[blk] + og * OCT + o);
#pragma unroll
            for (int p = 0; p < 4; ++p) acc[o][p] += bias;
        }
        store_act(acc, nullptr);
        w = acquire();
        prefetch_stage(++stage);
#pragma unroll
        for (int o = 0; o < OCT; ++o)
#pragma unroll
            for (int p = 0; p < 4; ++p) acc[o][p] = 0.f;
        conv3x3_acc<HL, W, OCT>(acc, bufA, CHS, w, kF, og, y, x0);
#pragma unroll
        for (int o = 0; o < OCT; ++o) {
            const float bias = __ldg(pk + L.b2[blk] + og * OCT + o);
#pragma unroll
            for (int p = 0; p < 4; ++p) xres[o][p] += acc[o][p] + bias;
        }
    }

    store_act(xres, pk + L.bnO);  // the 1x1 output conv needs no halo, but the barrier pairing stays uniform
    for (int c = 0; c < n_out; ++c) {
        const float* w = acquire();
        prefetch_stage(++stage);
#pragma unroll
        for (int o = 0; o < OCT; ++o)
#pragma unroll
            for (int p = 0; p < 4; ++p) acc[o][p] = 0.f;
        const float* a = bufA + (y + 1) * W + x0;
#pragma unroll 4
        for (int ci = 0; ci < kF; ++ci) {
            const float4 v = ld4(a + ci * CHS);
            float wv[OCT];
            load_w<OCT>(wv, w + ci * kF + og * OCT);
#pragma unroll
            for (int o = 0; o < OCT; ++o) {
                acc[o][0] = fmaf(wv[o], v.x, acc[o][0]);
                acc[o][1] = fmaf(wv[o], v.y, acc[o][1]);
                acc[o][2] = fmaf(wv[o], v.z, acc[o][2]);
                acc[o][3] = fmaf(wv[o], v.w, acc[o][3]);
            }
        }
#pragma unroll
        for (int o = 0; o < OCT; ++o) {
            const int oc = c * kF + og * OCT + o;
            if (oc < Cout) {
                const float bias = __ldg(pk + L.bout + oc);
                st4(out + ((static_cast<size_t>(b) * Cout + oc) * HF + gy) * W + x0,
                    make_float4(acc[o][0] + bias, acc[o][1] + bias, acc[o][2] + bias, acc[o][3] + bias));
            }
        }
    }
    cluster.sync();  // no CTA may exit while its peer can still write into its shared memory
}

template <int MODE>
static int launch_cluster32(const float* zsrc, float* out, const float* pk, const SplitGeom& g, int Cin, int Cout, int B,
                            cudaStream_t st) {
    constexpr size_t smem = (static_cast<size_t>(kF) * 18 * 32 + 2 * kWStage) * sizeof(float);
    auto kern = convnet_cluster32_kernel<MODE>;
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    kern<<<2 * B, 512, smem, st>>>(zsrc, out, pk, g, Cin, Cout, B);
    return launch_status();
}

template <int MODE>
static int dispatch_convnet(const float* zsrc, float* out, const float* pk, const SplitGeom& g, int Cin, int Cout, int B,
                            int h, int w, int flags, cudaStream_t st) {
    if (!(flags & NFB_CONV_FFMA)) {  // default: tensor-core (tcgen05, split-precision) kernel
        const int rc = convnet_tc_dispatch(zsrc, out, pk, g, MODE, Cin, Cout, B, h, w, flags, st);
        if (rc != NFB_ERR_UNSUPPORTED) return rc;
    }
    const int variant = flags & NFB_CONV_VARIANT_MASK;
#define NFB_CONV(H_, W_, NT_, OCT_) return launch_convnet<H_, W_, NT_, OCT_, MODE>(zsrc, out, pk, g, Cin, Cout, B, st)
    if (h == 16 && w == 16) {
        switch (variant) {
            case 1: NFB_CONV(16, 16, 512, 4);
            default: NFB_CONV(16, 16, 256, 8);
        }
    }
    if (h == 8 && w == 8) {
        switch (variant) {
            case 1: NFB_CONV(8, 8, 64, 8);
            case 2: NFB_CONV(8, 8, 256, 2);
            case 3: NFB_CONV(8, 8, 256, 4);
            default: NFB_CONV(8, 8, 128, 4);
        }
    }
    if (h == 4 && w == 4) {
        switch (variant) {
            case 1: NFB_CONV(4, 4, 32, 8);
            case 2: NFB_CONV(4, 4, 64, 2);
            case 3: NFB_CONV(4, 4, 128, 4);
            case 4: NFB_CONV(4, 4, 256, 2);
            default: NFB_CONV(4, 4, 128, 2);
        }
    }
#undef NFB_CONV
    if (h == 32 && w == 32) return launch_cluster32<MODE>(zsrc, out, pk, g, Cin, Cout, B, st);
    return NFB_ERR_UNSUPPORTED;
}

// ---------------------------------------------------------------------------------------------------------
// fused MLP (1-D couplings): one thread per sample, the 32 hidden activations in registers, weights in shared memory
// ---------------------------------------------------------------------------------------------------------
template <int MODE>  // NFB_SPLIT_1D gathers z1 from z (row stride g.D); MODE < 0: x is (B, Cin)
__global__ void __launch_bounds__(256) mlp_fused_kernel(const float* __restrict__ zsrc, float* __restrict__ out,
                                                       const float* __restrict__ pk, SplitGeom g, int Cin, int Cout,
                                                       int B) {
    extern __shared__ __align__(16) float sw[];  // the whole packed network, then one 32x33 transpose tile per warp
    const PackLayout L = pack_layout(Cin, Cout, 1);
    for (int i = threadIdx.x * 4; i < L.total; i += blockDim.x * 4) st4(sw + i, ldg4(pk + i));
    __syncthreads();
    const int CoutPad = (Cout + 31) & ~31;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* tile = sw + ((L.total + 3) & ~3) + warp * (32 * 33);
    // warp-uniform loop (every lane of a warp iterates the same number of times: the tile transpose is collective)
    for (int b0 = (blockIdx.x * (blockDim.x >> 5) + warp) * 32; b0 < B; b0 += gridDim.x * blockDim.x) {
        const int b = b0 + lane;
        const bool live = b < B;
        float x[kF], a[kF], y[kF];
#pragma unroll
        for (int o = 0; o < kF; ++o) x[o] = sw[L.b0 + o];
        for (int ci = 0; ci < Cin; ++ci) {
            float v = 0.f;
            if (live) {
                if (MODE < 0) v = __ldg(zsrc + static_cast<size_t>(b) * Cin + ci);
                else v = __ldg(zsrc + static_cast<size_t>(b) * g.D + 2 * ci + (g.odd ? 0 : 1));  // z1 of squeeze1d
            }
            const float* wr = sw + L.w0 + ci * kF;
#pragma unroll
            for (int o = 0; o < kF; ++o) x[o] = fmaf(wr[o], v, x[o]);
        }
#pragma unroll 1
        for (int blk = 0; blk < 2; ++blk) {
#pragma unroll
            for (int o = 0; o < kF; ++o)
                a[o] = fmaxf(fmaf(x[o], sw[L.bnA[blk] + o], sw[L.bnA[blk] + kF + o]), 0.f);
#pragma unroll
            for (int o = 0; o < kF; ++o) y[o] = sw[L.b1[blk] + o];
#pragma unroll
            for (int ci = 0; ci < kF; ++ci) {
                const float* wr = sw + L.w1[blk] + ci * kF;
#pragma unroll
                for (int o = 0; o < kF; ++o) y[o] = fmaf(wr[o], a[ci], y[o]);
            }
#pragma unroll
            for (int o = 0; o < kF; ++o) { a[o] = fmaxf(y[o], 0.f); y[o] = sw[L.b2[blk] + o]; }
#pragma unroll
            for (int ci = 0; ci < kF; ++ci) {
                const float* wr = sw + L.w2[blk] + ci * kF;
#pragma unroll
                for (int o = 0; o < kF; ++o) y[o] = fmaf(wr[o], a[ci], y[o]);
            }
#pragma unroll
            for (int o = 0; o < kF; ++o) x[o] += y[o];
        }
#pragma unroll
        for (int o = 0; o < kF; ++o) a[o] = fmaxf(fmaf(x[o], sw[L.bnO + o], sw[L.bnO + kF + o]), 0.f);
        for (int c = 0; c < CoutPad / kF; ++c) {
#pragma unroll
            for (int o = 0; o < kF; ++o) y[o] = sw[L.bout + c * kF + o];
#pragma unroll
            for (int ci = 0; ci < kF; ++ci) {
                const float* wr = sw + L.wout + (c * kF + ci) * kF;
#pragma unroll
                for (int o = 0; o < kF; ++o) y[o] = fmaf(wr[o], a[ci], y[o]);
            }
            // lane owns a sample; transpose the 32 samples x 32 outputs block through shared memory so that every
            // warp store is one contiguous 128-byte row of the (B, Cout) output
            __syncwarp();
#pragma unroll
            for (int o = 0; o < kF; ++o) tile[lane * 33 + o] = y[o];
            __syncwarp();
            const int oc = c * kF + lane;
            if (oc < Cout) {
                const int rows = (B - b0) < 32 ? (B - b0) : 32;
                for (int r = 0; r < rows; ++r) out[static_cast<size_t>(b0 + r) * Cout + oc] = tile[r * 33 + lane];
            }
        }
    }
}

}  // namespace nfb

using namespace nfb;

extern "C" int nfb_resnet_pack_size(int in_ch, int out_ch, int conv) {
    if (in_ch <= 0 || out_ch <= 0) return NFB_ERR_SHAPE;
    if (!conv) return pack_layout(in_ch, out_ch, 1).total;
    const TcPlan T = tc_plan(in_ch, out_ch, 1);  // FFMA section | tensor-core 3xTF32 section | FP16-split section
    return T.base + T.total;
}

extern "C" int nfb_resnet_pack(const float* const* t, float* packed, int in_ch, int out_ch, int conv, float wn_eps,
                               float bn_eps, nfb_stream_t stream) {
    // t[0..17]: (v, g, bias) of in_block.0, mid0.net.2, mid0.net.5, mid1.net.2, mid1.net.5, out_block.2
    // t[18..37]: (weight, bias, running_mean, running_var) of mid0.net.0, mid0.net.3, mid1.net.0, mid1.net.3, out_block.0
    if (!t || !packed) return NFB_ERR_NULL;
    for (int i = 0; i < 38; ++i)
        if (!t[i]) return NFB_ERR_NULL;
    if (in_ch <= 0 || out_ch <= 0) return NFB_ERR_SHAPE;
    const int kk = conv ? 9 : 1;
    const PackLayout L = pack_layout(in_ch, out_ch, kk);
    const int CoutPad = (out_ch + 31) & ~31;
    cudaStream_t st = as_stream(stream);
    auto wn = [&](int li, int bn /* -1: none */, float* w, float* b, int O, int J, int Opad) {
        const float* const* q = t + 3 * li;
        const float* const* n = bn >= 0 ? t + 18 + 4 * bn : nullptr;
        pack_wn_kernel<<<(J + 127) / 128, 128, 0, st>>>(q[0], q[1], q[2], n ? n[0] : nullptr, n ? n[1] : nullptr,
                                                        n ? n[2] : nullptr, n ? n[3] : nullptr, w, b, O, J, Opad, wn_eps,
                                                        bn_eps);
        return launch_status();
    };
    auto bn = [&](int bi, float* dst) {
        const float* const* n = t + 18 + 4 * bi;
        pack_bn_kernel<<<1, kF, 0, st>>>(n[0], n[1], n[2], n[3], dst, dst + kF, kF, bn_eps);
        return launch_status();
    };
    int rc;
    if ((rc = wn(0, -1, packed + L.w0, packed + L.b0, kF, in_ch * kk, kF))) return rc;
    for (int i = 0; i < 2; ++i) {
        if ((rc = bn(2 * i, packed + L.bnA[i]))) return rc;
        if ((rc = wn(1 + 2 * i, 2 * i + 1, packed + L.w1[i], packed + L.b1[i], kF, kF * kk, kF))) return rc;
        if ((rc = wn(2 + 2 * i, -1, packed + L.w2[i], packed + L.b2[i], kF, kF * kk, kF))) return rc;
    }
    if ((rc = bn(4, packed + L.bnO))) return rc;
    if ((rc = wn(5, -1, packed + L.wout, packed + L.bout, out_ch, kF, CoutPad))) return rc;
    if (!conv) return NFB_OK;
    const int rc_tc = pack_tc_launch(packed, packed + tc_plan(in_ch, out_ch, 0).base, in_ch, out_ch, 0, st);
    if (rc_tc != NFB_OK) return rc_tc;
    return pack_tc_launch(packed, packed + tc_plan(in_ch, out_ch, 1).base, in_ch, out_ch, 1, st);
}

extern "C" int nfb_convnet_fwd_ex(const float* src, float* params_out, const float* packed, int B, int C, int H, int W,
                                  int mode, int odd, int in_ch, int out_ch, int flags, nfb_stream_t stream) {
    if (!src || !params_out || !packed) return NFB_ERR_NULL;
    if (in_ch <= 0 || out_ch <= 0) return NFB_ERR_SHAPE;
    cudaStream_t st = as_stream(stream);
    SplitGeom g;
    if (mode < 0) {  // src is the (B, in_ch, H, W) conditioner input itself
        if (B <= 0 || H <= 0 || W <= 0) return NFB_ERR_SHAPE;
        g = SplitGeom{};
        return dispatch_convnet<-1>(src, params_out, packed, g, in_ch, out_ch, B, H, W, flags, st);
    }
    const int rc = make_geom(g, B, C, H, W, mode, odd);
    if (rc != NFB_OK) return rc;
    if (g.c0 != in_ch) return NFB_ERR_SHAPE;
    if (mode == NFB_SPLIT_CHECKER) return dispatch_convnet<NFB_SPLIT_CHECKER>(src, params_out, packed, g, in_ch, out_ch, B, g.h, g.w, flags, st);
    if (mode == NFB_SPLIT_CHANNEL) return dispatch_convnet<NFB_SPLIT_CHANNEL>(src, params_out, packed, g, in_ch, out_ch, B, g.h, g.w, flags, st);
    return NFB_ERR_SHAPE;
}

extern "C" int nfb_convnet_fwd(const float* src, float* params_out, const float* packed, int B, int C, int H, int W,
                               int mode, int odd, int in_ch, int out_ch, nfb_stream_t stream) {
    return nfb_convnet_fwd_ex(src, params_out, packed, B, C, H, W, mode, odd, in_ch, out_ch, 0, stream);
}

extern "C" int nfb_convnet_affine_fwd(float* z, float* ldj, const float* packed, const float* s_log_scale,
                                      const float* s_bias, int B, int C, int H, int W, int mode, int odd, int flags,
                                      nfb_stream_t stream) {
    if (!z || !ldj || !packed || !s_log_scale || !s_bias) return NFB_ERR_NULL;
    SplitGeom g;
    const int rc = make_geom(g, B, C, H, W, mode, odd);
    if (rc != NFB_OK) return rc;
    if (mode != NFB_SPLIT_CHECKER && mode != NFB_SPLIT_CHANNEL) return NFB_ERR_UNSUPPORTED;
    if (flags & NFB_CONV_FFMA) return NFB_ERR_UNSUPPORTED;
    return convnet_affine_tc_dispatch(z, ldj, packed, g, mode, g.c0, 2 * g.c0, B, s_log_scale,
                                      s_bias, flags, as_stream(stream));
}

extern "C" int nfb_convnet_affine_step_fwd(float* z, float* ldj, const float* packed, const float* s_log_scale,
                                           const float* s_bias, const float* next_log_scale, const float* next_bias,
                                           const float* next_W, const float* next_log_s, int B, int C, int H, int W, int mode,
                                           int odd, int flags, nfb_stream_t stream) {
    if (!z || !ldj || !packed || !s_log_scale || !s_bias || !next_log_scale || !next_bias || !next_W || !next_log_s)
        return NFB_ERR_NULL;
    SplitGeom g;
    const int rc = make_geom(g, B, C, H, W, mode, odd);
    if (rc != NFB_OK) return rc;
    if (mode != NFB_SPLIT_CHECKER && mode != NFB_SPLIT_CHANNEL) return NFB_ERR_UNSUPPORTED;
    if (flags & NFB_CONV_FFMA) return NFB_ERR_UNSUPPORTED;
    return convnet_affine_step_tc_dispatch(z, ldj, packed, g, mode, g.c0, 2 * g.c0, B,
                                           s_log_scale, s_bias, next_log_scale, next_bias, next_W, next_log_s, flags,
                                           as_stream(stream));
}

extern "C" int nfb_mlp_fwd(const float* src, float* params_out, const float* packed, int B, int C, int mode, int odd,
                           int in_ch, int out_ch, nfb_stream_t stream) {
    if (!src || !params_out || !packed) return NFB_ERR_NULL;
    if (in_ch <= 0 || out_ch <= 0 || B <= 0) return NFB_ERR_SHAPE;
    const PackLayout L = pack_layout(in_ch, out_ch, 1);
    const int threads = 256;
    const size_t smem = (static_cast<size_t>((L.total + 3) & ~3) + (threads / 32) * 32 * 33) * sizeof(float);
    if (smem > 220 * 1024) return NFB_ERR_UNSUPPORTED;
    cudaStream_t st = as_stream(stream);
    SplitGeom g{};
    int blocks = (B + threads - 1) / threads;
    if (blocks > kSMs * 4) blocks = kSMs * 4;
    if (mode < 0) {
        auto kern = mlp_fused_kernel<-1>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        kern<<<blocks, threads, smem, st>>>(src, params_out, packed, g, in_ch, out_ch, B);
        return launch_status();
    }
    const int rc = make_geom(g, B, C, 1, 1, NFB_SPLIT_1D, odd);
    if (rc != NFB_OK) return rc;
    if (mode != NFB_SPLIT_1D || g.c0 != in_ch) return NFB_ERR_SHAPE;
    auto kern = mlp_fused_kernel<NFB_SPLIT_1D>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    kern<<<blocks, threads, smem, st>>>(src, params_out, packed, g, in_ch, out_ch, B);
    return launch_status();
}
