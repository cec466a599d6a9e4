// conditioner_train.cu -- TRAIN-mode ConvNet conditioner (modules.py:416-438 under net.train(): BatchNorm on batch
// statistics, WeightNorm recomputed every step, weight_norm.py:35-45) and its backward pass, layer by layer.
//
// The eval-mode conditioner is ONE kernel because BatchNorm folds into the weights; in train mode every BatchNorm needs
// statistics over the whole batch, i.e. a grid-wide dependency between layers, so the network runs as a short chain of
// kernels over L2-resident (B, 32, h, w) activations (8 MB at B = 256, 16x16):
//   forward :  wn_pack  ->  conv(+bias,+skip,+channel moments)  ->  bn_relu  ->  conv ...
//   backward:  wgrad (+bias grad)  |  dgrad = the same conv kernel over flipped/transposed packed weights
//              -> bn_relu_bwd_reduce (ReLU mask + the two BatchNorm sums) -> bn_bwd_apply (+ residual add) ...
// The convolution core is the register-tiled FP32 FFMA routine of the fused kernel (conv3x3_acc, conditioner.cuh):
// OCT channels x 4 pixels per thread, zero rows above/below in shared memory, x neighbours by warp shuffle.
// Batch-wide sums (moments, BatchNorm backward sums) are accumulated in fp64 atomics; weight gradients in fp32 atomics.
#include "conditioner.cuh"

namespace nfb {

// ---------------------------------------------------------------------------------------------------------------------
// WeightNorm + packing.  v (O, I, KK), g (I, KK): w = v * g / (||v||_{dim 0} + eps)  (weight_norm.py:40)
//   w_nat (O, I, KK)                      natural layout (what the weight gradient refers to)
//   w_fwd [oc][ic][ci 32][tap][o 32]      forward conv: 32-output x 32-input chunks, zero padded (caller zero-fills)
//   w_bwd [ic'][oc'][o 32][KK-1-tap][i 32] data-gradient conv: roles of O and I swapped, taps flipped
// ---------------------------------------------------------------------------------------------------------------------
// block = 32 columns j (threadIdx.x) x 8 row slices (threadIdx.y: o = y, y+8, ...); the column norm is reduced through
// shared memory.  Padding entries of the packed layouts (o >= O or i >= I inside a 32 x 32 chunk) are written as zeros,
// so the caller needs no zero-fill.
constexpr int kMaxWnLayers = 8;
struct WnLayer {
    const float* v; const float* g; const float* gw;   // gw: weight gradient (backward only)
    float* o0; float* o1; float* o2;                   // pack: w_nat, w_fwd, w_bwd;  backward: gv, gg, unused
    int O, I, KK;
};
struct WnBatch { WnLayer l[kMaxWnLayers]; };

__device__ __forceinline__ void wn_pack_body(const float* __restrict__ v, const float* __restrict__ gw,
                                             float* __restrict__ w_nat, float* __restrict__ w_fwd,
                                             float* __restrict__ w_bwd, int O, int I, int KK, float eps) {
    __shared__ float part[8][33];
    const int J = I * KK;
    const int n_ic = (I + 31) >> 5, n_oc = (O + 31) >> 5;
    const int Jp = n_ic * 32 * KK, Op = n_oc * 32;
    const int j = blockIdx.x * 32 + threadIdx.x;   // column in the padded (I_pad * KK) range
    const int ty = threadIdx.y;
    const bool col = j < J;
    float ss = 0.f;
    if (col)
        for (int o = ty; o < O; o += 8) { const float x = __ldg(v + static_cast<size_t>(o) * J + j); ss = fmaf(x, x, ss); }
    part[ty][threadIdx.x] = ss;
    __syncthreads();
    if (j >= Jp) return;
    ss = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) ss += part[k][threadIdx.x];
    const float scale = col ? __fdiv_rn(__ldg(gw + j), __fadd_rn(sqrtf(ss), eps)) : 0.f;
    const int i = j / KK, tap = j - i * KK;
    const size_t chunk = static_cast<size_t>(32) * KK * 32;
    for (int o = ty; o < Op; o += 8) {
        float w = 0.f;
        if (col && o < O) {
            w = __fmul_rn(__ldg(v + static_cast<size_t>(o) * J + j), scale);
            w_nat[static_cast<size_t>(o) * J + j] = w;
        }
        w_fwd[(static_cast<size_t>(o >> 5) * n_ic + (i >> 5)) * chunk + ((i & 31) * KK + tap) * 32 + (o & 31)] = w;
        w_bwd[(static_cast<size_t>(i >> 5) * n_oc + (o >> 5)) * chunk + ((o & 31) * KK + (KK - 1 - tap)) * 32 + (i & 31)] = w;
    }
}
__global__ void __launch_bounds__(256) wn_pack_train_kernel(const float* __restrict__ v, const float* __restrict__ gw,
                                                           float* __restrict__ w_nat, float* __restrict__ w_fwd,
                                                           float* __restrict__ w_bwd, int O, int I, int KK, float eps) {
    wn_pack_body(v, gw, w_nat, w_fwd, w_bwd, O, I, KK, eps);
}
__global__ void __launch_bounds__(256) wn_pack_train_multi_kernel(WnBatch bt, float eps) {
    const WnLayer& l = bt.l[blockIdx.y];
    if (blockIdx.x * 32 >= ((l.I + 31) >> 5) * 32 * l.KK) return;  // block-uniform: this layer has fewer columns
    wn_pack_body(l.v, l.g, l.o0, l.o1, l.o2, l.O, l.I, l.KK, eps);
}

// gradient of the WeightNorm map: s_j = g_j / (n_j + eps), n_j = ||v[:, j]||, d_j = sum_o gw[o,j] v[o,j]
//   gg_j = d_j / (n_j + eps);   gv[o,j] = gw[o,j] s_j - v[o,j] d_j g_j / ((n_j + eps)^2 n_j)
__device__ __forceinline__ void wn_bwd_body(const float* __restrict__ v, const float* __restrict__ g,
                                            const float* __restrict__ gw, float* __restrict__ gv, float* __restrict__ gg,
                                            int O, int J, float eps) {
    __shared__ float p_ss[8][33], p_d[8][33];
    const int j = blockIdx.x * 32 + threadIdx.x;
    const int ty = threadIdx.y;
    float ss = 0.f, d = 0.f;
    if (j < J)
        for (int o = ty; o < O; o += 8) {
            const float x = __ldg(v + static_cast<size_t>(o) * J + j);
            ss = fmaf(x, x, ss);
            d = fmaf(__ldg(gw + static_cast<size_t>(o) * J + j), x, d);
        }
    p_ss[ty][threadIdx.x] = ss;
    p_d[ty][threadIdx.x] = d;
    __syncthreads();
    if (j >= J) return;
    ss = d = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { ss += p_ss[k][threadIdx.x]; d += p_d[k][threadIdx.x]; }
    const float n = sqrtf(ss), ne = n + eps, gj = __ldg(g + j);
    const float sc = gj / ne;
    const float c = n > 0.f ? d * gj / (ne * ne * n) : 0.f;
    if (ty == 0) gg[j] = d / ne;
    for (int o = ty; o < O; o += 8) {
        const size_t k = static_cast<size_t>(o) * J + j;
        gv[k] = fmaf(__ldg(gw + k), sc, -__ldg(v + k) * c);
    }
}
__global__ void __launch_bounds__(256) wn_bwd_kernel(const float* __restrict__ v, const float* __restrict__ g,
                                                    const float* __restrict__ gw, float* __restrict__ gv,
                                                    float* __restrict__ gg, int O, int J, float eps) {
    wn_bwd_body(v, g, gw, gv, gg, O, J, eps);
}
__global__ void __launch_bounds__(256) wn_bwd_multi_kernel(WnBatch bt, float eps) {
    const WnLayer& l = bt.l[blockIdx.y];
    if (blockIdx.x * 32 >= l.I * l.KK) return;
    wn_bwd_body(l.v, l.g, l.gw, l.o0, l.o1, l.O, l.I * l.KK, eps);
}

// ---------------------------------------------------------------------------------------------------------------------
// convolution layer (forward, and data gradient with w_bwd): out = conv_KS(in; w) + bias (+ skip), optional per-channel
// moments of the output (sum, sum of squares) for the BatchNorm that follows.  Cin / Cout in chunks of 32.
// ---------------------------------------------------------------------------------------------------------------------
struct ConvArgs {
    const float* in;     // (B, Cin, H, W)
    const float* w;      // packed chunks [oc][ic][32][KK][32]
    const float* bias;   // [Cout] or null
    const float* skip;   // (B, Cout, H, W) or null
    float* out;          // (B, Cout, H, W)
    double* stats;       // [2*Cout] (sum | sum of squares), accumulated; or null
    int Cin, Cout, B;
    // data-gradient mode (mask_a != null): the BatchNorm+ReLU backward reduction is fused into the epilogue --
    // out = U = conv * [mask_a > 0], stats = (sum U | sum U * xhat) with xhat = (mask_x - mean) * rstd
    const float* mask_a;
    const float* mask_x;
    const float* mean_rstd;
};

// TILES > 1: the image is H * TILES rows high and a CTA owns one horizontal band of H rows of one sample (32x32 inputs
// as four 8 x 32 bands); its halo rows come from the neighbouring bands in global memory instead of being zero.
template <int H, int W, int NT, int OCT, int KS, int TILES>
__global__ void __launch_bounds__(NT) conv_layer_kernel(ConvArgs A) {
    constexpr int KK = KS * KS;
    constexpr int HF = H * TILES;  // full image height
    constexpr int PGS = H * W / 4, NOG = kF / OCT, NPG = NT / NOG, S = NPG / PGS;
    static_assert(S >= 1 && S * PGS == NPG, "tile must hold whole samples");
    static_assert(TILES == 1 || S == 1, "banded images: one band per CTA");
    constexpr int CHS = S * (H + 2) * W;
    constexpr int RW = NPG < 32 ? NPG : 32;  // lanes that share one output-channel group inside a warp
    extern __shared__ __align__(16) float smem[];
    float* bufA = smem;              // [32][S][H+2][W]
    float* wsm = smem + kF * CHS;    // [32 ci][KK][32 o]
    const int t = threadIdx.x;
    const int pg = t % NPG, og = t / NPG;
    const int s = pg / PGS, r = pg % PGS;
    const int y = r / (W / 4), x0 = 4 * (r % (W / 4));
    const int b = TILES == 1 ? blockIdx.x * S + s : blockIdx.x / TILES;
    const int r0 = TILES == 1 ? 0 : (blockIdx.x % TILES) * H;  // first image row of this CTA's band
    const bool valid = b < A.B;
    const int sbase = s * (H + 2) * W;
    const int n_ic = (A.Cin + 31) >> 5, n_oc = (A.Cout + 31) >> 5;
    constexpr int HWv = H * W / 4;

    if (TILES == 1) {  // zero halo rows above / below every sample (banded images load theirs); channels past Cin are never read
        for (int i = t; i < kF * S * 2 * (W / 4); i += NT) {
            const int ci = i / (S * 2 * (W / 4));
            int rem = i - ci * (S * 2 * (W / 4));
            const int ss = rem / (2 * (W / 4));
            rem -= ss * (2 * (W / 4));
            const int row = rem / (W / 4) ? H + 1 : 0, xv = rem % (W / 4);
            st4(bufA + ci * CHS + ss * (H + 2) * W + row * W + 4 * xv, make_float4(0.f, 0.f, 0.f, 0.f));
        }
    }
    for (int oc = 0; oc < n_oc; ++oc) {
        float acc[OCT][4];
#pragma unroll
        for (int o = 0; o < OCT; ++o) acc[o][0] = acc[o][1] = acc[o][2] = acc[o][3] = 0.f;
        for (int ic = 0; ic < n_ic; ++ic) {
            __syncthreads();  // previous chunk fully consumed (also orders the zero fill)
            if (TILES > 1 && (n_ic > 1 || oc == 0)) {
                // band + its two halo rows: image rows r0-1 .. r0+H (zeros outside the image)
                constexpr int RWv = W / 4;
                for (int i = t; i < kF * (H + 2) * RWv; i += NT) {
                    const int ci = i / ((H + 2) * RWv);
                    const int rem = i - ci * ((H + 2) * RWv);
                    const int rr = rem / RWv, xv = rem - rr * RWv;
                    const int gr = r0 + rr - 1, ch = ic * 32 + ci;
                    float* dst = bufA + ci * CHS + rr * W + 4 * xv;
                    if (valid && ch < A.Cin && gr >= 0 && gr < HF)
                        cp_async16(dst, A.in + ((static_cast<size_t>(b) * A.Cin + ch) * HF + gr) * W + 4 * xv);
                    else
                        st4(dst, make_float4(0.f, 0.f, 0.f, 0.f));
                }
            } else if (n_ic > 1 || oc == 0) {
                // input chunk: channels [32 ic, 32 ic + 32) of the CTA's S samples, float4 per thread
                for (int i = t; i < kF * S * HWv; i += NT) {
                    const int ci = i / (S * HWv);
                    int rem = i - ci * (S * HWv);
                    const int ss = rem / HWv;
                    rem -= ss * HWv;
                    const int bb = blockIdx.x * S + ss, ch = ic * 32 + ci;
                    // asynchronous copies: all of a thread's requests are in flight together (no register staging);
                    // samples past the batch and channels past Cin are never consumed
                    if (bb < A.B && ch < A.Cin)
                        cp_async16(bufA + ci * CHS + ss * (H + 2) * W + W + 4 * rem,
                                   A.in + (static_cast<size_t>(bb) * A.Cin + ch) * (H * W) + 4 * rem);
                }
            }
            const float* wsrc = A.w + (static_cast<size_t>(oc) * n_ic + ic) * (32 * KK * 32);
            for (int i = t * 4; i < 32 * KK * 32; i += NT * 4) cp_async16(wsm + i, wsrc + i);
            cp_async_commit();
            cp_async_wait_all();
            __syncthreads();
            const int CI = (A.Cin - ic * 32) < kF ? (A.Cin - ic * 32) : kF;
            if (KS == 3) {
                conv3x3_acc<H, W, OCT>(acc, bufA + sbase, CHS, wsm, CI, og, y, x0);
            } else {
                const float* a = bufA + sbase + (y + 1) * W + x0;
#pragma unroll 4
                for (int ci = 0; ci < CI; ++ci) {
                    const float4 v = ld4(a + ci * CHS);
                    float wv[OCT];
                    load_w<OCT>(wv, wsm + ci * kF + og * OCT);
#pragma unroll
                    for (int o = 0; o < OCT; ++o) {
                        acc[o][0] = fmaf(wv[o], v.x, acc[o][0]);
                        acc[o][1] = fmaf(wv[o], v.y, acc[o][1]);
                        acc[o][2] = fmaf(wv[o], v.z, acc[o][2]);
                        acc[o][3] = fmaf(wv[o], v.w, acc[o][3]);
                    }
                }
            }
        }
        // epilogue of this chunk of 32 output channels
#pragma unroll
        for (int o = 0; o < OCT; ++o) {
            const int ch = oc * 32 + og * OCT + o;
            // moments in fp64 from the start: var = E[x^2] - mean^2 cancels when |mean| >> std (a bias in front of a
            // BatchNorm), and fp32 squares would leave the variance with ~1e-5 relative error
            double s1 = 0.0, s2 = 0.0;
            if (valid && ch < A.Cout) {
                const float bias = A.bias ? __ldg(A.bias + ch) : 0.f;
                float4 v = make_float4(acc[o][0] + bias, acc[o][1] + bias, acc[o][2] + bias, acc[o][3] + bias);
                const size_t off = ((static_cast<size_t>(b) * A.Cout + ch) * HF + r0 + y) * W + x0;
                if (A.skip) {
                    const float4 k = ldg4(A.skip + off);
                    v.x += k.x; v.y += k.y; v.z += k.z; v.w += k.w;
                }
                if (A.mask_a) {
                    const float4 a4 = ldg4(A.mask_a + off), x4 = ldg4(A.mask_x + off);
                    const float m = __ldg(A.mean_rstd + ch), rs = __ldg(A.mean_rstd + A.Cout + ch);
                    v.x = a4.x > 0.f ? v.x : 0.f; v.y = a4.y > 0.f ? v.y : 0.f;
                    v.z = a4.z > 0.f ? v.z : 0.f; v.w = a4.w > 0.f ? v.w : 0.f;
                    s2 = static_cast<double>(v.x) * ((x4.x - m) * rs) + static_cast<double>(v.y) * ((x4.y - m) * rs) +
                         static_cast<double>(v.z) * ((x4.z - m) * rs) + static_cast<double>(v.w) * ((x4.w - m) * rs);
                } else if (A.stats) {
                    s2 = static_cast<double>(v.x) * v.x + static_cast<double>(v.y) * v.y + static_cast<double>(v.z) * v.z +
                         static_cast<double>(v.w) * v.w;
                }
                st4(A.out + off, v);
                s1 = (static_cast<double>(v.x) + v.y) + (static_cast<double>(v.z) + v.w);
            }
            if (A.stats) {  // uniform branch
#pragma unroll
                for (int off = RW >> 1; off > 0; off >>= 1) {
                    s1 += __shfl_xor_sync(0xffffffffu, s1, off);
                    s2 += __shfl_xor_sync(0xffffffffu, s2, off);
                }
                if ((t & (RW - 1)) == 0 && ch < A.Cout) {
                    atomicAdd(A.stats + ch, s1);
                    atomicAdd(A.stats + A.Cout + ch, s2);
                }
            }
        }
    }
}

template <int H, int W, int NT, int OCT, int KS, int TILES = 1>
static int launch_conv_layer(const ConvArgs& A, cudaStream_t st) {
    constexpr int S = (NT / (kF / OCT)) / (H * W / 4);
    constexpr size_t smem = (static_cast<size_t>(kF) * S * (H + 2) * W + 32 * KS * KS * 32) * sizeof(float);
    auto kern = conv_layer_kernel<H, W, NT, OCT, KS, TILES>;
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    kern<<<TILES == 1 ? (A.B + S - 1) / S : A.B * TILES, NT, smem, st>>>(A);
    return launch_status();
}

// ---------------------------------------------------------------------------------------------------------------------
// weight gradient: gw[o, ci, tap] = sum_{b,y,x} gy[b,o,y,x] a[b,ci,y+dy-1,x+dx-1];  gb[o] = sum gy[b,o,y,x]
// CTA = one (32 o) x (32 ci) chunk pair x a group of samples; thread = 2 o x 2 ci x KK accumulators; one image row of gy
// and three rows of a live in registers at a time.  Every sample group writes its partial gradient (natural layout, then
// the bias partial) with plain stores; wgrad_reduce_kernel sums the groups -- no atomics, deterministic.
// ---------------------------------------------------------------------------------------------------------------------
template <int W>
__device__ __forceinline__ void load_row(float (&dst)[W], const float* __restrict__ p) {
#pragma unroll
    for (int i = 0; i < W; i += 4) {
        const float4 v = ld4(p + i);
        dst[i] = v.x; dst[i + 1] = v.y; dst[i + 2] = v.z; dst[i + 3] = v.w;
    }
}

// SPI samples are staged per iteration (16 image rows in every configuration: 1 x 16, 4 x 8 or 8 x 4), so the small
// feature maps do not pay two block barriers per 16 pixels.
// TILES > 1 (SPI = 1): a work unit is one horizontal band of H rows of one sample (image height H * TILES); the halo rows
// of `a` come from the neighbouring bands.  B then counts units (samples x TILES).
template <int H, int W, int KS, int SPI, int TILES>
__global__ void __launch_bounds__(256) wgrad_kernel(const float* __restrict__ gy, const float* __restrict__ a,
                                                   float* __restrict__ partial, int Cin, int Cout, int B, int spc) {
    constexpr int KK = KS * KS;
    constexpr int HF = H * TILES;
    static_assert(TILES == 1 || SPI == 1, "banded images: one band per iteration");
    constexpr int GS = SPI * H * W + 4;          // channel stride of the gy tile (stride/4 odd: conflict-free float4 reads)
    constexpr int AS = SPI * (H + 2) * W + 4;    // channel stride of the a tile (zero rows above / below every sample)
    static_assert(((GS / 4) & 1) == 1 && ((AS / 4) & 1) == 1, "padded strides");
    extern __shared__ __align__(16) float wg_smem[];
    float* gys = wg_smem;            // [32][SPI][H*W]
    float* as = wg_smem + 32 * GS;   // [32][SPI][(H+2)*W]
    const int n_ic = (Cin + 31) >> 5;
    const int oc = blockIdx.y / n_ic, ic = blockIdx.y - oc * n_ic;
    const int t = threadIdx.x;
    const int op = t >> 4, cp = t & 15;    // o in {op, op+16}, ci in {cp, cp+16}
    float acc[2][2][KK];
    float gbacc[2] = {0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int k = 0; k < KK; ++k) acc[i][j][k] = 0.f;
    for (int i = t; i < 32 * AS; i += 256) as[i] = 0.f;
    constexpr int HWv = H * W / 4;
    const int b0 = blockIdx.x * spc;
    const int b1 = (b0 + spc) < B ? (b0 + spc) : B;
    for (int b = b0; b < b1; b += SPI) {
        __syncthreads();
        if (TILES > 1) {
            const int smp = b / TILES, r0 = (b % TILES) * H;
            constexpr int RWv = W / 4;
            for (int i = t; i < 32 * (H + 2) * RWv; i += 256) {
                const int ch = i / ((H + 2) * RWv);
                const int rem = i - ch * ((H + 2) * RWv);
                const int rr = rem / RWv, xv = rem - rr * RWv;
                const int gr = r0 + rr - 1;  // image row of padded row rr
                float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ic * 32 + ch < Cin && gr >= 0 && gr < HF)
                    a4 = ldg4(a + ((static_cast<size_t>(smp) * Cin + ic * 32 + ch) * HF + gr) * W + 4 * xv);
                st4(as + ch * AS + rr * W + 4 * xv, a4);
                if (rr >= 1 && rr <= H) {
                    float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (oc * 32 + ch < Cout) g4 = ldg4(gy + ((static_cast<size_t>(smp) * Cout + oc * 32 + ch) * HF + gr) * W + 4 * xv);
                    st4(gys + ch * GS + (rr - 1) * W + 4 * xv, g4);
                }
            }
        } else
        for (int i = t; i < 32 * SPI * HWv; i += 256) {
            const int ch = i / (SPI * HWv);
            int rem = i - ch * (SPI * HWv);
            const int sp = rem / HWv;
            rem -= sp * HWv;
            float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f), a4 = g4;
            if (b + sp < b1) {
                if (oc * 32 + ch < Cout) g4 = ldg4(gy + (static_cast<size_t>(b + sp) * Cout + oc * 32 + ch) * (H * W) + 4 * rem);
                if (ic * 32 + ch < Cin) a4 = ldg4(a + (static_cast<size_t>(b + sp) * Cin + ic * 32 + ch) * (H * W) + 4 * rem);
            }
            st4(gys + ch * GS + sp * (H * W) + 4 * rem, g4);
            st4(as + ch * AS + sp * ((H + 2) * W) + W + 4 * rem, a4);
        }
        __syncthreads();
#pragma unroll 1
        for (int r = 0; r < SPI * H; ++r) {
            const int sp = r / H, y = r - sp * H;
            float gr[2][W];
            load_row<W>(gr[0], gys + op * GS + sp * (H * W) + y * W);
            load_row<W>(gr[1], gys + (op + 16) * GS + sp * (H * W) + y * W);
            if (cp == 0) {
#pragma unroll
                for (int x = 0; x < W; ++x) { gbacc[0] += gr[0][x]; gbacc[1] += gr[1][x]; }
            }
#pragma unroll
            for (int dy = 0; dy < KS; ++dy) {
                float ar[2][W];
                const int row = KS == 3 ? y + dy : y + 1;  // padded row index (row 0 = zeros above the image)
                load_row<W>(ar[0], as + cp * AS + sp * ((H + 2) * W) + row * W);
                load_row<W>(ar[1], as + (cp + 16) * AS + sp * ((H + 2) * W) + row * W);
#pragma unroll
                for (int dx = 0; dx < KS; ++dx) {
                    const int sh = KS == 3 ? dx - 1 : 0;
#pragma unroll
                    for (int x = 0; x < W; ++x) {
                        const int xx = x + sh;
                        if (xx < 0 || xx >= W) continue;
#pragma unroll
                        for (int i = 0; i < 2; ++i)
#pragma unroll
                            for (int j = 0; j < 2; ++j)
                                acc[i][j][dy * KS + dx] = fmaf(gr[i][x], ar[j][xx], acc[i][j][dy * KS + dx]);
                    }
                }
            }
        }
    }
    const size_t n_w = static_cast<size_t>(Cout) * Cin * KK;
    float* pw = partial + blockIdx.x * (n_w + Cout);  // this sample group's partial: gw (natural layout) | gb
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int o = oc * 32 + op + 16 * i;
        if (o >= Cout) continue;
        if (cp == 0 && ic == 0) pw[n_w + o] = gbacc[i];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int ci = ic * 32 + cp + 16 * j;
            if (ci >= Cin) continue;
#pragma unroll
            for (int k = 0; k < KK; ++k) pw[(static_cast<size_t>(o) * Cin + ci) * KK + k] = acc[i][j][k];
        }
    }
}

// gw[i] = sum over groups of partial[g][i] (i < n_w), gb[i - n_w] likewise
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ gw,
                                                          float* __restrict__ gb, int n_w, int n_b, int groups) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_w + n_b) return;
    const size_t stride = static_cast<size_t>(n_w) + n_b;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int g = 0;
    for (; g + 4 <= groups; g += 4) {
        a0 += __ldg(partial + (g + 0) * stride + i);
        a1 += __ldg(partial + (g + 1) * stride + i);
        a2 += __ldg(partial + (g + 2) * stride + i);
        a3 += __ldg(partial + (g + 3) * stride + i);
    }
    for (; g < groups; ++g) a0 += __ldg(partial + g * stride + i);
    const float r = (a0 + a1) + (a2 + a3);
    if (i < n_w) gw[i] = r;
    else if (gb) gb[i - n_w] = r;
}

// ---------------------------------------------------------------------------------------------------------------------
// BatchNorm (train) + ReLU, C channels, float4 never straddles a channel (HW % 4 == 0)
// ---------------------------------------------------------------------------------------------------------------------
// a = relu((x - mean) * rstd * gamma + beta) from the accumulated moments; the first CTA also stores (mean | rstd) for
// the backward pass and updates the running statistics (momentum; running_var takes the UNBIASED variance, as nn.BatchNorm).
__global__ void __launch_bounds__(256) bn_relu_fwd_kernel(const float* __restrict__ x, const double* __restrict__ stats,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         float* __restrict__ running_mean, float* __restrict__ running_var,
                                                         float momentum, float eps, float* __restrict__ a_out,
                                                         float* __restrict__ mean_rstd, int B, int C, int HW) {
    extern __shared__ float ks[];  // mean[C] | rstd[C] | gamma[C] | beta[C]
    const double n = static_cast<double>(B) * HW;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const double m = stats[c] / n;
        double var = stats[C + c] / n - m * m;
        var = var > 0.0 ? var : 0.0;
        const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
        ks[c] = static_cast<float>(m);
        ks[C + c] = rstd;
        ks[2 * C + c] = __ldg(gamma + c);
        ks[3 * C + c] = __ldg(beta + c);
        if (blockIdx.x == 0) {
            mean_rstd[c] = static_cast<float>(m);
            mean_rstd[C + c] = rstd;
            if (running_mean) {
                const double unb = n > 1.0 ? var * n / (n - 1.0) : var;
                running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * static_cast<float>(m);
                running_var[c] = (1.f - momentum) * running_var[c] + momentum * static_cast<float>(unb);
            }
        }
    }
    __syncthreads();
    const long long nv = static_cast<long long>(B) * C * HW / 4;
    for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < nv;
         v += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(((v << 2) / HW) % C);
        // subtract the mean first: x*k + (beta - mean*k) cancels when |mean| >> std; this is also exactly the xhat of
        // the backward pass, so the ReLU mask and the gradient see the same pre-activation
        const float m = ks[c], rs = ks[C + c], g = ks[2 * C + c], bt = ks[3 * C + c];
        const float4 q = ldg4(x + (v << 2));
        st4(a_out + (v << 2), make_float4(fmaxf(fmaf((q.x - m) * rs, g, bt), 0.f), fmaxf(fmaf((q.y - m) * rs, g, bt), 0.f),
                                          fmaxf(fmaf((q.z - m) * rs, g, bt), 0.f), fmaxf(fmaf((q.w - m) * rs, g, bt), 0.f)));
    }
}

// U = ga * [a > 0];  sums[c] += sum U,  sums[C + c] += sum U * xhat   (xhat = (x - mean) * rstd)
__global__ void __launch_bounds__(256) bn_relu_bwd_reduce_kernel(const float* __restrict__ ga, const float* __restrict__ a,
                                                                const float* __restrict__ x,
                                                                const float* __restrict__ mean_rstd, float* __restrict__ U,
                                                                double* __restrict__ sums, int B, int C, int HW) {
    extern __shared__ float acc[];  // 2C
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) acc[i] = 0.f;
    __syncthreads();
    const long long nv = static_cast<long long>(B) * C * HW / 4;
    const long long n_round = ((nv + 31) / 32) * 32;
    for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < n_round;
         v += static_cast<long long>(gridDim.x) * blockDim.x) {
        int c = -1;
        float s1 = 0.f, s2 = 0.f;
        if (v < nv) {
            c = static_cast<int>(((v << 2) / HW) % C);
            const float m = __ldg(mean_rstd + c), rs = __ldg(mean_rstd + C + c);
            const float4 g4 = ldg4(ga + (v << 2)), a4 = ldg4(a + (v << 2)), x4 = ldg4(x + (v << 2));
            const float4 u = make_float4(a4.x > 0.f ? g4.x : 0.f, a4.y > 0.f ? g4.y : 0.f, a4.z > 0.f ? g4.z : 0.f,
                                         a4.w > 0.f ? g4.w : 0.f);
            st4(U + (v << 2), u);
            s1 = (u.x + u.y) + (u.z + u.w);
            s2 = fmaf(u.x, (x4.x - m) * rs, fmaf(u.y, (x4.y - m) * rs, fmaf(u.z, (x4.z - m) * rs, u.w * ((x4.w - m) * rs))));
        }
        const int c0 = __shfl_sync(0xffffffffu, c, 0);
        if (__all_sync(0xffffffffu, c == c0)) {
            s1 = warp_sum(s1);
            s2 = warp_sum(s2);
            if ((threadIdx.x & 31) == 0 && c0 >= 0) { atomicAdd(acc + c0, s1); atomicAdd(acc + C + c0, s2); }
        } else if (c >= 0) {
            atomicAdd(acc + c, s1);
            atomicAdd(acc + C + c, s2);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x)
        if (acc[i] != 0.f) atomicAdd(sums + i, static_cast<double>(acc[i]));
}

// gx = gamma * rstd * (U - s1/n - xhat * s2/n) (+ add);  first CTA: g_gamma = s2, g_beta = s1
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ U, const float* __restrict__ x,
                                                          const float* __restrict__ mean_rstd,
                                                          const float* __restrict__ gamma, const double* __restrict__ sums,
                                                          const float* __restrict__ add, float* __restrict__ gx,
                                                          float* __restrict__ g_gamma, float* __restrict__ g_beta, int B,
                                                          int C, int HW) {
    extern __shared__ float cs[];  // k[C] | m1[C] | m2[C] | mean[C] | rstd[C]
    const double n = static_cast<double>(B) * HW;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float rs = __ldg(mean_rstd + C + c);
        cs[c] = __ldg(gamma + c) * rs;
        cs[C + c] = static_cast<float>(sums[c] / n);
        cs[2 * C + c] = static_cast<float>(sums[C + c] / n);
        cs[3 * C + c] = __ldg(mean_rstd + c);
        cs[4 * C + c] = rs;
        if (blockIdx.x == 0) {
            if (g_gamma) g_gamma[c] = static_cast<float>(sums[C + c]);
            if (g_beta) g_beta[c] = static_cast<float>(sums[c]);
        }
    }
    __syncthreads();
    const long long nv = static_cast<long long>(B) * C * HW / 4;
    for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < nv;
         v += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(((v << 2) / HW) % C);
        const float k = cs[c], m1 = cs[C + c], m2 = cs[2 * C + c], m = cs[3 * C + c], rs = cs[4 * C + c];
        const float4 u = ldg4(U + (v << 2)), x4 = ldg4(x + (v << 2));
        float4 o = make_float4(k * (u.x - m1 - (x4.x - m) * rs * m2), k * (u.y - m1 - (x4.y - m) * rs * m2),
                               k * (u.z - m1 - (x4.z - m) * rs * m2), k * (u.w - m1 - (x4.w - m) * rs * m2));
        if (add) {
            const float4 d = ldg4(add + (v << 2));
            o.x += d.x; o.y += d.y; o.z += d.z; o.w += d.w;
        }
        st4(gx + (v << 2), o);
    }
}

// (B, C) rows <-> (B/HW, C, HW) channel planes: lets the 1-D (MLP) conditioner run on the convolution-layer kernels (a
// Linear layer over rows is a 1x1 convolution over the "pixels" of a plane; BatchNorm1d statistics over the batch are
// the BatchNorm2d statistics over (group, pixel)).  32 x 32 tiles through shared memory, both sides coalesced.
template <bool TO_PLANES>
__global__ void __launch_bounds__(256) rows_planes_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int C,
                                                         int HW) {
    __shared__ float tile[32][33];
    const int g = blockIdx.z;                      // group of HW rows
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const size_t rbase = static_cast<size_t>(g) * HW;        // first row of the group
    if (TO_PLANES) {
        for (int k = ty; k < 32; k += 8) {         // read rows (p0+k), columns c0+tx
            const int p = p0 + k, c = c0 + tx;
            tile[k][tx] = (p < HW && c < C) ? __ldg(src + (rbase + p) * C + c) : 0.f;
        }
        __syncthreads();
        for (int k = ty; k < 32; k += 8) {         // write channel c0+k, pixels p0+tx
            const int c = c0 + k, p = p0 + tx;
            if (c < C && p < HW) dst[(static_cast<size_t>(g) * C + c) * HW + p] = tile[tx][k];
        }
    } else {
        for (int k = ty; k < 32; k += 8) {
            const int c = c0 + k, p = p0 + tx;
            tile[k][tx] = (c < C && p < HW) ? __ldg(src + (static_cast<size_t>(g) * C + c) * HW + p) : 0.f;
        }
        __syncthreads();
        for (int k = ty; k < 32; k += 8) {
            const int p = p0 + k, c = c0 + tx;
            if (p < HW && c < C) dst[(rbase + p) * C + c] = tile[tx][k];
        }
    }
}

inline int ew_grid(long long nv) {
    long long blocks = (nv + 255) / 256;
    if (blocks > kSMs * 8) blocks = kSMs * 8;
    if (blocks < 1) blocks = 1;
    return static_cast<int>(blocks);
}

}  // namespace nfb

using namespace nfb;

extern "C" int nfb_wn_pack_train(const float* v, const float* g, float* w_nat, float* w_fwd, float* w_bwd, int O, int I,
                                 int KK, float eps, nfb_stream_t stream) {
    if (!v || !g || !w_nat || !w_fwd || !w_bwd) return NFB_ERR_NULL;
    if (O <= 0 || I <= 0 || (KK != 1 && KK != 9)) return NFB_ERR_SHAPE;
    const int Jp = ((I + 31) / 32) * 32 * KK;
    wn_pack_train_kernel<<<(Jp + 31) / 32, dim3(32, 8), 0, as_stream(stream)>>>(v, g, w_nat, w_fwd, w_bwd, O, I, KK, eps);
    return launch_status();
}

// all WeightNorm layers of one conditioner in one launch.  ptrs: HOST array of 5 device pointers per layer
// (pack: v, g, w_nat, w_fwd, w_bwd; backward: v, g, gw, gv, gg); dims: HOST array of (O, I, KK) per layer.
static int wn_multi(const void* const* ptrs, const int* dims, int n, float eps, bool backward, nfb_stream_t stream) {
    if (!ptrs || !dims) return NFB_ERR_NULL;
    if (n <= 0 || n > kMaxWnLayers) return NFB_ERR_SHAPE;
    WnBatch bt;
    int max_cols = 0;
    for (int i = 0; i < n; ++i) {
        WnLayer& l = bt.l[i];
        for (int k = 0; k < 5; ++k)
            if (!ptrs[5 * i + k]) return NFB_ERR_NULL;
        l.O = dims[3 * i]; l.I = dims[3 * i + 1]; l.KK = dims[3 * i + 2];
        if (l.O <= 0 || l.I <= 0 || (l.KK != 1 && l.KK != 9)) return NFB_ERR_SHAPE;
        l.v = static_cast<const float*>(ptrs[5 * i]);
        l.g = static_cast<const float*>(ptrs[5 * i + 1]);
        if (!backward) {
            l.gw = nullptr;
            l.o0 = static_cast<float*>(const_cast<void*>(ptrs[5 * i + 2]));
            l.o1 = static_cast<float*>(const_cast<void*>(ptrs[5 * i + 3]));
            l.o2 = static_cast<float*>(const_cast<void*>(ptrs[5 * i + 4]));
        } else {
            l.gw = static_cast<const float*>(ptrs[5 * i + 2]);
            l.o0 = static_cast<float*>(const_cast<void*>(ptrs[5 * i + 3]));
            l.o1 = static_cast<float*>(const_cast<void*>(ptrs[5 * i + 4]));
            l.o2 = nullptr;
        }
        const int cols = backward ? l.I * l.KK : ((l.I + 31) / 32) * 32 * l.KK;
        if (cols > max_cols) max_cols = cols;
    }
    dim3 grid((max_cols + 31) / 32, n);
    if (backward) wn_bwd_multi_kernel<<<grid, dim3(32, 8), 0, as_stream(stream)>>>(bt, eps);
    else wn_pack_train_multi_kernel<<<grid, dim3(32, 8), 0, as_stream(stream)>>>(bt, eps);
    return launch_status();
}
extern "C" int nfb_wn_pack_train_multi(const void* const* ptrs, const int* dims, int n, float eps, nfb_stream_t stream) {
    return wn_multi(ptrs, dims, n, eps, false, stream);
}
extern "C" int nfb_wn_bwd_multi(const void* const* ptrs, const int* dims, int n, float eps, nfb_stream_t stream) {
    return wn_multi(ptrs, dims, n, eps, true, stream);
}

extern "C" int nfb_wn_bwd(const float* v, const float* g, const float* gw, float* gv, float* gg, int O, int Ikk, float eps,
                          nfb_stream_t stream) {
    if (!v || !g || !gw || !gv || !gg) return NFB_ERR_NULL;
    if (O <= 0 || Ikk <= 0) return NFB_ERR_SHAPE;
    wn_bwd_kernel<<<(Ikk + 31) / 32, dim3(32, 8), 0, as_stream(stream)>>>(v, g, gw, gv, gg, O, Ikk, eps);
    return launch_status();
}

static int conv_dispatch(const ConvArgs& A, int h, int w, int ks, cudaStream_t st) {
#define NFB_CL(H_, W_, NT_, OCT_)                                                     \
    return ks == 3 ? launch_conv_layer<H_, W_, NT_, OCT_, 3>(A, st) : launch_conv_layer<H_, W_, NT_, OCT_, 1>(A, st)
    if (h == 16 && w == 16) { NFB_CL(16, 16, 256, 8); }
    if (h == 8 && w == 8) { NFB_CL(8, 8, 128, 4); }
    if (h == 4 && w == 4) { NFB_CL(4, 4, 128, 2); }
#undef NFB_CL
    if (h == 32 && w == 32)  // four 8 x 32 bands per sample
        return ks == 3 ? launch_conv_layer<8, 32, 256, 8, 3, 4>(A, st) : launch_conv_layer<8, 32, 256, 8, 1, 4>(A, st);
    return NFB_ERR_UNSUPPORTED;
}

extern "C" int nfb_conv_train(const float* in, const float* w_packed, const float* bias, const float* skip, float* out,
                              double* stats, int stats_zeroed, int B, int Cin, int Cout, int h, int w, int ks,
                              nfb_stream_t stream) {
    if (!in || !w_packed || !out) return NFB_ERR_NULL;
    if (B <= 0 || Cin <= 0 || Cout <= 0 || (ks != 1 && ks != 3)) return NFB_ERR_SHAPE;
    if (!aligned16(in) || !aligned16(out) || !aligned16(w_packed) || (skip && !aligned16(skip))) return NFB_ERR_UNSUPPORTED;
    cudaStream_t st = as_stream(stream);
    if (stats && !stats_zeroed) cudaMemsetAsync(stats, 0, sizeof(double) * 2 * Cout, st);
    const ConvArgs A{in, w_packed, bias, skip, out, stats, Cin, Cout, B, nullptr, nullptr, nullptr};
    return conv_dispatch(A, h, w, ks, st);
}

extern "C" int nfb_conv_train_dgrad_bnrelu(const float* gy, const float* w_bwd, const float* a, const float* x,
                                           const float* mean_rstd, float* U, double* sums, int sums_zeroed, int B, int Cin,
                                           int Cout, int h, int w, int ks, nfb_stream_t stream) {
    if (!gy || !w_bwd || !a || !x || !mean_rstd || !U || !sums) return NFB_ERR_NULL;
    if (B <= 0 || Cin <= 0 || Cout <= 0 || (ks != 1 && ks != 3)) return NFB_ERR_SHAPE;
    if (!aligned16(gy) || !aligned16(w_bwd) || !aligned16(a) || !aligned16(x) || !aligned16(U)) return NFB_ERR_UNSUPPORTED;
    cudaStream_t st = as_stream(stream);
    if (!sums_zeroed) cudaMemsetAsync(sums, 0, sizeof(double) * 2 * Cout, st);
    const ConvArgs A{gy, w_bwd, nullptr, nullptr, U, sums, Cin, Cout, B, a, x, mean_rstd};
    return conv_dispatch(A, h, w, ks, st);
}

// samples per CTA: large enough that the partial-sum traffic (groups x |gw|) stays small next to the FMA work
// work units: samples, or bands of 8 rows for 32x32 inputs
static int wgrad_units(int B, int h, int w) { return (h == 32 && w == 32) ? B * 4 : B; }

static int wgrad_spc(int B, int Cin, int Cout, int h, int w) {
    const int pairs = ((Cin + 31) / 32) * ((Cout + 31) / 32);
    B = wgrad_units(B, h, w);
    int spc = h * w >= 256 ? 2 : (h * w >= 64 ? 4 : 8);  // = a multiple of the samples staged per iteration
    const int fill = (B * pairs + kSMs * 2 - 1) / (kSMs * 2);  // never more than about two CTAs per SM in total
    if (spc < fill) spc = fill;
    if (spc > B) spc = B;
    return spc < 1 ? 1 : spc;
}

extern "C" long long nfb_conv_train_wgrad_scratch(int B, int Cin, int Cout, int h, int w, int ks) {
    if (B <= 0 || Cin <= 0 || Cout <= 0 || (ks != 1 && ks != 3)) return NFB_ERR_SHAPE;
    const int spc = wgrad_spc(B, Cin, Cout, h, w);
    const long long groups = (wgrad_units(B, h, w) + spc - 1) / spc;
    return groups * (static_cast<long long>(Cout) * Cin * ks * ks + Cout);
}

extern "C" int nfb_conv_train_wgrad(const float* gy, const float* a, float* gw, float* gb, float* scratch, int B, int Cin,
                                    int Cout, int h, int w, int ks, nfb_stream_t stream) {
    if (!gy || !a || !gw || !scratch) return NFB_ERR_NULL;
    if (B <= 0 || Cin <= 0 || Cout <= 0 || (ks != 1 && ks != 3)) return NFB_ERR_SHAPE;
    if (!aligned16(gy) || !aligned16(a)) return NFB_ERR_UNSUPPORTED;
    cudaStream_t st = as_stream(stream);
    const int pairs = ((Cin + 31) / 32) * ((Cout + 31) / 32);
    const int spc = wgrad_spc(B, Cin, Cout, h, w);
    const int units = wgrad_units(B, h, w);
    const int groups = (units + spc - 1) / spc;
    dim3 grid(groups, pairs);
    int rc = NFB_ERR_UNSUPPORTED;
#define NFB_WG(H_, W_, SPI_, TILES_)                                                                         \
    do {                                                                                                     \
        constexpr size_t smem = sizeof(float) * 32 * ((SPI_) * (H_) * (W_) + 4 + (SPI_) * ((H_) + 2) * (W_) + 4); \
                    cudaFuncSetAttribute(wgrad_kernel<H_, W_, 3, SPI_, TILES_>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)); \
            cudaFuncSetAttribute(wgrad_kernel<H_, W_, 1, SPI_, TILES_>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)); \
        if (ks == 3) wgrad_kernel<H_, W_, 3, SPI_, TILES_><<<grid, 256, smem, st>>>(gy, a, scratch, Cin, Cout, units, spc); \
        else wgrad_kernel<H_, W_, 1, SPI_, TILES_><<<grid, 256, smem, st>>>(gy, a, scratch, Cin, Cout, units, spc); \
        rc = launch_status();                                                                                \
    } while (0)
    if (h == 16 && w == 16) NFB_WG(16, 16, 1, 1);
    else if (h == 8 && w == 8) NFB_WG(8, 8, 4, 1);
    else if (h == 4 && w == 4) NFB_WG(4, 4, 8, 1);
    else if (h == 32 && w == 32) NFB_WG(8, 32, 1, 4);
#undef NFB_WG
    if (rc != NFB_OK) return rc;
    const int n_w = Cout * Cin * ks * ks;
    wgrad_reduce_kernel<<<(n_w + Cout + 255) / 256, 256, 0, st>>>(scratch, gw, gb, n_w, Cout, groups);
    return launch_status();
}

extern "C" int nfb_bn_relu_fwd(const float* x, const double* stats, const float* gamma, const float* beta,
                               float* running_mean, float* running_var, float momentum, float eps, float* a_out,
                               float* mean_rstd, int B, int C, int HW, nfb_stream_t stream) {
    if (!x || !stats || !gamma || !beta || !a_out || !mean_rstd) return NFB_ERR_NULL;
    if (B <= 0 || C <= 0 || HW <= 0) return NFB_ERR_SHAPE;
    if (HW % 4 || C > 1024 || !aligned16(x) || !aligned16(a_out)) return NFB_ERR_UNSUPPORTED;
    const long long nv = static_cast<long long>(B) * C * HW / 4;
    bn_relu_fwd_kernel<<<ew_grid(nv), 256, sizeof(float) * 4 * C, as_stream(stream)>>>(
        x, stats, gamma, beta, running_mean, running_var, momentum, eps, a_out, mean_rstd, B, C, HW);
    return launch_status();
}

extern "C" int nfb_bn_relu_bwd_reduce(const float* ga, const float* a, const float* x, const float* mean_rstd, float* U,
                                      double* sums, int sums_zeroed, int B, int C, int HW, nfb_stream_t stream) {
    if (!ga || !a || !x || !mean_rstd || !U || !sums) return NFB_ERR_NULL;
    if (B <= 0 || C <= 0 || HW <= 0) return NFB_ERR_SHAPE;
    if (HW % 4 || C > 1024 || !aligned16(ga) || !aligned16(a) || !aligned16(x) || !aligned16(U)) return NFB_ERR_UNSUPPORTED;
    cudaStream_t st = as_stream(stream);
    if (!sums_zeroed) cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, st);
    const long long nv = static_cast<long long>(B) * C * HW / 4;
    bn_relu_bwd_reduce_kernel<<<ew_grid(nv), 256, sizeof(float) * 2 * C, st>>>(ga, a, x, mean_rstd, U, sums, B, C, HW);
    return launch_status();
}

extern "C" int nfb_bn_bwd_apply(const float* U, const float* x, const float* mean_rstd, const float* gamma,
                                const double* sums, const float* add, float* gx, float* g_gamma, float* g_beta, int B, int C,
                                int HW, nfb_stream_t stream) {
    if (!U || !x || !mean_rstd || !gamma || !sums || !gx) return NFB_ERR_NULL;
    if (B <= 0 || C <= 0 || HW <= 0) return NFB_ERR_SHAPE;
    if (HW % 4 || C > 1024 || !aligned16(U) || !aligned16(x) || !aligned16(gx) || (add && !aligned16(add)))
        return NFB_ERR_UNSUPPORTED;
    const long long nv = static_cast<long long>(B) * C * HW / 4;
    bn_bwd_apply_kernel<<<ew_grid(nv), 256, sizeof(float) * 5 * C, as_stream(stream)>>>(U, x, mean_rstd, gamma, sums, add, gx,
                                                                                      g_gamma, g_beta, B, C, HW);
    return launch_status();
}

extern "C" int nfb_rows_to_planes(const float* rows, float* planes, int B, int C, int HW, nfb_stream_t stream) {
    if (!rows || !planes) return NFB_ERR_NULL;
    if (B <= 0 || C <= 0 || HW <= 0 || B % HW) return NFB_ERR_SHAPE;
    if (B / HW > 65535) return NFB_ERR_UNSUPPORTED;
    dim3 grid((HW + 31) / 32, (C + 31) / 32, B / HW);
    rows_planes_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(rows, planes, B, C, HW);
    return launch_status();
}

extern "C" int nfb_planes_to_rows(const float* planes, float* rows, int B, int C, int HW, nfb_stream_t stream) {
    if (!rows || !planes) return NFB_ERR_NULL;
    if (B <= 0 || C <= 0 || HW <= 0 || B % HW) return NFB_ERR_SHAPE;
    if (B / HW > 65535) return NFB_ERR_UNSUPPORTED;
    dim3 grid((HW + 31) / 32, (C + 31) / 32, B / HW);
    rows_planes_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(planes, rows, B, C, HW);
    return launch_status();
}
