// flowpp_mlp.cu -- the Flow++ conditioner of the 1-D couplings (coupling.py:142-158: Linear(in,32) -> GatedLinear ->
// LayerNorm(32) -> GatedAttn((32,)) -> LayerNorm(32) -> Linear(32, (2+3K) c0); modules.py:500-517, 538-578) as ONE kernel.
//
// A 1-D input is a single "token" for the attention block: the score matrix V^T K is 1 x 1, its softmax over dim 2 is
// exactly 1, so A = Q and GatedAttn reduces to  q = W1[Q rows] (x + pos) + b1[Q rows];  y = W2 q + b2;
// x += y[:32] * sigmoid(y[32:])  (the V and K projections do not influence the result; (V, K, Q) split order of
// modules.py:566).  Everything is per sample: one thread owns one sample, its 32 hidden features live in registers, the
// whole network (a few thousand floats for the 2-D toy densities the reference trains these on) sits in shared memory
// and is read with warp-broadcast loads.
#include "common.cuh"

namespace nfb {

constexpr int kH = 32;  // base_filters

struct FppMlpArgs {
    const float *w0, *b0, *wg, *bg, *ln1w, *ln1b, *pos, *w1, *b1, *w2, *b2, *ln2w, *ln2b, *w5, *b5;
};

__device__ __forceinline__ float elu_f(float x) { return x > 0.f ? x : expm1f(x); }            // F.elu, alpha = 1
__device__ __forceinline__ float sigmoid_f(float x) { return __fdiv_rn(1.f, 1.f + expf(-x)); }  // torch.sigmoid

__device__ __forceinline__ void layer_norm32(float (&h)[kH], const float* __restrict__ w, const float* __restrict__ b) {
    float m = 0.f;
#pragma unroll
    for (int i = 0; i < kH; ++i) m += h[i];
    m *= (1.f / kH);
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < kH; ++i) { const float d = h[i] - m; v = fmaf(d, d, v); }
    const float rs = __fdiv_rn(1.f, sqrtf(v * (1.f / kH) + 1.0e-5f));  // biased variance, eps inside the root
#pragma unroll
    for (int i = 0; i < kH; ++i) h[i] = fmaf((h[i] - m) * rs, w[i], b[i]);
}

template <int MODE>  // NFB_SPLIT_1D: z1 gathered from z (row stride D); MODE < 0: src is (B, in_ch)
__global__ void __launch_bounds__(128) flowpp_mlp_kernel(const float* __restrict__ src, float* __restrict__ out, FppMlpArgs A,
                                                        int B, int D, int odd, int Cin, int Cout) {
    extern __shared__ __align__(16) float sm[];
    // shared-memory image of the network
    float* w0 = sm;                       // [32][Cin]
    float* b0 = w0 + kH * Cin;            // [32]
    float* wg = b0 + kH;                  // [32][64]
    float* bg = wg + kH * 2 * kH;         // [32]
    float* l1w = bg + kH;                 // [32] x 2
    float* l1b = l1w + kH;
    float* pos = l1b + kH;                // [32]
    float* wq = pos + kH;                 // [32][32]  rows 64..95 of conv1
    float* bq = wq + kH * kH;             // [32]
    float* w2 = bq + kH;                  // [64][32]
    float* b2 = w2 + 2 * kH * kH;         // [64]
    float* l2w = b2 + 2 * kH;             // [32] x 2
    float* l2b = l2w + kH;
    float* w5 = l2b + kH;                 // [Cout][32]
    float* b5 = w5 + static_cast<size_t>(Cout) * kH;  // [Cout]
    const int t = threadIdx.x, nt = blockDim.x;
    for (int i = t; i < kH * Cin; i += nt) w0[i] = __ldg(A.w0 + i);
    for (int i = t; i < kH * 2 * kH; i += nt) wg[i] = __ldg(A.wg + i);
    for (int i = t; i < kH * kH; i += nt) wq[i] = __ldg(A.w1 + 2 * kH * kH + i);  // Q = third group of conv1 rows
    for (int i = t; i < 2 * kH * kH; i += nt) w2[i] = __ldg(A.w2 + i);
    for (int i = t; i < Cout * kH; i += nt) w5[i] = __ldg(A.w5 + i);
    for (int i = t; i < Cout; i += nt) b5[i] = __ldg(A.b5 + i);
    for (int i = t; i < kH; i += nt) {
        b0[i] = __ldg(A.b0 + i); bg[i] = __ldg(A.bg + i); l1w[i] = __ldg(A.ln1w + i); l1b[i] = __ldg(A.ln1b + i);
        pos[i] = __ldg(A.pos + i); bq[i] = __ldg(A.b1 + 2 * kH + i); l2w[i] = __ldg(A.ln2w + i); l2b[i] = __ldg(A.ln2b + i);
    }
    for (int i = t; i < 2 * kH; i += nt) b2[i] = __ldg(A.b2 + i);
    __syncthreads();

    for (int b = blockIdx.x * nt + t; b < B; b += gridDim.x * nt) {
        float h[kH], y[kH];
        // Linear(in, 32)
#pragma unroll
        for (int o = 0; o < kH; ++o) h[o] = b0[o];
        for (int ci = 0; ci < Cin; ++ci) {
            const float v = MODE < 0 ? __ldg(src + static_cast<size_t>(b) * Cin + ci)
                                     : __ldg(src + static_cast<size_t>(b) * D + 2 * ci + (odd ? 0 : 1));  // z1 of squeeze1d
#pragma unroll
            for (int o = 0; o < kH; ++o) h[o] = fmaf(w0[o * Cin + ci], v, h[o]);
        }
        // GatedLinear (modules.py:500-517): y = W elu([x, -x]) + b;  y = elu([y, -y]);  x += y[:32] * sigmoid(y[32:])
#pragma unroll
        for (int o = 0; o < kH; ++o) y[o] = bg[o];
#pragma unroll 4
        for (int i = 0; i < kH; ++i) {
            const float p = elu_f(h[i]), q = elu_f(-h[i]);
#pragma unroll
            for (int o = 0; o < kH; ++o) y[o] = fmaf(wg[o * 2 * kH + kH + i], q, fmaf(wg[o * 2 * kH + i], p, y[o]));
        }
#pragma unroll
        for (int o = 0; o < kH; ++o) h[o] = fmaf(elu_f(y[o]), sigmoid_f(elu_f(-y[o])), h[o]);
        layer_norm32(h, l1w, l1b);
        // GatedAttn with a single token: A = Q
#pragma unroll
        for (int o = 0; o < kH; ++o) y[o] = bq[o];
#pragma unroll 4
        for (int i = 0; i < kH; ++i) {
            const float xi = h[i] + pos[i];
#pragma unroll
            for (int o = 0; o < kH; ++o) y[o] = fmaf(wq[o * kH + i], xi, y[o]);
        }
        {
            float ya[kH], yb[kH];
#pragma unroll
            for (int o = 0; o < kH; ++o) { ya[o] = b2[o]; yb[o] = b2[kH + o]; }
#pragma unroll 4
            for (int i = 0; i < kH; ++i) {
#pragma unroll
                for (int o = 0; o < kH; ++o) {
                    ya[o] = fmaf(w2[o * kH + i], y[i], ya[o]);
                    yb[o] = fmaf(w2[(kH + o) * kH + i], y[i], yb[o]);
                }
            }
#pragma unroll
            for (int o = 0; o < kH; ++o) h[o] = fmaf(ya[o], sigmoid_f(yb[o]), h[o]);
        }
        layer_norm32(h, l2w, l2b);
        // Linear(32, Cout)
        float* orow = out + static_cast<size_t>(b) * Cout;
        for (int o = 0; o < Cout; ++o) {
            float acc = b5[o];
#pragma unroll
            for (int i = 0; i < kH; ++i) acc = fmaf(w5[o * kH + i], h[i], acc);
            orow[o] = acc;
        }
    }
}

}  // namespace nfb

using namespace nfb;

extern "C" int nfb_flowpp_mlp_fwd(const float* const* tensors, const float* src, float* params_out, int B, int D, int mode,
                                  int odd, int in_ch, int out_ch, nfb_stream_t stream) {
    if (!tensors || !src || !params_out) return NFB_ERR_NULL;
    for (int i = 0; i < 15; ++i)
        if (!tensors[i]) return NFB_ERR_NULL;
    if (B <= 0 || in_ch <= 0 || out_ch <= 0) return NFB_ERR_SHAPE;
    if (mode >= 0 && (mode != NFB_SPLIT_1D || D <= 0 || D % 2 || D / 2 != in_ch)) return NFB_ERR_SHAPE;
    const size_t floats = static_cast<size_t>(kH) * in_ch + kH * 2 * kH + kH * kH + 2 * kH * kH +
                          static_cast<size_t>(out_ch) * (kH + 1) + 10 * kH;
    const size_t smem = floats * sizeof(float);
    if (smem > 200 * 1024) return NFB_ERR_UNSUPPORTED;  // network does not fit one SM's shared memory: library path
    const FppMlpArgs A{tensors[0], tensors[1], tensors[2],  tensors[3],  tensors[4],  tensors[5],  tensors[6], tensors[7],
                       tensors[8], tensors[9], tensors[10], tensors[11], tensors[12], tensors[13], tensors[14]};
    cudaStream_t st = as_stream(stream);
    int grid = (B + 127) / 128;
    if (grid > kSMs * 4) grid = kSMs * 4;
    if (mode < 0) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(flowpp_mlp_kernel<-1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        flowpp_mlp_kernel<-1><<<grid, 128, smem, st>>>(src, params_out, A, B, D, odd, in_ch, out_ch);
    } else {
        if (smem > 48 * 1024) cudaFuncSetAttribute(flowpp_mlp_kernel<NFB_SPLIT_1D>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        flowpp_mlp_kernel<NFB_SPLIT_1D><<<grid, 128, smem, st>>>(src, params_out, A, B, D, odd, in_ch, out_ch);
    }
    return launch_status();
}
