// common.cuh -- shared device helpers for libnfb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "nfb200.h"

namespace nfb {

constexpr int kSMs = 148;  // B200

extern unsigned long long g_launches;  // host-side counter behind nfb_launch_count()

inline int launch_status() {
    ++g_launches;
    return static_cast<int>(cudaPeekAtLastError());
}

inline cudaStream_t as_stream(nfb_stream_t s) { return static_cast<cudaStream_t>(s); }

__host__ __device__ inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---------------------------------------------------------------------------------------------
// geometry of the coupling split (squeeze.py) -- see DESIGN.md "index formulas"
// ---------------------------------------------------------------------------------------------
struct SplitGeom {
    int B, C, H, W;
    int HW;   // H*W
    int D;    // C*H*W   elements per sample
    int n0;   // D/2     elements per sample in each half
    int h, w; // spatial size of the halves (H/2,W/2 for checker; H,W for channel; 1,1 for 1-D)
    int c0;   // channels of each half (2C checker, C/2 channel / 1-D)
    int odd;
};

inline int make_geom(SplitGeom& g, int B, int C, int H, int W, int mode, int odd) {
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return NFB_ERR_SHAPE;
    g.B = B; g.C = C; g.H = H; g.W = W; g.HW = H * W;
    long long D = 1LL * C * H * W;
    if (D > (1LL << 30)) return NFB_ERR_SHAPE;
    g.D = static_cast<int>(D); g.n0 = g.D / 2; g.odd = odd ? 1 : 0;
    switch (mode) {
        case NFB_SPLIT_1D:
            if (H != 1 || W != 1) return NFB_ERR_SHAPE;
            if (C % 2) return NFB_ERR_SPLIT;
            g.h = g.w = 1; g.c0 = C / 2; break;
        case NFB_SPLIT_CHECKER:
            if ((H % 2) || (W % 2)) return NFB_ERR_SPLIT;
            g.h = H / 2; g.w = W / 2; g.c0 = 2 * C; break;
        case NFB_SPLIT_CHANNEL:
            if (C % 2) return NFB_ERR_SPLIT;
            g.h = H; g.w = W; g.c0 = C / 2; break;
        default: return NFB_ERR_SHAPE;
    }
    return NFB_OK;
}

// Which half does element e (offset inside one sample, original layout) belong to, and where does it
// sit inside that half's contiguous (c0, h, w) tensor?  Returns true for the transformed half z0.
template <int MODE>
__device__ __forceinline__ bool classify(const SplitGeom& g, int e, int& idx) {
    if (MODE == NFB_SPLIT_1D) {
        idx = e >> 1;
        return ((e & 1) ^ g.odd) == 0;
    } else if (MODE == NFB_SPLIT_CHANNEL) {
        const bool first = e < g.n0;
        idx = first ? e : e - g.n0;
        return first != (g.odd != 0);
    } else {
        const int c = e / g.HW;
        const int r = e - c * g.HW;
        const int y = r / g.W;
        const int x = r - y * g.W;
        const int k = 4 * c + 2 * (y & 1) + (x & 1);
        const int q = k / g.C;  // which of the 4 blocks of C squeezed channels
        const int m = (q == 0) ? k : (q == 3) ? k - 2 * g.C : k - g.C;
        idx = (m * g.h + (y >> 1)) * g.w + (x >> 1);
        return ((q == 0 || q == 3) ? 1 : 0) != g.odd;
    }
}

// block classification for the checkerboard: squeezed channel k -> (is z0, channel m inside its half)
__device__ __forceinline__ bool checker_block(const SplitGeom& g, int k, int& m) {
    const int q = k / g.C;
    m = (q == 0) ? k : (q == 3) ? k - 2 * g.C : k - g.C;
    return ((q == 0 || q == 3) ? 1 : 0) != g.odd;
}

// Inverse map: entry j of half `second` (0 = transformed half z0, 1 = pass-through half z1), both stored as a
// contiguous (c0, h, w) tensor, -> offset inside one sample in the original layout.
template <int MODE>
__device__ __forceinline__ int half_offset(const SplitGeom& g, int j, int second) {
    const bool outer = (second ^ g.odd) == 0;  // even entries / first channel half / blocks (a, d)
    if (MODE == NFB_SPLIT_1D) return 2 * j + (outer ? 0 : 1);
    if (MODE == NFB_SPLIT_CHANNEL) return j + (outer ? 0 : g.n0);
    const int hw = g.h * g.w;
    const int m = j / hw;
    const int r = j - m * hw;
    const int i = r / g.w;
    const int jj = r - i * g.w;
    const int k = outer ? (m < g.C ? m : m + 2 * g.C) : m + g.C;
    return (k >> 2) * g.HW + (2 * i + ((k >> 1) & 1)) * g.W + 2 * jj + (k & 1);
}

// ---------------------------------------------------------------------------------------------
// vector memory helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }

// ---------------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the whole CTA; result valid in every thread.  `red` = 33 floats of shared memory.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();  // protect `red` from the previous use
    if (lane == 0) red[wid] = v;
    __syncthreads();
    if (wid == 0) {
        T t = lane < nw ? red[lane] : T(0);
        t = warp_sum(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

// ---------------------------------------------------------------------------------------------
// scalar math with PyTorch's CPU semantics (no fast-math anywhere in this library)
// ---------------------------------------------------------------------------------------------
// F.softplus(beta=1, threshold=20): x if x > 20 else log1p(exp(x))   
__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }
// F.logsigmoid: min(x,0) - log1p(exp(-|x|))
__device__ __forceinline__ float logsigmoid_f(float x) { return fminf(x, 0.f) - log1pf(expf(-fabsf(x))); }
// modules.py:19-21
__device__ __forceinline__ float log_dsigmoid_f(float x) { return __fsub_rn(x, __fmul_rn(2.f, softplus_f(x))); }

// ---------------------------------------------------------------------------------------------
// "row" drivers: one sample (row) reduces its own log-det contribution.
//   F: struct with  int items;  __device__ float operator()(int row, int item) const;
//                   __device__ float finish(float acc) const;   (sign / scaling of the row sum)
// rows_big:   one CTA per row (grid-stride), block-level reduce -- deterministic, no atomics.
// rows_small: G (power of two <= 32) lanes per row, shuffle reduce.
// ---------------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(512) rows_big_kernel(F f, const float* ldj_in, float* ldj_out, int B) {
    __shared__ float red[33];
    for (int row = blockIdx.x; row < B; row += gridDim.x) {
        float acc = 0.f;
        for (int it = threadIdx.x; it < f.items; it += blockDim.x) acc += f(row, it);
        acc = block_sum(acc, red);
        if (threadIdx.x == 0 && ldj_out) ldj_out[row] = ldj_in[row] + f.finish(acc);
    }
}

template <class F>
__global__ void __launch_bounds__(256) rows_small_kernel(F f, const float* ldj_in, float* ldj_out, int B,
                                                        int G) {
    const int gl = threadIdx.x & (G - 1);
    const long long gid = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) / G;
    const bool live = gid < B;
    const int row = static_cast<int>(gid);
    float acc = 0.f;
    if (live)
        for (int it = gl; it < f.items; it += G) acc += f(row, it);
    for (int o = G >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (live && gl == 0 && ldj_out) ldj_out[row] = ldj_in[row] + f.finish(acc);
}

// rows_warp: one WARP per row, lanes stride over the items, shuffle reduce -- no block barrier at all.  Used when
// the batch alone fills the machine (streaming sizes): resident warps are then limited only by registers.
template <class F>
__global__ void __launch_bounds__(256, 4) rows_warp_kernel(F f, const float* ldj_in, float* ldj_out, int B) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < B; row += gridDim.x * wpb) {
        float acc = 0.f;
        for (int it = lane; it < f.items; it += 32) acc += f(row, it);
        acc = warp_sum(acc);
        if (lane == 0 && ldj_out) ldj_out[row] = ldj_in[row] + f.finish(acc);
    }
}

template <class F>
inline int launch_rows(const F& f, const float* ldj_in, float* ldj_out, int B, cudaStream_t st) {
    if (f.items > 32 && f.items <= 4096 && B >= kSMs * 64) {  // several waves of warps from the batch alone
        const int rows_per_block = 8;
        long long grid = (static_cast<long long>(B) + rows_per_block - 1) / rows_per_block;
        if (grid > kSMs * 64) grid = kSMs * 64;
        rows_warp_kernel<F><<<static_cast<int>(grid), 32 * rows_per_block, 0, st>>>(f, ldj_in, ldj_out, B);
        return launch_status();
    }
    if (f.items > 32) {
        int threads = f.items >= 512 ? 512 : ((f.items + 31) / 32) * 32;
        // a sample needing several passes is better served by 256 threads x more CTAs per SM
        if (f.items > 512) threads = 256;
        const int grid = B < kSMs * 16 ? B : kSMs * 16;
        rows_big_kernel<F><<<grid, threads, 0, st>>>(f, ldj_in, ldj_out, B);
    } else {
        int G = 1;
        while (G < f.items) G <<= 1;
        const long long total = static_cast<long long>(B) * G;
        const int grid = static_cast<int>((total + 255) / 256);
        rows_small_kernel<F><<<grid, 256, 0, st>>>(f, ldj_in, ldj_out, B, G);
    }
    return launch_status();
}

}  // namespace nfb
