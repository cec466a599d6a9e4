// permute_nll.cu -- bit-exact index permutations (squeeze.py), the Gaussian NLL reduction (main.py:83-85),
// the WeightNorm fold (weight_norm.py:40) and the library-level glue (version, error strings, launch counter).
#include "common.cuh"

namespace nfb {

unsigned long long g_launches = 0;

// ---- Squeeze2d / Unsqueeze2d -------------------------------------------------------------------------
// thread = 8 consecutive x of one input line (c, y): even x -> squeezed channel k = 4c+2dy, odd x -> k+1.
// With odd != 0 the two halves of the 4C output channels are swapped (squeeze.py:103-105 + cat).
template <bool UNSQ>
__global__ void __launch_bounds__(256) squeeze2d_vec(const float* __restrict__ src, float* __restrict__ dst, int B, int C,
                                                    int H, int W, int odd) {
    const int HW = H * W, h = H >> 1, w = W >> 1;
    const long long nitems = static_cast<long long>(B) * C * HW / 8;
    for (long long it = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; it < nitems;
         it += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long e = it * 8;  // offset in the (B,C,H,W) tensor
        const long long bc = e / HW;
        const int r = static_cast<int>(e - bc * HW);
        const int y = r / W, x0 = r - y * W;
        const int c = static_cast<int>(bc % C);
        const long long b = bc / C;
        int k = 4 * c + 2 * (y & 1);
        int k0 = k, k1 = k + 1;
        if (odd) { k0 = (k0 + 2 * C) % (4 * C); k1 = (k1 + 2 * C) % (4 * C); }
        const long long o0 = ((b * 4 * C + k0) * h + (y >> 1)) * w + (x0 >> 1);
        const long long o1 = ((b * 4 * C + k1) * h + (y >> 1)) * w + (x0 >> 1);
        if (!UNSQ) {
            const float4 v0 = ldg4(src + e), v1 = ldg4(src + e + 4);
            st4(dst + o0, make_float4(v0.x, v0.z, v1.x, v1.z));
            st4(dst + o1, make_float4(v0.y, v0.w, v1.y, v1.w));
        } else {
            const float4 a = ldg4(src + o0), d = ldg4(src + o1);
            st4(dst + e, make_float4(a.x, d.x, a.y, d.y));
            st4(dst + e + 4, make_float4(a.z, d.z, a.w, d.w));
        }
    }
}

template <bool UNSQ>
__global__ void __launch_bounds__(256) squeeze2d_scalar(const float* __restrict__ src, float* __restrict__ dst, int B,
                                                       int C, int H, int W, int odd) {
    const int HW = H * W, h = H >> 1, w = W >> 1;
    const long long total = static_cast<long long>(B) * C * HW;
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long bc = e / HW;
        const int r = static_cast<int>(e - bc * HW);
        const int y = r / W, x = r - y * W;
        const int c = static_cast<int>(bc % C);
        const long long b = bc / C;
        int k = 4 * c + 2 * (y & 1) + (x & 1);
        if (odd) k = (k + 2 * C) % (4 * C);
        const long long o = ((b * 4 * C + k) * h + (y >> 1)) * w + (x >> 1);
        if (!UNSQ) dst[o] = src[e];
        else dst[e] = src[o];
    }
}

template <bool UNSQ>
static int launch_squeeze(const float* src, float* dst, int B, int C, int H, int W, int odd, nfb_stream_t stream) {
    if (!src || !dst) return NFB_ERR_NULL;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return NFB_ERR_SHAPE;
    if ((H % 2) || (W % 2)) return NFB_ERR_SPLIT;
    if (src == dst) return NFB_ERR_UNSUPPORTED;
    const long long total = static_cast<long long>(B) * C * H * W;
    const bool vec = (W % 8 == 0) && aligned16(src) && aligned16(dst);
    const long long work = vec ? total / 8 : total;
    long long blocks = (work + 255) / 256;
    if (blocks > kSMs * 16) blocks = kSMs * 16;
    cudaStream_t st = as_stream(stream);
    if (vec) squeeze2d_vec<UNSQ><<<static_cast<int>(blocks), 256, 0, st>>>(src, dst, B, C, H, W, odd);
    else squeeze2d_scalar<UNSQ><<<static_cast<int>(blocks), 256, 0, st>>>(src, dst, B, C, H, W, odd);
    return launch_status();
}

// ---- coupling split / merge (materialising; the fused coupling kernels do NOT use these) ---------------
template <int MODE, bool MERGE>
__global__ void __launch_bounds__(256) split_merge_kernel(const float* zfull_in, float* zfull_out, const float* h0_in,
                                                         const float* h1_in, float* h0_out, float* h1_out, SplitGeom g) {
    const long long total = static_cast<long long>(g.B) * g.D;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long row = i / g.D;
        const int e = static_cast<int>(i - row * g.D);
        int idx;
        const bool first = classify<MODE>(g, e, idx);
        const long long hoff = row * g.n0 + idx;
        if (!MERGE) {
            const float v = zfull_in[i];
            if (first) { if (h0_out) h0_out[hoff] = v; }
            else       { if (h1_out) h1_out[hoff] = v; }
        } else {
            zfull_out[i] = first ? (h0_in ? h0_in[hoff] : 0.f) : (h1_in ? h1_in[hoff] : 0.f);  // NULL half -> zeros
        }
    }
}

template <bool MERGE>
static int launch_split_merge(const float* zin, float* zout, const float* h0i, const float* h1i, float* h0o, float* h1o,
                              int B, int C, int H, int W, int mode, int odd, nfb_stream_t stream) {
    SplitGeom g;
    const int rc = make_geom(g, B, C, H, W, mode, odd);
    if (rc != NFB_OK) return rc;
    const long long total = static_cast<long long>(B) * g.D;
    long long blocks = (total + 255) / 256;
    if (blocks > kSMs * 16) blocks = kSMs * 16;
    cudaStream_t st = as_stream(stream);
    const int nb = static_cast<int>(blocks);
    if (mode == NFB_SPLIT_1D) split_merge_kernel<NFB_SPLIT_1D, MERGE><<<nb, 256, 0, st>>>(zin, zout, h0i, h1i, h0o, h1o, g);
    else if (mode == NFB_SPLIT_CHECKER) split_merge_kernel<NFB_SPLIT_CHECKER, MERGE><<<nb, 256, 0, st>>>(zin, zout, h0i, h1i, h0o, h1o, g);
    else split_merge_kernel<NFB_SPLIT_CHANNEL, MERGE><<<nb, 256, 0, st>>>(zin, zout, h0i, h1i, h0o, h1o, g);
    return launch_status();
}

// ---- Gaussian NLL ---------------------------------------------------------------------------------------
// one CTA per sample: ||z||^2 in fp64 (D up to 12288 terms; fp32 would lose ~1e-6 relative), then
// nll_b = 0.5||z||^2 + 0.5 D log(2 pi) - ldj_b.  The batch sum is a second, single-CTA fp64 pass: deterministic.
__global__ void __launch_bounds__(256) nll_rows_kernel(const float* __restrict__ z, const float* __restrict__ ldj,
                                                      float* __restrict__ nll_rows, int B, int D) {
    __shared__ double red[33];
    for (int row = blockIdx.x; row < B; row += gridDim.x) {
        const float* zr = z + static_cast<size_t>(row) * D;
        double acc = 0.0;
        if ((D & 3) == 0 && aligned16(zr)) {
            for (int i = threadIdx.x; i < (D >> 2); i += blockDim.x) {
                const float4 v = ldg4(zr + 4 * i);
                acc += static_cast<double>(v.x) * v.x + static_cast<double>(v.y) * v.y +
                       static_cast<double>(v.z) * v.z + static_cast<double>(v.w) * v.w;
            }
        } else {
            for (int i = threadIdx.x; i < D; i += blockDim.x) { const double v = __ldg(zr + i); acc += v * v; }
        }
        acc = block_sum(acc, red);
        if (threadIdx.x == 0) {
            const double nll = 0.5 * acc + 0.5 * static_cast<double>(D) * 1.8378770664093453 /* log(2 pi) */
                               - static_cast<double>(__ldg(ldj + row));
            nll_rows[row] = static_cast<float>(nll);
        }
    }
}

// Streaming batches (the batch alone fills the machine): one WARP per sample, no block barrier; every lane squares its
// float4s in fp32 and adds them pairwise into an fp64 accumulator (one DADD per 4 elements), shuffle-reduced in fp64.
__global__ void __launch_bounds__(256, 4) nll_rows_warp_kernel(const float* __restrict__ z, const float* __restrict__ ldj,
                                                              float* __restrict__ nll_rows, int B, int D) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < B; row += gridDim.x * wpb) {
        const float* zr = z + static_cast<size_t>(row) * D;
        double acc = 0.0;
        const int n4 = D >> 2;
        int i = lane;
        for (; i + 96 < n4; i += 128) {  // four independent 16-byte loads in flight
            const float4 a = ldg4(zr + 4 * i), b = ldg4(zr + 4 * (i + 32)), c = ldg4(zr + 4 * (i + 64)), d = ldg4(zr + 4 * (i + 96));
            acc += static_cast<double>(fmaf(a.x, a.x, a.y * a.y) + fmaf(a.z, a.z, a.w * a.w));
            acc += static_cast<double>(fmaf(b.x, b.x, b.y * b.y) + fmaf(b.z, b.z, b.w * b.w));
            acc += static_cast<double>(fmaf(c.x, c.x, c.y * c.y) + fmaf(c.z, c.z, c.w * c.w));
            acc += static_cast<double>(fmaf(d.x, d.x, d.y * d.y) + fmaf(d.z, d.z, d.w * d.w));
        }
        for (; i < n4; i += 32) {
            const float4 a = ldg4(zr + 4 * i);
            acc += static_cast<double>(fmaf(a.x, a.x, a.y * a.y) + fmaf(a.z, a.z, a.w * a.w));
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            const double nll = 0.5 * acc + 0.5 * static_cast<double>(D) * 1.8378770664093453 - static_cast<double>(__ldg(ldj + row));
            nll_rows[row] = static_cast<float>(nll);
        }
    }
}

__global__ void __launch_bounds__(1024) nll_sum_kernel(const float* __restrict__ rows, double* sum_out, int B) {
    __shared__ double red[33];
    double tot = 0.0;
    int i = threadIdx.x;
    for (; i + 7 * 1024 < B && blockDim.x == 1024; i += 8 * 1024) {  // eight independent loads in flight, same summation order
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = __ldg(rows + i + k * 1024);
#pragma unroll
        for (int k = 0; k < 8; ++k) tot += static_cast<double>(v[k]);
    }
    for (; i < B; i += blockDim.x) tot += static_cast<double>(rows[i]);
    tot = block_sum(tot, red);
    if (threadIdx.x == 0) { sum_out[0] = tot; sum_out[1] = static_cast<double>(B); }
}

// ---- WeightNorm fold ----------------------------------------------------------------------------------
// w[o, j] = v[o, j] * (g[j] / (||v[:, j]||_2 + eps));  norm over dim 0 (the OUTPUT channels), weight_norm.py:21,40
__global__ void __launch_bounds__(256) weight_norm_kernel(const float* __restrict__ v, const float* __restrict__ g,
                                                         float* __restrict__ w, int O, int J, float eps) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= J) return;
    float ss = 0.f;
    for (int o = 0; o < O; ++o) { const float x = __ldg(v + static_cast<size_t>(o) * J + j); ss = fmaf(x, x, ss); }
    const float scale = __fdiv_rn(__ldg(g + j), __fadd_rn(sqrtf(ss), eps));
    for (int o = 0; o < O; ++o) w[static_cast<size_t>(o) * J + j] = __fmul_rn(__ldg(v + static_cast<size_t>(o) * J + j), scale);
}

}  // namespace nfb

using namespace nfb;

extern "C" int nfb_version(void) { return 100; }

extern "C" unsigned long long nfb_launch_count(void) { return g_launches; }

extern "C" const char* nfb_error_string(int code) {
    switch (code) {
        case NFB_OK: return "ok";
        case NFB_ERR_NULL: return "nfb200: required pointer is NULL";
        case NFB_ERR_SHAPE: return "nfb200: non-positive or inconsistent dimension";
        case NFB_ERR_SPLIT: return "nfb200: shape cannot be split in this mode (odd H/W or odd channel count)";
        case NFB_ERR_UNSUPPORTED: return "nfb200: request outside the implemented range";
        default: break;
    }
    if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
    return "nfb200: unknown error code";
}

extern "C" int nfb_squeeze2d(const float* z_in, float* z_out, int B, int C, int H, int W, int odd, nfb_stream_t stream) {
    return launch_squeeze<false>(z_in, z_out, B, C, H, W, odd, stream);
}
extern "C" int nfb_unsqueeze2d(const float* z_in, float* z_out, int B, int C, int H, int W, int odd, nfb_stream_t stream) {
    // z_in is the squeezed (B,4C,H/2,W/2) tensor, z_out the (B,C,H,W) one
    return launch_squeeze<true>(z_in, z_out, B, C, H, W, odd, stream);
}

extern "C" int nfb_coupling_split(const float* z, float* z0_out, float* z1_out, int B, int C, int H, int W, int mode,
                                  int odd, nfb_stream_t stream) {
    if (!z || (!z0_out && !z1_out)) return NFB_ERR_NULL;
    return launch_split_merge<false>(z, nullptr, nullptr, nullptr, z0_out, z1_out, B, C, H, W, mode, odd, stream);
}
extern "C" int nfb_coupling_merge(const float* z0, const float* z1, float* z_out, int B, int C, int H, int W, int mode,
                                  int odd, nfb_stream_t stream) {
    if ((!z0 && !z1) || !z_out) return NFB_ERR_NULL;
    return launch_split_merge<true>(nullptr, z_out, z0, z1, nullptr, nullptr, B, C, H, W, mode, odd, stream);
}

extern "C" int nfb_gauss_nll(const float* z, const float* ldj, float* nll_rows, double* sum_out, int B, int D,
                             nfb_stream_t stream) {
    if (!z || !ldj || !nll_rows) return NFB_ERR_NULL;
    if (B <= 0 || D <= 0) return NFB_ERR_SHAPE;
    cudaStream_t st = as_stream(stream);
    if (B >= kSMs * 64 && D >= 512 && (D & 3) == 0 && aligned16(z)) {
        long long grid = (static_cast<long long>(B) + 7) / 8;
        if (grid > kSMs * 64) grid = kSMs * 64;
        nll_rows_warp_kernel<<<static_cast<int>(grid), 256, 0, st>>>(z, ldj, nll_rows, B, D);
    } else {
        const int grid = B < kSMs * 8 ? B : kSMs * 8;
        nll_rows_kernel<<<grid, D >= 1024 ? 256 : 64, 0, st>>>(z, ldj, nll_rows, B, D);
    }
    const int rc = launch_status();
    if (rc != NFB_OK || !sum_out) return rc;
    nll_sum_kernel<<<1, 1024, 0, st>>>(nll_rows, sum_out, B);  // single CTA, fixed order: deterministic
    return launch_status();
}

extern "C" int nfb_weight_norm(const float* v, const float* g, float* w_out, int O, int Ikk, float eps,
                               nfb_stream_t stream) {
    if (!v || !g || !w_out) return NFB_ERR_NULL;
    if (O <= 0 || Ikk <= 0) return NFB_ERR_SHAPE;
    weight_norm_kernel<<<(Ikk + 255) / 256, 256, 0, as_stream(stream)>>>(v, g, w_out, O, Ikk, eps);
    return launch_status();
}
