// mixlogcdf.cu -- the standalone MixLogCDF module (modules.py:186-212) with the logistic helpers it is made of
// (logistic_logpdf / logistic_logcdf / mix_logistic_logpdf / mix_logistic_logcdf, modules.py:64-97), log domain exactly
// as the reference writes them.  (MixLogAttnCoupling does not go through here: coupling_mixlog.cu fuses the same
// arithmetic with the logit / affine stages of the coupling.)
//   x: (B, n)   log_pi, mu, s: (B, K, n)   y = exp(logsumexp_k(log_pi + logsigmoid(u)))   ldj += sum_n logsumexp_k(log_pi + lpdf)
#include "common.cuh"

namespace nfb {

// -> mixture log-cdf at x; logpdf = mixture log-pdf (modules.py:76-97)
template <bool PDF>
__device__ __forceinline__ float mixcdf_eval(const float* __restrict__ lp, const float* __restrict__ mu,
                                             const float* __restrict__ s, size_t stride, int K, float x, float& logpdf) {
    float cmax = -INFINITY, pmax = -INFINITY;
    for (int k = 0; k < K; ++k) {
        const float sk = __ldg(s + k * stride);
        const float u = __fmul_rn(__fsub_rn(x, __ldg(mu + k * stride)), expf(-sk));  // modules.py:66,72
        cmax = fmaxf(cmax, __ldg(lp + k * stride) + logsigmoid_f(u));
        if (PDF) pmax = fmaxf(pmax, __ldg(lp + k * stride) + (__fsub_rn(u, sk) - 2.f * softplus_f(u)));
    }
    float cs = 0.f, ps = 0.f;
    for (int k = 0; k < K; ++k) {
        const float sk = __ldg(s + k * stride);
        const float u = __fmul_rn(__fsub_rn(x, __ldg(mu + k * stride)), expf(-sk));
        cs += expf(__ldg(lp + k * stride) + logsigmoid_f(u) - cmax);
        if (PDF) ps += expf(__ldg(lp + k * stride) + (__fsub_rn(u, sk) - 2.f * softplus_f(u)) - pmax);
    }
    if (PDF) logpdf = logf(ps) + pmax;
    return logf(cs) + cmax;
}

struct MixCdfFwd {
    const float* x;
    float* y;
    const float* __restrict__ lp;
    const float* __restrict__ mu;
    const float* __restrict__ s;
    int items;  // n
    int K;
    __device__ __forceinline__ float finish(float acc) const { return acc; }
    __device__ __forceinline__ float operator()(int row, int j) const {
        const size_t n = static_cast<size_t>(items), o = static_cast<size_t>(row) * K * n + j;
        float lpdf;
        const float lcdf = mixcdf_eval<true>(lp + o, mu + o, s + o, n, K, x[static_cast<size_t>(row) * n + j], lpdf);
        y[static_cast<size_t>(row) * n + j] = expf(lcdf);  // modules.py:194
        return lpdf;                                       // modules.py:191-192
    }
};

// bisection of modules.py:197-206: phase 0 = 25 steps from [-1e3, 1e3] (+ stall flag: the reference's global break
// test `all(|hi - lo| < 1e-4)` fails only if some element hit val == target), phase 1 = the remaining 75 steps when
// the flag is up
__global__ void __launch_bounds__(256) mixcdf_bisect(const float* __restrict__ y, const float* __restrict__ lp,
                                                    const float* __restrict__ mu, const float* __restrict__ s,
                                                    float* __restrict__ scratch, int* flag, int B, int n, int K, int phase) {
    const long long total = static_cast<long long>(B) * n;
    if (phase == 1 && *reinterpret_cast<volatile int*>(flag) == 0) return;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int row = static_cast<int>(i / n), j = static_cast<int>(i - static_cast<long long>(row) * n);
        const size_t o = static_cast<size_t>(row) * K * n + j;
        const float target = y[i];
        float lo, hi;
        int iters;
        if (phase == 0) { lo = -1.0e3f; hi = 1.0e3f; iters = 25; }
        else { lo = scratch[2 * i]; hi = scratch[2 * i + 1]; iters = 75; }
        for (int it = 0; it < iters; ++it) {
            const float mid = __fmul_rn(__fadd_rn(lo, hi), 0.5f);
            float dummy;
            const float val = expf(mixcdf_eval<false>(lp + o, mu + o, s + o, static_cast<size_t>(n), K, mid, dummy));
            lo = val < target ? mid : lo;
            hi = val > target ? mid : hi;
        }
        scratch[2 * i] = lo;
        scratch[2 * i + 1] = hi;
        if (phase == 0 && !(fabsf(hi - lo) < 1.0e-4f)) atomicOr(flag, 1);
    }
}

struct MixCdfInvFinish {
    float* x;
    const float* __restrict__ lp;
    const float* __restrict__ mu;
    const float* __restrict__ s;
    const float* __restrict__ scratch;
    int items;
    int K;
    __device__ __forceinline__ float finish(float acc) const { return -acc; }  // modules.py:212
    __device__ __forceinline__ float operator()(int row, int j) const {
        const size_t n = static_cast<size_t>(items), i = static_cast<size_t>(row) * n + j, o = static_cast<size_t>(row) * K * n + j;
        const float xv = __fmul_rn(__fadd_rn(scratch[2 * i], scratch[2 * i + 1]), 0.5f);  // modules.py:208
        float lpdf;
        mixcdf_eval<true>(lp + o, mu + o, s + o, n, K, xv, lpdf);
        x[i] = xv;
        return lpdf;
    }
};

}  // namespace nfb

using namespace nfb;

extern "C" int nfb_mixlogcdf_fwd(const float* x, float* y, const float* log_pi, const float* mu, const float* s,
                                 const float* ldj_in, float* ldj_out, int B, int n, int K, nfb_stream_t stream) {
    if (!x || !y || !log_pi || !mu || !s || !ldj_in || !ldj_out) return NFB_ERR_NULL;
    if (B <= 0 || n <= 0 || K <= 0) return NFB_ERR_SHAPE;
    MixCdfFwd f{x, y, log_pi, mu, s, n, K};
    return launch_rows(f, ldj_in, ldj_out, B, as_stream(stream));
}

extern "C" int nfb_mixlogcdf_inv(const float* y, float* x, const float* log_pi, const float* mu, const float* s,
                                 const float* ldj_in, float* ldj_out, float* scratch, int* stall_flag, int B, int n, int K,
                                 nfb_stream_t stream) {
    if (!y || !x || !log_pi || !mu || !s || !ldj_in || !ldj_out || !scratch || !stall_flag) return NFB_ERR_NULL;
    if (B <= 0 || n <= 0 || K <= 0) return NFB_ERR_SHAPE;
    cudaStream_t st = as_stream(stream);
    cudaError_t e = cudaMemsetAsync(stall_flag, 0, sizeof(int), st);
    if (e != cudaSuccess) return static_cast<int>(e);
    const long long total = static_cast<long long>(B) * n;
    long long blocks = (total + 255) / 256;
    if (blocks > kSMs * 32) blocks = kSMs * 32;
    for (int phase = 0; phase < 2; ++phase) {
        mixcdf_bisect<<<static_cast<int>(blocks), 256, 0, st>>>(y, log_pi, mu, s, scratch, stall_flag, B, n, K, phase);
        const int rc = launch_status();
        if (rc != NFB_OK) return rc;
    }
    MixCdfInvFinish f{x, log_pi, mu, s, scratch, n, K};
    return launch_rows(f, ldj_in, ldj_out, B, st);
}
