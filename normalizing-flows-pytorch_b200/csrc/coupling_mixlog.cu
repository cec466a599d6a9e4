// coupling_mixlog.cu -- Flow++ logistic-mixture-CDF coupling (coupling.py:125-210, modules.py:64-97,186-212).
// forward : one kernel = log_softmax over K + mixture log-pdf/log-cdf (two logsumexps) + Logit(1e-5) +
//           tanh-gated affine + the three log-det terms reduced per sample.
// inverse : affine^-1 -> sigmoid -> bisection on [-1e3, 1e3] entirely in registers (no host sync), two
//           phases to reproduce the reference's global stop rule (25 iterations, or 100 if any element stalls).
//
// Work decomposition: thread = one transformed element, walking z0 in its own (c0,h,w) order so that the
// (2+3K) parameter loads of a warp are each one contiguous 128-byte line.  The kernel is transcendental-bound,
// not HBM-bound (~45 exp/log per element at K = 8), so exp(-|u|) / log1p are shared between softplus(u) and
// logsigmoid(u).
#include "common.cuh"

namespace nfb {

constexpr float kLogitEps = 1.0e-5f;  // Logit() default inside MixLogAttnCoupling (coupling.py:169)

template <int KT>
struct MixParams {
    float logpi[KT ? KT : NFB_MAX_MIXTURES];
    float mu[KT ? KT : NFB_MAX_MIXTURES];
    float s[KT ? KT : NFB_MAX_MIXTURES];
    float inv[KT ? KT : NFB_MAX_MIXTURES];  // exp(-s)
};

// load section values for element j and normalise logpi with log_softmax over k (coupling.py:180)
template <int KT>
__device__ __forceinline__ void load_mix(MixParams<KT>& p, const float* __restrict__ prow, int n0, int j, int K) {
    const int kk = KT ? KT : K;
    const float* base = prow + 2 * static_cast<size_t>(n0) + j;
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < kk; ++k) {
        p.logpi[k] = __ldg(base + static_cast<size_t>(k) * n0);
        p.mu[k] = __ldg(base + static_cast<size_t>(kk + k) * n0);
        p.s[k] = __ldg(base + static_cast<size_t>(2 * kk + k) * n0);
        mx = fmaxf(mx, p.logpi[k]);
    }
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < kk; ++k) sum += expf(p.logpi[k] - mx);
    const float lse = logf(sum);
#pragma unroll
    for (int k = 0; k < kk; ++k) {
        p.logpi[k] = (p.logpi[k] - mx) - lse;
        p.inv[k] = expf(-p.s[k]);
    }
}

// mixture log-cdf (modules.py:88-97) and optionally log-pdf (modules.py:76-85) at x
template <int KT, bool PDF>
__device__ __forceinline__ float mix_eval(const MixParams<KT>& p, float x, int K, float& logpdf) {
    const int kk = KT ? KT : K;
    float cterm[KT ? KT : NFB_MAX_MIXTURES];
    float pterm[KT ? KT : NFB_MAX_MIXTURES];
    float cmax = -INFINITY, pmax = -INFINITY;
#pragma unroll
    for (int k = 0; k < kk; ++k) {
        const float u = __fmul_rn(__fsub_rn(x, p.mu[k]), p.inv[k]);  // modules.py:66,72
        const float l = log1pf(expf(-fabsf(u)));                    // shared by softplus and logsigmoid
        cterm[k] = p.logpi[k] + (fminf(u, 0.f) - l);                // logpi + logsigmoid(u)
        cmax = fmaxf(cmax, cterm[k]);
        if (PDF) {
            const float sp = u > 20.f ? u : fmaxf(u, 0.f) + l;      // softplus(u)
            pterm[k] = p.logpi[k] + ((u - p.s[k]) - 2.f * sp);      // logpi + (u - s - 2 softplus(u))
            pmax = fmaxf(pmax, pterm[k]);
        }
    }
    float cs = 0.f, ps = 0.f;
#pragma unroll
    for (int k = 0; k < kk; ++k) {
        cs += expf(cterm[k] - cmax);
        if (PDF) ps += expf(pterm[k] - pmax);
    }
    if (PDF) logpdf = logf(ps) + pmax;
    return logf(cs) + cmax;
}

// log-domain evaluation exactly as modules.py:76-97,190-194 -- out of line so the fast path keeps its registers
template <int KT>
__device__ __noinline__ float mix_fwd_logdomain(const float* __restrict__ prow, int n0, int j, int K, float x, float& logpdf) {
    MixParams<KT> p;
    load_mix<KT>(p, prow, n0, j, K);
    return expf(mix_eval<KT, true>(p, x, K, logpdf));
}

// ---- forward -------------------------------------------------------------------------------------------
template <int MODE, int KT>
struct MixFwd {
    const float* zin;
    float* zout;
    const float* __restrict__ params;
    const float* __restrict__ pa;
    const float* __restrict__ pb;
    SplitGeom g;
    int items;  // n0
    int K;
    bool inplace;

    __device__ __forceinline__ float finish(float acc) const { return acc; }

    __device__ __forceinline__ float operator()(int row, int j) const {
        const int kk = KT ? KT : K;
        const size_t zbase = static_cast<size_t>(row) * g.D;
        const float* prow = params + static_cast<size_t>(row) * (2 + 3 * kk) * g.n0;
        const int e = half_offset<MODE>(g, j, 0);
        const float x = zin[zbase + e];
        const float a = __fadd_rn(__fmul_rn(tanhf(__ldg(prow + j)), __ldg(pa)), __ldg(pb));  // coupling.py:178
        const float b = __ldg(prow + g.n0 + j);
        // Fast path, linear domain: with e_k = exp(-|u_k|) the logistic cdf is 1/(1+e_k) (u >= 0) or e_k/(1+e_k), and the
        // pdf is exp(-s_k) e_k/(1+e_k)^2; the mixture is a softmax-weighted sum.  2 exp + 1 reciprocal per component
        // instead of 4 exp + 1 log1p.  exp(logsumexp(.)) of modules.py:85,97,194 is the same quantity; the log-domain
        // formulation below takes over when the density underflows (far tails), so extreme inputs behave like the
        // reference.
        float ld1, y;
        {
            const float* base = prow + 2 * static_cast<size_t>(g.n0) + j;
            float lp[KT ? KT : NFB_MAX_MIXTURES], mu[KT ? KT : NFB_MAX_MIXTURES], sv[KT ? KT : NFB_MAX_MIXTURES];
#pragma unroll
            for (int k = 0; k < kk; ++k) {  // all loads first: 3K independent requests in flight
                lp[k] = __ldg(base + static_cast<size_t>(k) * g.n0);
                mu[k] = __ldg(base + static_cast<size_t>(kk + k) * g.n0);
                sv[k] = __ldg(base + static_cast<size_t>(2 * kk + k) * g.n0);
            }
            float mx = lp[0];
#pragma unroll
            for (int k = 1; k < kk; ++k) mx = fmaxf(mx, lp[k]);
            float wsum = 0.f, cdf = 0.f, pdf = 0.f;
#pragma unroll
            for (int k = 0; k < kk; ++k) {
                const float w = expf(lp[k] - mx);
                const float inv = expf(-sv[k]);
                const float u = __fmul_rn(__fsub_rn(x, mu[k]), inv);
                const float ek = expf(-fabsf(u));
                const float r = __fdiv_rn(1.f, 1.f + ek);
                const float c_lo = ek * r;                            // sigma(-|u|); r = sigma(|u|)
                wsum += w;
                cdf = fmaf(w, u >= 0.f ? r : c_lo, cdf);
                pdf = fmaf(w * inv, r * c_lo, pdf);
            }
            const float rw = __fdiv_rn(1.f, wsum);
            y = cdf * rw;
            ld1 = logf(pdf * rw);
            if (!(pdf * rw > 1.0e-30f)) y = mix_fwd_logdomain<KT>(prow, g.n0, j, K, x, ld1);  // far tails: reference formulation
        }
        y = fminf(fmaxf(y, kLogitEps), 1.f - kLogitEps);               // Logit.forward, modules.py:147
        const float l0 = logf(y), l1 = logf(__fsub_rn(1.f, y));        // logit = log y - log(1-y); log-det = -(log y + log(1-y))
        const float lg = __fsub_rn(l0, l1);
        const float ld2 = -__fadd_rn(l0, l1);
        zout[zbase + e] = __fadd_rn(__fmul_rn(lg, expf(a)), b);        // coupling.py:187
        if (!inplace) {
            const int e1 = half_offset<MODE>(g, j, 1);
            zout[zbase + e1] = zin[zbase + e1];
        }
        return (ld1 + ld2) + a;
    }
};

// ---- inverse, phase A: affine^-1, sigmoid, 25 bisection steps -> (lo, hi) in scratch, stall flag ------------
template <int MODE, int KT>
__global__ void __launch_bounds__(256) mix_inv_bisect(const float* __restrict__ zin, const float* __restrict__ params,
                                                     const float* __restrict__ pa, const float* __restrict__ pb,
                                                     float* __restrict__ scratch, int* flag, SplitGeom g, int K,
                                                     int phase) {
    const int kk = KT ? KT : K;
    const long long total = static_cast<long long>(g.B) * g.n0;
    if (phase == 1 && *reinterpret_cast<volatile int*>(flag) == 0) return;  // nobody stalled: reference stopped at 25
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int row = static_cast<int>(i / g.n0);
        const int j = static_cast<int>(i - static_cast<long long>(row) * g.n0);
        const float* prow = params + static_cast<size_t>(row) * (2 + 3 * kk) * g.n0;
        const float a = __fadd_rn(__fmul_rn(tanhf(__ldg(prow + j)), __ldg(pa)), __ldg(pb));
        const float b = __ldg(prow + g.n0 + j);
        const float z = zin[static_cast<size_t>(row) * g.D + half_offset<MODE>(g, j, 0)];
        const float xa = __fmul_rn(expf(-a), __fsub_rn(z, b));       // coupling.py:204
        const float target = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-xa)));  // Logit.backward: sigmoid
        MixParams<KT> p;
        load_mix<KT>(p, prow, g.n0, j, K);
        float lo, hi;
        int iters;
        if (phase == 0) { lo = -1.0e3f; hi = 1.0e3f; iters = 25; }
        else { lo = scratch[2 * i]; hi = scratch[2 * i + 1]; iters = 75; }
        for (int it = 0; it < iters; ++it) {                          // modules.py:199-203
            const float mid = __fmul_rn(__fadd_rn(lo, hi), 0.5f);
            float dummy;
            const float val = expf(mix_eval<KT, false>(p, mid, K, dummy));
            lo = val < target ? mid : lo;
            hi = val > target ? mid : hi;
        }
        scratch[2 * i] = lo;
        scratch[2 * i + 1] = hi;
        if (phase == 0 && !(fabsf(hi - lo) < 1.0e-4f)) atomicOr(flag, 1);  // modules.py:205 would not break
    }
}

// ---- inverse, phase B: x = (lo+hi)/2, log-pdf at x, write, reduce the three log-det terms ---------------------
template <int MODE, int KT>
struct MixInvFinish {
    const float* zin;
    float* zout;
    const float* __restrict__ params;
    const float* __restrict__ pa;
    const float* __restrict__ pb;
    const float* __restrict__ scratch;
    SplitGeom g;
    int items;
    int K;
    bool inplace;

    __device__ __forceinline__ float finish(float acc) const { return acc; }

    __device__ __forceinline__ float operator()(int row, int j) const {
        const int kk = KT ? KT : K;
        const size_t zbase = static_cast<size_t>(row) * g.D;
        const float* prow = params + static_cast<size_t>(row) * (2 + 3 * kk) * g.n0;
        const int e = half_offset<MODE>(g, j, 0);
        const float a = __fadd_rn(__fmul_rn(tanhf(__ldg(prow + j)), __ldg(pa)), __ldg(pb));
        const float b = __ldg(prow + g.n0 + j);
        const float xa = __fmul_rn(expf(-a), __fsub_rn(zin[zbase + e], b));
        const float ld2 = log_dsigmoid_f(xa);                          // Logit.backward, modules.py:153
        MixParams<KT> p;
        load_mix<KT>(p, prow, g.n0, j, K);
        const size_t i = static_cast<size_t>(row) * g.n0 + j;
        const float x = __fmul_rn(__fadd_rn(scratch[2 * i], scratch[2 * i + 1]), 0.5f);  // modules.py:208
        float ld3;
        mix_eval<KT, true>(p, x, K, ld3);
        zout[zbase + e] = x;
        if (!inplace) {
            const int e1 = half_offset<MODE>(g, j, 1);
            zout[zbase + e1] = zin[zbase + e1];
        }
        return (ld2 - a) - ld3;  // coupling.py:205, modules.py:155, modules.py:212
    }
};

template <int MODE, int KT>
static int mix_fwd_launch(const float* zi, float* zo, const float* pr, const float* li, float* lo, const float* a,
                          const float* b, const SplitGeom& g, int K, cudaStream_t st) {
    MixFwd<MODE, KT> f{zi, zo, pr, a, b, g, g.n0, K, zi == zo};
    return launch_rows(f, li, lo, g.B, st);
}

template <int MODE, int KT>
static int mix_inv_launch(const float* zi, float* zo, const float* pr, const float* li, float* lo, const float* a,
                          const float* b, float* scratch, int* flag, const SplitGeom& g, int K, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(flag, 0, sizeof(int), st);
    if (e != cudaSuccess) return static_cast<int>(e);
    const long long total = static_cast<long long>(g.B) * g.n0;
    long long blocks = (total + 255) / 256;
    if (blocks > kSMs * 32) blocks = kSMs * 32;
    for (int phase = 0; phase < 2; ++phase) {
        mix_inv_bisect<MODE, KT><<<static_cast<int>(blocks), 256, 0, st>>>(zi, pr, a, b, scratch, flag, g, K, phase);
        const int rc = launch_status();
        if (rc != NFB_OK) return rc;
    }
    MixInvFinish<MODE, KT> f{zi, zo, pr, a, b, scratch, g, g.n0, K, zi == zo};
    return launch_rows(f, li, lo, g.B, st);
}

#define NFB_DISPATCH_MODE_K(FN, ...)                                                             \
    switch (mode) {                                                                              \
        case NFB_SPLIT_1D:                                                                       \
            return K == 4 ? FN<NFB_SPLIT_1D, 4>(__VA_ARGS__) : K == 8 ? FN<NFB_SPLIT_1D, 8>(__VA_ARGS__) \
                                                                      : FN<NFB_SPLIT_1D, 0>(__VA_ARGS__); \
        case NFB_SPLIT_CHECKER:                                                                  \
            return K == 4 ? FN<NFB_SPLIT_CHECKER, 4>(__VA_ARGS__) : K == 8 ? FN<NFB_SPLIT_CHECKER, 8>(__VA_ARGS__) \
                                                                           : FN<NFB_SPLIT_CHECKER, 0>(__VA_ARGS__); \
        default:                                                                                 \
            return K == 4 ? FN<NFB_SPLIT_CHANNEL, 4>(__VA_ARGS__) : K == 8 ? FN<NFB_SPLIT_CHANNEL, 8>(__VA_ARGS__) \
                                                                           : FN<NFB_SPLIT_CHANNEL, 0>(__VA_ARGS__); \
    }

}  // namespace nfb

using namespace nfb;

extern "C" int nfb_mixlog_coupling_fwd(const float* z_in, float* z_out, const float* params, const float* ldj_in,
                                       float* ldj_out, const float* a_log_scale, const float* a_bias, int B, int C,
                                       int H, int W, int mode, int odd, int K, nfb_stream_t stream) {
    if (!z_in || !z_out || !params || !ldj_in || !ldj_out || !a_log_scale || !a_bias) return NFB_ERR_NULL;
    if (K <= 0) return NFB_ERR_SHAPE;
    if (K > NFB_MAX_MIXTURES) return NFB_ERR_UNSUPPORTED;
    SplitGeom g;
    const int rc = make_geom(g, B, C, H, W, mode, odd);
    if (rc != NFB_OK) return rc;
    cudaStream_t st = as_stream(stream);
    NFB_DISPATCH_MODE_K(mix_fwd_launch, z_in, z_out, params, ldj_in, ldj_out, a_log_scale, a_bias, g, K, st)
}

extern "C" int nfb_mixlog_coupling_inv(const float* z_in, float* z_out, const float* params, const float* ldj_in,
                                       float* ldj_out, const float* a_log_scale, const float* a_bias, float* scratch,
                                       int* stall_flag, int B, int C, int H, int W, int mode, int odd, int K,
                                       nfb_stream_t stream) {
    if (!z_in || !z_out || !params || !ldj_in || !ldj_out || !a_log_scale || !a_bias || !scratch || !stall_flag)
        return NFB_ERR_NULL;
    if (K <= 0) return NFB_ERR_SHAPE;
    if (K > NFB_MAX_MIXTURES) return NFB_ERR_UNSUPPORTED;
    SplitGeom g;
    const int rc = make_geom(g, B, C, H, W, mode, odd);
    if (rc != NFB_OK) return rc;
    cudaStream_t st = as_stream(stream);
    NFB_DISPATCH_MODE_K(mix_inv_launch, z_in, z_out, params, ldj_in, ldj_out, a_log_scale, a_bias, scratch, stall_flag,
                        g, K, st)
}
