// channel_layers.cu -- per-channel affine bijections: ActNorm and the flow BatchNorm (modules.py:225-322),
// their cross-sample statistics (ActNorm first-call init, BatchNorm train mode), and Logit (modules.py:141-155).
//
// ActNorm / BatchNorm: flat float4 elementwise pass; the log-det of these layers is the same scalar for every
// sample (sum over channels x pixels), so the leading CTAs also apply it to ldj -- one launch per layer.
#include "common.cuh"

namespace nfb {

enum { OP_ACTNORM_FWD = 0, OP_ACTNORM_INV = 1, OP_BN_FWD = 2, OP_BN_INV = 3 };

struct ChanArgs {
    const float* p0;  // actnorm: log_scale   | bn: mean
    const float* p1;  // actnorm: bias        | bn: var
    const float* p2;  //                      | bn: log_gamma
    const float* p3;  //                      | bn: beta
};

template <int OP>
__device__ __forceinline__ void chan_coeff(const ChanArgs& a, int c, float& u, float& v, float& w, float& x) {
    if (OP == OP_ACTNORM_FWD || OP == OP_ACTNORM_INV) {
        u = expf(__ldg(a.p0 + c));  // exp(log_scale)
        v = __ldg(a.p1 + c);        // bias
        w = x = 0.f;
    } else {
        u = __ldg(a.p0 + c);               // mean
        v = sqrtf(__ldg(a.p1 + c));        // sqrt(var)
        w = expf(__ldg(a.p2 + c));         // exp(log_gamma)
        x = __ldg(a.p3 + c);               // beta
    }
}

template <int OP>
__device__ __forceinline__ float chan_apply(float z, float u, float v, float w, float x) {
    if (OP == OP_ACTNORM_FWD) return __fdiv_rn(__fsub_rn(z, v), u);                    // modules.py:246
    if (OP == OP_ACTNORM_INV) return __fadd_rn(__fmul_rn(z, u), v);                    // modules.py:253
    if (OP == OP_BN_FWD) return __fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(z, u), v), w), x);  // modules.py:300-301
    return __fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(z, x), w), v), u);                  // modules.py:315-316
}

// per-channel term of the log-det (before the * HW)
template <int OP>
__device__ __forceinline__ float chan_logdet_term(const ChanArgs& a, int c) {
    if (OP == OP_ACTNORM_FWD) return -__ldg(a.p0 + c);
    if (OP == OP_ACTNORM_INV) return __ldg(a.p0 + c);
    const float t = __fsub_rn(__ldg(a.p2 + c), __fmul_rn(0.5f, logf(__ldg(a.p1 + c))));  // modules.py:304
    return OP == OP_BN_FWD ? t : -t;
}

template <int OP, bool VEC>
__global__ void __launch_bounds__(256) chan_kernel(const float* zin, float* zout, const float* ldj_in,
                                                  float* ldj_out, ChanArgs a, int B, int C, int HW) {
    // (1) log-det: the first ceil(B/256) CTAs own one sample per thread
    if (static_cast<long long>(blockIdx.x) * blockDim.x < B) {
        float part = 0.f;
        for (int c = threadIdx.x & 31; c < C; c += 32) part += chan_logdet_term<OP>(a, c);
        part = warp_sum(part);  // torch.sum(log_scale) -- fp32, order differs by a few ulp at most
        const int b = blockIdx.x * blockDim.x + threadIdx.x;
        if (b < B) ldj_out[b] = __fadd_rn(ldj_in[b], __fmul_rn(part, static_cast<float>(HW)));
    }
    // (2) the elementwise map
    const long long total = static_cast<long long>(B) * C * HW;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    if (VEC) {  // HW % 4 == 0: a float4 never straddles a channel
        const long long nv = total >> 2;
        for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nv; i += stride) {
            const int c = static_cast<int>(((i << 2) / HW) % C);
            float u, v, w, x;
            chan_coeff<OP>(a, c, u, v, w, x);
            float4 z = ld4(zin + (i << 2));
            z.x = chan_apply<OP>(z.x, u, v, w, x);
            z.y = chan_apply<OP>(z.y, u, v, w, x);
            z.z = chan_apply<OP>(z.z, u, v, w, x);
            z.w = chan_apply<OP>(z.w, u, v, w, x);
            st4(zout + (i << 2), z);
        }
    } else {
        for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
            const int c = static_cast<int>((i / HW) % C);
            float u, v, w, x;
            chan_coeff<OP>(a, c, u, v, w, x);
            zout[i] = chan_apply<OP>(zin[i], u, v, w, x);
        }
    }
}

template <int OP>
static int launch_chan(const float* zin, float* zout, const float* ldj_in, float* ldj_out, ChanArgs a, int B, int C,
                       int HW, nfb_stream_t stream) {
    if (!zin || !zout || !ldj_in || !ldj_out || !a.p0 || !a.p1) return NFB_ERR_NULL;
    if ((OP == OP_BN_FWD || OP == OP_BN_INV) && (!a.p2 || !a.p3)) return NFB_ERR_NULL;
    if (B <= 0 || C <= 0 || HW <= 0) return NFB_ERR_SHAPE;
    const long long total = static_cast<long long>(B) * C * HW;
    const bool vec = (HW % 4 == 0) && aligned16(zin) && aligned16(zout);
    const long long work = vec ? total / 4 : total;
    long long blocks = (work + 255) / 256;
    const long long need = (B + 255) / 256;  // CTAs that own the ldj update
    if (blocks > kSMs * 16) blocks = kSMs * 16;
    if (blocks < need) blocks = need;
    cudaStream_t st = as_stream(stream);
    if (vec) chan_kernel<OP, true><<<static_cast<int>(blocks), 256, 0, st>>>(zin, zout, ldj_in, ldj_out, a, B, C, HW);
    else chan_kernel<OP, false><<<static_cast<int>(blocks), 256, 0, st>>>(zin, zout, ldj_in, ldj_out, a, B, C, HW);
    return launch_status();
}

// ---- cross-sample statistics: one CTA per channel, fp64 accumulation of (sum, sum of squares) ---------
// mode 0: ActNorm init  -> out0 = log(std_unbiased + eps), out1 = mean      (modules.py:238-244)
// mode 1: BatchNorm     -> out0 = mean,                   out1 = biased var + eps (modules.py:285-287)
__global__ void __launch_bounds__(512) chan_stats_kernel(const float* __restrict__ z, float* out0, float* out1, int B,
                                                        int C, int HW, float eps, int mode) {
    __shared__ double red[33];
    const int c = blockIdx.x;
    const long long n = static_cast<long long>(B) * HW;
    // pass 1: mean
    double s = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const long long b = i / HW, p = i - b * HW;
        s += static_cast<double>(__ldg(z + (b * C + c) * HW + p));
    }
    const double mean = block_sum(s, red) / static_cast<double>(n);
    // pass 2: centred second moment (two-pass: no cancellation)
    double q = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const long long b = i / HW, p = i - b * HW;
        const double d = static_cast<double>(__ldg(z + (b * C + c) * HW + p)) - mean;
        q += d * d;
    }
    q = block_sum(q, red);
    if (threadIdx.x == 0) {
        if (mode == 0) {
            const double var = q / static_cast<double>(n > 1 ? n - 1 : 1);
            out0[c] = logf(__fadd_rn(static_cast<float>(sqrt(var)), eps));
            out1[c] = static_cast<float>(mean);
        } else {
            out0[c] = static_cast<float>(mean);
            out1[c] = __fadd_rn(static_cast<float>(q / static_cast<double>(n)), eps);
        }
    }
}

// per-channel raw moments in fp64: out[c] = sum x, out[C + c] = sum x^2 over (batch, pixels).  Additive across ranks:
// all-reduce (SUM) the 2C doubles and the element count, then finalise (parallel.py) -> statistics identical to a
// single-process pass over the global batch (SURVEY.md 8e exceptions: ActNorm init, BatchNorm train mode).
__global__ void __launch_bounds__(512) chan_moments_kernel(const float* __restrict__ z, double* __restrict__ out, int B, int C,
                                                          int HW) {
    __shared__ double red[33];
    const int c = blockIdx.x;
    const long long n = static_cast<long long>(B) * HW;
    double s = 0.0, q = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const long long b = i / HW, p = i - b * HW;
        const double v = static_cast<double>(__ldg(z + (b * C + c) * HW + p));
        s += v;
        q += v * v;
    }
    s = block_sum(s, red);
    q = block_sum(q, red);
    if (threadIdx.x == 0) { out[c] = s; out[C + c] = q; }
}

// ---- Logit ------------------------------------------------------------------------------------------
template <bool INV, bool VEC>
struct LogitRow {
    const float* zin;
    float* zout;
    float lo, hi;
    int D;
    int items;

    __device__ __forceinline__ float finish(float acc) const { return acc; }

    __device__ __forceinline__ float one(float& x) const {
        if (!INV) {
            x = fminf(fmaxf(x, lo), hi);                     // modules.py:147
            // logit(x) = log x - log(1-x) and -(y - 2 softplus(y)) = -(log x + log(1-x)) exactly (softplus(logit x) =
            // -log(1-x)); two logarithms instead of div + logf + expf + log1pf (modules.py:29-32,148-150).
            // The logarithms are lg2.approx * ln 2 (__logf): with libm's logf the layer was issue-bound at 0.56 of the HBM
            // roofline (~50 instructions per element for 8 bytes).  Documented error of __logf: <= 2^-21.41 absolute for
            // arguments in [0.5, 2], <= 3 ulp elsewhere.  Both results here are SUMS of the two logarithms, one of which is
            // always >= ln 2 in magnitude, so an absolute error of 3.6e-7 per term is the rounding level of the results
            // themselves (ulp(0.69) = 6e-8, ulp(4.6) = 4.8e-7); worst case over a 3072-pixel sample 1.1e-3 nats on a
            // log-likelihood of ~8e3 nats = 1.4e-7 relative, inside the 1e-5 bits/dim budget.
            const float l0 = __logf(x), l1 = __logf(__fsub_rn(1.f, x));
            x = __fsub_rn(l0, l1);
            return -__fadd_rn(l0, l1);
        }
        const float ld = log_dsigmoid_f(x);                   // modules.py:153
        x = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-x)));         // torch.sigmoid
        return ld;
    }

    __device__ __forceinline__ float operator()(int row, int it) const {
        const size_t base = static_cast<size_t>(row) * D;
        if (VEC) {
            float4 v = ld4(zin + base + 4 * it);
            float acc = one(v.x);
            acc += one(v.y);
            acc += one(v.z);
            acc += one(v.w);
            st4(zout + base + 4 * it, v);
            return acc;
        }
        float x = zin[base + it];
        const float ld = one(x);
        zout[base + it] = x;
        return ld;
    }
};

template <bool INV>
static int launch_logit(const float* zin, float* zout, const float* ldj_in, float* ldj_out, float lo, float hi, int B,
                        int D, nfb_stream_t stream) {
    if (!zin || !zout || !ldj_in || !ldj_out) return NFB_ERR_NULL;
    if (B <= 0 || D <= 0) return NFB_ERR_SHAPE;
    cudaStream_t st = as_stream(stream);
    if (D % 4 == 0 && aligned16(zin) && aligned16(zout)) {
        LogitRow<INV, true> f{zin, zout, lo, hi, D, D / 4};
        return launch_rows(f, ldj_in, ldj_out, B, st);
    }
    LogitRow<INV, false> f{zin, zout, lo, hi, D, D};
    return launch_rows(f, ldj_in, ldj_out, B, st);
}

}  // namespace nfb

using namespace nfb;

extern "C" int nfb_actnorm_fwd(const float* z_in, float* z_out, const float* ldj_in, float* ldj_out,
                               const float* log_scale, const float* bias, int B, int C, int HW, nfb_stream_t stream) {
    return launch_chan<OP_ACTNORM_FWD>(z_in, z_out, ldj_in, ldj_out, ChanArgs{log_scale, bias, nullptr, nullptr}, B, C, HW, stream);
}
extern "C" int nfb_actnorm_inv(const float* z_in, float* z_out, const float* ldj_in, float* ldj_out,
                               const float* log_scale, const float* bias, int B, int C, int HW, nfb_stream_t stream) {
    return launch_chan<OP_ACTNORM_INV>(z_in, z_out, ldj_in, ldj_out, ChanArgs{log_scale, bias, nullptr, nullptr}, B, C, HW, stream);
}
extern "C" int nfb_bnflow_fwd(const float* z_in, float* z_out, const float* ldj_in, float* ldj_out, const float* mean,
                              const float* var, const float* log_gamma, const float* beta, int B, int C, int HW,
                              nfb_stream_t stream) {
    return launch_chan<OP_BN_FWD>(z_in, z_out, ldj_in, ldj_out, ChanArgs{mean, var, log_gamma, beta}, B, C, HW, stream);
}
extern "C" int nfb_bnflow_inv(const float* z_in, float* z_out, const float* ldj_in, float* ldj_out, const float* mean,
                              const float* var, const float* log_gamma, const float* beta, int B, int C, int HW,
                              nfb_stream_t stream) {
    return launch_chan<OP_BN_INV>(z_in, z_out, ldj_in, ldj_out, ChanArgs{mean, var, log_gamma, beta}, B, C, HW, stream);
}

static int launch_stats(const float* z, float* o0, float* o1, int B, int C, int HW, float eps, int mode,
                        nfb_stream_t stream) {
    if (!z || !o0 || !o1) return NFB_ERR_NULL;
    if (B <= 0 || C <= 0 || HW <= 0) return NFB_ERR_SHAPE;
    chan_stats_kernel<<<C, 512, 0, as_stream(stream)>>>(z, o0, o1, B, C, HW, eps, mode);
    return launch_status();
}
extern "C" int nfb_actnorm_init(const float* z, float* log_scale_out, float* bias_out, int B, int C, int HW, float eps,
                                nfb_stream_t stream) {
    return launch_stats(z, log_scale_out, bias_out, B, C, HW, eps, 0, stream);
}
extern "C" int nfb_bnflow_batch_stats(const float* z, float* mean_out, float* var_out, int B, int C, int HW, float eps,
                                      nfb_stream_t stream) {
    return launch_stats(z, mean_out, var_out, B, C, HW, eps, 1, stream);
}

extern "C" int nfb_logit_fwd(const float* z_in, float* z_out, const float* ldj_in, float* ldj_out, float lo, float hi,
                             int B, int D, nfb_stream_t stream) {
    return launch_logit<false>(z_in, z_out, ldj_in, ldj_out, lo, hi, B, D, stream);
}
extern "C" int nfb_logit_inv(const float* z_in, float* z_out, const float* ldj_in, float* ldj_out, int B, int D,
                             nfb_stream_t stream) {
    return launch_logit<true>(z_in, z_out, ldj_in, ldj_out, 0.f, 1.f, B, D, stream);
}

extern "C" int nfb_channel_moments(const float* z, double* moments_out, int B, int C, int HW, nfb_stream_t stream) {
    if (!z || !moments_out) return NFB_ERR_NULL;
    if (B <= 0 || C <= 0 || HW <= 0) return NFB_ERR_SHAPE;
    chan_moments_kernel<<<C, 512, 0, as_stream(stream)>>>(z, moments_out, B, C, HW);
    return launch_status();
}
