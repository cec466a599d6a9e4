// backward.cu -- gradient kernels of the bijective layers: what `loss.backward()` of the reference's
// training step (main.py:78-92) computes through coupling.py:104-112, modules.py:141-150, 246-250, 300-305, 470-482 and
// the NLL of main.py:85, as explicit kernels.  Each layer's backward is one pass that recomputes the cheap forward
// quantities (tanh, exp) from the saved INPUT of the layer, so the forward pass keeps no extra activations.
//
//   upstream: gy  = dLoss/d(layer output z)      (same layout as z)
//             gl  = dLoss/d(ldj_out)             (B,)   -- every layer passes it through unchanged to ldj_in
//   produces: gz  = dLoss/d(layer input z), parameter gradients.
//
// Parameter gradients that reduce over the whole batch are accumulated in fp64 (block partials in fp32, one fp64
// atomicAdd per CTA) into a small scratch buffer and rounded to fp32 once by a 1-CTA finish kernel.
#include "common.cuh"

namespace nfb {

__device__ __forceinline__ double block_sum_d(float v, double* red) { return block_sum(static_cast<double>(v), red); }

// =====================================================================================================================
// AffineCoupling (coupling.py:104-112):  s = tanh(r)*a + b;  y0 = z0*exp(s) + t;  y1 = z1;  ldj += sum s
//   gz0 = gy0*exp(s);  gz1 = gy1;  gt = gy0;  gs = gy0*z0*exp(s) + gl[b];  gr = gs*a*(1 - tanh(r)^2)
//   ga = sum gs*tanh(r);  gb = sum gs
// Thread = 4 consecutive elements of z in its ORIGINAL layout (float4 loads of z / gy, float4 store of gz); the
// (t, r) slots of the transformed elements follow from classify<MODE> (DESIGN.md index formulas).
// =====================================================================================================================
template <int MODE>
__device__ __forceinline__ void affine_bwd_elem(const SplitGeom& g, int e, float z, float gyv, float glv,
                                                const float* __restrict__ pr, float* __restrict__ gpr, float a, float b,
                                                float& gz, float& acc_a, float& acc_b) {
    int idx;
    if (classify<MODE>(g, e, idx)) {
        const float th = tanhf(__ldg(pr + g.n0 + idx));
        const float es = expf(__fadd_rn(__fmul_rn(th, a), b));
        gz = gyv * es;
        const float gs = fmaf(gyv * z, es, glv);
        gpr[idx] = gyv;
        gpr[g.n0 + idx] = gs * a * (1.f - th * th);
        acc_a = fmaf(gs, th, acc_a);
        acc_b += gs;
    } else {
        gz = gyv;
    }
}

template <int MODE, bool VEC>
__global__ void __launch_bounds__(256) affine_bwd_kernel(const float* __restrict__ z, const float* __restrict__ params,
                                                        const float* __restrict__ gy, const float* __restrict__ gl,
                                                        float* __restrict__ gz, float* __restrict__ gparams,
                                                        double* __restrict__ gab, const float* __restrict__ pa,
                                                        const float* __restrict__ pb, SplitGeom g) {
    __shared__ double red[33];
    const float a = __ldg(pa), b = __ldg(pb);
    float acc_a = 0.f, acc_b = 0.f;
    const long long total = static_cast<long long>(g.B) * g.D;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    if (VEC) {
        for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < (total >> 2); v += stride) {
            const long long i = v << 2;
            const long long row = i / g.D;
            const int e = static_cast<int>(i - row * g.D);
            const float glv = gl ? __ldg(gl + row) : 0.f;
            const float* pr = params + row * g.D;
            float* gpr = gparams + row * g.D;
            const float4 zv = ldg4(z + i), gv = ldg4(gy + i);
            float4 o;
            affine_bwd_elem<MODE>(g, e, zv.x, gv.x, glv, pr, gpr, a, b, o.x, acc_a, acc_b);
            affine_bwd_elem<MODE>(g, e + 1, zv.y, gv.y, glv, pr, gpr, a, b, o.y, acc_a, acc_b);
            affine_bwd_elem<MODE>(g, e + 2, zv.z, gv.z, glv, pr, gpr, a, b, o.z, acc_a, acc_b);
            affine_bwd_elem<MODE>(g, e + 3, zv.w, gv.w, glv, pr, gpr, a, b, o.w, acc_a, acc_b);
            st4(gz + i, o);
        }
    } else {
        for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
            const long long row = i / g.D;
            const int e = static_cast<int>(i - row * g.D);
            float o;
            affine_bwd_elem<MODE>(g, e, __ldg(z + i), __ldg(gy + i), gl ? __ldg(gl + row) : 0.f, params + row * g.D,
                                  gparams + row * g.D, a, b, o, acc_a, acc_b);
            gz[i] = o;
        }
    }
    const double sa = block_sum_d(acc_a, red), sb = block_sum_d(acc_b, red);
    if (threadIdx.x == 0) { atomicAdd(gab, sa); atomicAdd(gab + 1, sb); }
}

__global__ void round_to_float_kernel(const double* __restrict__ src, float* dst0, float* dst1) {
    if (threadIdx.x == 0 && dst0) dst0[0] = static_cast<float>(src[0]);
    if (threadIdx.x == 1 && dst1) dst1[0] = static_cast<float>(src[1]);
}

inline int flat_grid(long long work) {
    long long blocks = (work + 255) / 256;
    if (blocks > kSMs * 8) blocks = kSMs * 8;
    if (blocks < 1) blocks = 1;
    return static_cast<int>(blocks);
}

// ---- vectorised variant: the forward kernel's item decomposition (coupling_affine.cu) -- 8 consecutive floats of z in its
// original layout per item, every access a 16-byte load/store, checkerboard classification warp-uniform (no integer
// division per element).  One warp per sample, lanes stride over the items.
struct Bwd4 {
    float4 gz, gr;
};
__device__ __forceinline__ Bwd4 affine_bwd4(float z0, float z1, float z2, float z3, float g0, float g1, float g2, float g3,
                                            const float4& sr, float glv, float a, float b, float& acc_a, float& acc_b) {
    Bwd4 o;
    const float zz[4] = {z0, z1, z2, z3}, gg[4] = {g0, g1, g2, g3}, ss[4] = {sr.x, sr.y, sr.z, sr.w};
    float gzv[4], grv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float th = tanhf(ss[i]);
        const float es = expf(__fadd_rn(__fmul_rn(th, a), b));
        gzv[i] = gg[i] * es;
        const float gs = fmaf(gg[i] * zz[i], es, glv);
        grv[i] = gs * a * (1.f - th * th);
        acc_a = fmaf(gs, th, acc_a);
        acc_b += gs;
    }
    o.gz = make_float4(gzv[0], gzv[1], gzv[2], gzv[3]);
    o.gr = make_float4(grv[0], grv[1], grv[2], grv[3]);
    return o;
}

template <int MODE>
__global__ void __launch_bounds__(256) affine_bwd_vec_kernel(const float* __restrict__ z, const float* __restrict__ params,
                                                            const float* __restrict__ gy, const float* __restrict__ gl,
                                                            float* __restrict__ gz, float* __restrict__ gparams,
                                                            double* __restrict__ gab, const float* __restrict__ pa,
                                                            const float* __restrict__ pb, SplitGeom g, int items,
                                                            int sh_ipl, int sh_wq) {
    __shared__ double red[33];
    const float a = __ldg(pa), b = __ldg(pb);
    float acc_a = 0.f, acc_b = 0.f;
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < g.B; row += gridDim.x * wpb) {
        const size_t base = static_cast<size_t>(row) * g.D;
        const float* zr = z + base;
        const float* gyr = gy + base;
        const float* pr = params + base;
        float* gzr = gz + base;
        float* gpr = gparams + base;
        const float glv = gl ? __ldg(gl + row) : 0.f;
        for (int it = lane; it < items; it += 32) {
            if (MODE == NFB_SPLIT_CHANNEL) {
                const int o0 = (g.odd ? g.n0 : 0) + 8 * it, o1 = (g.odd ? 0 : g.n0) + 8 * it;
                const float4 v0 = ldg4(zr + o0), v1 = ldg4(zr + o0 + 4), q0 = ldg4(gyr + o0), q1 = ldg4(gyr + o0 + 4);
                const float4 s0 = ldg4(pr + g.n0 + 8 * it), s1 = ldg4(pr + g.n0 + 8 * it + 4);
                st4(gzr + o1, ldg4(gyr + o1));
                st4(gzr + o1 + 4, ldg4(gyr + o1 + 4));
                const Bwd4 r0 = affine_bwd4(v0.x, v0.y, v0.z, v0.w, q0.x, q0.y, q0.z, q0.w, s0, glv, a, b, acc_a, acc_b);
                const Bwd4 r1 = affine_bwd4(v1.x, v1.y, v1.z, v1.w, q1.x, q1.y, q1.z, q1.w, s1, glv, a, b, acc_a, acc_b);
                st4(gzr + o0, r0.gz); st4(gzr + o0 + 4, r1.gz);
                st4(gpr + 8 * it, q0); st4(gpr + 8 * it + 4, q1);
                st4(gpr + g.n0 + 8 * it, r0.gr); st4(gpr + g.n0 + 8 * it + 4, r1.gr);
            } else if (MODE == NFB_SPLIT_1D) {
                const float4 v0 = ldg4(zr + 8 * it), v1 = ldg4(zr + 8 * it + 4);
                float4 q0 = ldg4(gyr + 8 * it), q1 = ldg4(gyr + 8 * it + 4);
                const float4 sr = ldg4(pr + g.n0 + 4 * it);
                if (!g.odd) {
                    const Bwd4 r = affine_bwd4(v0.x, v0.z, v1.x, v1.z, q0.x, q0.z, q1.x, q1.z, sr, glv, a, b, acc_a, acc_b);
                    st4(gpr + 4 * it, make_float4(q0.x, q0.z, q1.x, q1.z));
                    q0.x = r.gz.x; q0.z = r.gz.y; q1.x = r.gz.z; q1.z = r.gz.w;
                    st4(gpr + g.n0 + 4 * it, r.gr);
                } else {
                    const Bwd4 r = affine_bwd4(v0.y, v0.w, v1.y, v1.w, q0.y, q0.w, q1.y, q1.w, sr, glv, a, b, acc_a, acc_b);
                    st4(gpr + 4 * it, make_float4(q0.y, q0.w, q1.y, q1.w));
                    q0.y = r.gz.x; q0.w = r.gz.y; q1.y = r.gz.z; q1.w = r.gz.w;
                    st4(gpr + g.n0 + 4 * it, r.gr);
                }
                st4(gzr + 8 * it, q0); st4(gzr + 8 * it + 4, q1);
            } else {
                // items ordered (c, dy, i, jb) as in the forward kernel: all items of one (c, dy) share the two squeezed
                // channels k (even x) and k+1 (odd x)
                const int wq = g.w >> 2, ipl = g.h * wq;
                int cd, rem, i, jb;
                if (sh_ipl >= 0) { cd = it >> sh_ipl; rem = it & (ipl - 1); i = rem >> sh_wq; jb = rem & (wq - 1); }
                else { cd = it / ipl; rem = it - cd * ipl; i = rem / wq; jb = rem - i * wq; }
                const int k = 2 * cd;
                const int e0 = (cd >> 1) * g.HW + (2 * i + (cd & 1)) * g.W + 8 * jb;
                const int qe = (k >= g.C) + (k >= 2 * g.C) + (k >= 3 * g.C);
                const int qo = (k + 1 >= g.C) + (k + 1 >= 2 * g.C) + (k + 1 >= 3 * g.C);
                const bool te = ((qe == 0 || qe == 3) ? 1 : 0) != g.odd, to = ((qo == 0 || qo == 3) ? 1 : 0) != g.odd;
                const int me = (qe == 0) ? k : (qe == 3) ? k - 2 * g.C : k - g.C;
                const int mo = (qo == 0) ? k + 1 : (qo == 3) ? k + 1 - 2 * g.C : k + 1 - g.C;
                const int hw = g.h * g.w, sp = i * g.w + 4 * jb;
                float4 q0 = ldg4(gyr + e0), q1 = ldg4(gyr + e0 + 4);
                float4 v0, v1;
                if (te || to) { v0 = ldg4(zr + e0); v1 = ldg4(zr + e0 + 4); }
                if (te) {
                    const float4 sr = ldg4(pr + g.n0 + me * hw + sp);
                    const Bwd4 r = affine_bwd4(v0.x, v0.z, v1.x, v1.z, q0.x, q0.z, q1.x, q1.z, sr, glv, a, b, acc_a, acc_b);
                    st4(gpr + me * hw + sp, make_float4(q0.x, q0.z, q1.x, q1.z));
                    st4(gpr + g.n0 + me * hw + sp, r.gr);
                    q0.x = r.gz.x; q0.z = r.gz.y; q1.x = r.gz.z; q1.z = r.gz.w;
                }
                if (to) {
                    const float4 sr = ldg4(pr + g.n0 + mo * hw + sp);
                    const Bwd4 r = affine_bwd4(v0.y, v0.w, v1.y, v1.w, q0.y, q0.w, q1.y, q1.w, sr, glv, a, b, acc_a, acc_b);
                    st4(gpr + mo * hw + sp, make_float4(q0.y, q0.w, q1.y, q1.w));
                    st4(gpr + g.n0 + mo * hw + sp, r.gr);
                    q0.y = r.gz.x; q0.w = r.gz.y; q1.y = r.gz.z; q1.w = r.gz.w;
                }
                st4(gzr + e0, q0); st4(gzr + e0 + 4, q1);
            }
        }
    }
    const double sa = block_sum_d(acc_a, red), sb = block_sum_d(acc_b, red);
    if (threadIdx.x == 0) { atomicAdd(gab, sa); atomicAdd(gab + 1, sb); }
}

template <int MODE>
static int launch_affine_bwd(const float* z, const float* params, const float* gy, const float* gl, float* gz,
                             float* gparams, double* gab, const float* a, const float* b, const SplitGeom& g,
                             cudaStream_t st) {
    const long long total = static_cast<long long>(g.B) * g.D;
    const bool al = aligned16(z) && aligned16(gy) && aligned16(gz) && aligned16(params) && aligned16(gparams);
    int items = 0;
    bool vec8 = al;
    if (MODE == NFB_SPLIT_CHANNEL) { vec8 = vec8 && (g.n0 % 8 == 0); items = g.n0 / 8; }
    else if (MODE == NFB_SPLIT_1D) { vec8 = vec8 && (g.D % 8 == 0); items = g.D / 8; }
    else                           { vec8 = vec8 && (g.W % 8 == 0); items = g.D / 8; }
    if (vec8 && items >= 32) {
        int sh_ipl = -1, sh_wq = 0;
        if (MODE == NFB_SPLIT_CHECKER) {
            const int wq = g.w / 4, ipl = g.h * wq;
            auto is_pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
            auto lg = [](int v) { int s = 0; while ((1 << s) < v) ++s; return s; };
            if (is_pow2(wq) && is_pow2(ipl)) { sh_ipl = lg(ipl); sh_wq = lg(wq); }
        }
        long long grid = (static_cast<long long>(g.B) + 7) / 8;
        if (grid > kSMs * 16) grid = kSMs * 16;
        affine_bwd_vec_kernel<MODE><<<static_cast<int>(grid), 256, 0, st>>>(z, params, gy, gl, gz, gparams, gab, a, b, g,
                                                                          items, sh_ipl, sh_wq);
        return launch_status();
    }
    const bool vec = (g.D % 4 == 0) && aligned16(z) && aligned16(gy) && aligned16(gz);
    if (vec) affine_bwd_kernel<MODE, true><<<flat_grid(total / 4), 256, 0, st>>>(z, params, gy, gl, gz, gparams, gab, a, b, g);
    else affine_bwd_kernel<MODE, false><<<flat_grid(total), 256, 0, st>>>(z, params, gy, gl, gz, gparams, gab, a, b, g);
    return launch_status();
}

// =====================================================================================================================
// per-channel affine layers: ActNorm (modules.py:246-250) and the flow BatchNorm (modules.py:300-305; the batch
// statistics are buffers written with .data.copy_, so the reference does NOT differentiate through them).
//   y = (z - m_c) * k_c + beta_c   with  ActNorm: m = bias, k = exp(-log_scale), beta = 0
//                                        BatchNorm: m = mean, k = exp(log_gamma)/sqrt(var)
//   gz = gy * k_c;   S1_c = sum gy,  S2_c = sum gy*z   (over batch and pixels, fp64)
//   ActNorm:   g_bias = -k*S1;  g_log_scale = -k*(S2 - m*S1) - HW * sum_b gl[b]
//   BatchNorm: g_beta = S1;     g_log_gamma =  k*(S2 - m*S1) + HW * sum_b gl[b]
// =====================================================================================================================
enum { CH_ACTNORM = 0, CH_BN = 1 };

template <int KIND>
__device__ __forceinline__ float chan_scale(const float* __restrict__ p0, const float* __restrict__ p1, int c) {
    if (KIND == CH_ACTNORM) return __fdiv_rn(1.f, expf(__ldg(p0 + c)));               // 1/exp(log_scale)
    return __fdiv_rn(expf(__ldg(p0 + c)), sqrtf(__ldg(p1 + c)));                       // exp(log_gamma)/sqrt(var)
}

// flat pass; the two per-channel sums go through shared-memory accumulators (warp-reduced first when the whole warp
// sits in one channel, which is always the case for HW % 128 == 0), then one fp64 atomic per channel per CTA.
template <int KIND, bool VEC>
__global__ void __launch_bounds__(256) chan_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ z,
                                                      const float* __restrict__ p0, const float* __restrict__ p1,
                                                      float* __restrict__ gz, double* __restrict__ sums, int B, int C,
                                                      int HW) {
    extern __shared__ float acc[];  // 2C
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) acc[i] = 0.f;
    __syncthreads();
    const long long total = static_cast<long long>(B) * C * HW;
    const long long n = VEC ? (total >> 2) : total;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const long long n_round = ((n + 31) / 32) * 32;  // keep whole warps in the loop for the shuffles
    for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < n_round; v += stride) {
        const bool live = v < n;
        int c = -1;
        float s1 = 0.f, s2 = 0.f;
        if (live) {
            const long long i = VEC ? (v << 2) : v;
            c = static_cast<int>((i / HW) % C);
            const float k = chan_scale<KIND>(p0, p1, c);
            if (VEC) {
                const float4 g4 = ldg4(gy + i), z4 = ldg4(z + i);
                st4(gz + i, make_float4(g4.x * k, g4.y * k, g4.z * k, g4.w * k));
                s1 = (g4.x + g4.y) + (g4.z + g4.w);
                s2 = fmaf(g4.x, z4.x, fmaf(g4.y, z4.y, fmaf(g4.z, z4.z, g4.w * z4.w)));
            } else {
                const float gv = __ldg(gy + i);
                gz[i] = gv * k;
                s1 = gv;
                s2 = gv * __ldg(z + i);
            }
        }
        const int c0 = __shfl_sync(0xffffffffu, c, 0);
        if (__all_sync(0xffffffffu, c == c0)) {
            s1 = warp_sum(s1);
            s2 = warp_sum(s2);
            if ((threadIdx.x & 31) == 0 && c0 >= 0) { atomicAdd(acc + c0, s1); atomicAdd(acc + C + c0, s2); }
        } else if (live) {
            atomicAdd(acc + c, s1);
            atomicAdd(acc + C + c, s2);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x)
        if (acc[i] != 0.f) atomicAdd(sums + i, static_cast<double>(acc[i]));
}

template <int KIND>
__global__ void __launch_bounds__(256) chan_bwd_finish_kernel(const double* __restrict__ sums, const float* __restrict__ gl,
                                                             const float* __restrict__ p0, const float* __restrict__ p1,
                                                             const float* __restrict__ pm, float* __restrict__ g_scale,
                                                             float* __restrict__ g_shift, int B, int C, int HW) {
    __shared__ double red[33];
    double t = 0.0;
    if (gl)
        for (int b = threadIdx.x; b < B; b += blockDim.x) t += static_cast<double>(__ldg(gl + b));
    const double gl_sum = block_sum(t, red) * static_cast<double>(HW);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const double k = static_cast<double>(chan_scale<KIND>(p0, p1, c));
        const double m = static_cast<double>(__ldg(pm + c));
        const double s1 = sums[c], s2 = sums[C + c];
        if (KIND == CH_ACTNORM) {
            if (g_scale) g_scale[c] = static_cast<float>(-k * (s2 - m * s1) - gl_sum);
            if (g_shift) g_shift[c] = static_cast<float>(-k * s1);
        } else {
            if (g_scale) g_scale[c] = static_cast<float>(k * (s2 - m * s1) + gl_sum);
            if (g_shift) g_shift[c] = static_cast<float>(s1);
        }
    }
}

template <int KIND>
static int launch_chan_bwd(const float* gy, const float* z, const float* gl, const float* p0, const float* p1,
                           const float* pm, float* gz, float* g_scale, float* g_shift, double* scratch, int B, int C,
                           int HW, nfb_stream_t stream) {
    if (!gy || !z || !p0 || !pm || !gz || !scratch) return NFB_ERR_NULL;
    if (KIND == CH_BN && !p1) return NFB_ERR_NULL;
    if (B <= 0 || C <= 0 || HW <= 0) return NFB_ERR_SHAPE;
    if (C > 4096) return NFB_ERR_UNSUPPORTED;
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * C, st);
    const long long total = static_cast<long long>(B) * C * HW;
    const bool vec = (HW % 4 == 0) && aligned16(gy) && aligned16(z) && aligned16(gz);
    const size_t smem = sizeof(float) * 2 * C;
    if (vec) chan_bwd_kernel<KIND, true><<<flat_grid(total / 4), 256, smem, st>>>(gy, z, p0, p1, gz, scratch, B, C, HW);
    else chan_bwd_kernel<KIND, false><<<flat_grid(total), 256, smem, st>>>(gy, z, p0, p1, gz, scratch, B, C, HW);
    int rc = launch_status();
    if (rc != NFB_OK || (!g_scale && !g_shift)) return rc;
    chan_bwd_finish_kernel<KIND><<<1, 256, 0, st>>>(scratch, gl, p0, p1, pm, g_scale, g_shift, B, C, HW);
    return launch_status();
}

// =====================================================================================================================
// InvertibleConv1x1 (modules.py:470-482): y[b,:,p] = W z[b,:,p]; ldj += sum(log_s)*HW
//   gz = W^T gy (nfb_invconv1x1_apply with the transposed matrix);  gW[i,j] = sum_{b,p} gy[b,i,p] z[b,j,p]  (here)
//   W = P A U', A = L o tril(-1) + I, U' = U o triu(1) + diag(d), d = sign_s exp(log_s):
//   M = P^T gW;  gL = (M U'^T) o tril(-1);  gU = (A^T M) o triu(1);  g_log_s = diag(A^T M) d + HW sum_b gl[b]
// =====================================================================================================================
// gW = GY (C x N) . Z^T (N x C), N = B*HW positions.  CTA = one T x T output tile x one chunk of positions; a thread
// owns a 4 x 4 block of the tile, so (T/4)^2 threads cover it and the 256 threads form G = 256/(T/4)^2 groups that walk
// disjoint positions of the staged chunk (T = 16: 16 groups -- the small channel counts C = 3, 12 still use every
// thread); the groups are summed through shared memory, every CTA stores its partial tile with plain stores and
// invconv_wgrad_reduce adds the chunks: no atomics, deterministic.
template <int T>
__global__ void __launch_bounds__(256) invconv_wgrad_kernel(const float* __restrict__ gy, const float* __restrict__ z,
                                                           float* __restrict__ partial, int B, int C, int HW, int chunk) {
    constexpr int TQ = T / 4;              // threads per tile side
    constexpr int TG = TQ * TQ;            // threads per group
    constexpr int G = 256 / TG;            // position groups
    constexpr int TP = T == 64 ? 64 : 128; // positions staged per step
    constexpr int LD = T + 4;              // padded row (positions-major tiles: [p][channel])
    __shared__ __align__(16) float sm[2 * TP * LD];
    float* gs = sm;
    float* zs = sm + TP * LD;
    const int nt = (C + T - 1) / T;
    const int ti = blockIdx.y / nt, tj = blockIdx.y - ti * nt;
    const int i0 = ti * T, j0 = tj * T;
    const int grp = threadIdx.x / TG, tl = threadIdx.x - grp * TG;
    const int ti4 = (tl / TQ) * 4, tj4 = (tl % TQ) * 4;
    const int npos = B * HW;
    const int q0 = blockIdx.x * chunk;
    const int q1 = q0 + chunk < npos ? q0 + chunk : npos;
    float acc[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = 0.f;
    for (int q = q0; q < q1; q += TP) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < T * TP; idx += 256) {  // consecutive threads -> consecutive positions
            const int ch = idx / TP, p = idx - ch * TP;
            const int pos = q + p;
            float a = 0.f, b = 0.f;
            if (pos < q1) {
                const int bb = pos / HW, pp = pos - bb * HW;
                if (i0 + ch < C) a = __ldg(gy + (static_cast<size_t>(bb) * C + i0 + ch) * HW + pp);
                if (j0 + ch < C) b = __ldg(z + (static_cast<size_t>(bb) * C + j0 + ch) * HW + pp);
            }
            gs[p * LD + ch] = a;
            zs[p * LD + ch] = b;
        }
        __syncthreads();
#pragma unroll 4
        for (int p = grp; p < TP; p += G) {
            const float4 a = ld4(gs + p * LD + ti4), b = ld4(zs + p * LD + tj4);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(av[u], bv[v], acc[u][v]);
        }
    }
    // sum the G position groups through shared memory (reusing the staging tiles), then one store per tile entry
    __syncthreads();
    float* red = sm;  // G * T * T = 4096 floats, both staging tiles together hold at least that
    static_assert(G * T * T <= 2 * TP * LD, "reduction buffer");
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) red[grp * T * T + (ti4 + u) * T + tj4 + v] = acc[u][v];
    __syncthreads();
    float* pw = partial + static_cast<size_t>(blockIdx.x) * C * C;
    for (int e = threadIdx.x; e < T * T; e += 256) {
        float r = 0.f;
#pragma unroll
        for (int k = 0; k < G; ++k) r += red[k * T * T + e];
        const int i = i0 + e / T, j = j0 + e % T;
        if (i < C && j < C) pw[i * C + j] = r;
    }
}

__global__ void __launch_bounds__(256) invconv_wgrad_reduce(const float* __restrict__ partial, float* __restrict__ gW, int n,
                                                           int chunks) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float a0 = 0.f, a1 = 0.f;
    int k = 0;
    for (; k + 2 <= chunks; k += 2) {
        a0 += __ldg(partial + static_cast<size_t>(k) * n + i);
        a1 += __ldg(partial + static_cast<size_t>(k + 1) * n + i);
    }
    if (k < chunks) a0 += __ldg(partial + static_cast<size_t>(k) * n + i);
    gW[i] = a0 + a1;
}

// one thread per (i, j): strictly-lower entries -> gL, strictly-upper -> gU, the diagonal -> g_log_s
__global__ void __launch_bounds__(256) invconv_weight_bwd_kernel(const float* __restrict__ gW, const float* __restrict__ P,
                                                                const float* __restrict__ L, const float* __restrict__ U,
                                                                const float* __restrict__ log_s,
                                                                const float* __restrict__ sign_s,
                                                                const float* __restrict__ gl, float* __restrict__ gL,
                                                                float* __restrict__ gU, float* __restrict__ g_log_s, int B,
                                                                int C, int HW) {
    extern __shared__ int perm[];  // perm[k] = r with P[r,k] = 1, i.e. (P^T gW)[k,:] = gW[perm[k],:]
    __shared__ double red[33];
    for (int k = threadIdx.x; k < C; k += blockDim.x) {
        int r = 0;
        for (int rr = 0; rr < C; ++rr)
            if (__ldg(P + rr * C + k) != 0.f) r = rr;
        perm[k] = r;
    }
    double t = 0.0;
    if (gl)  // every CTA needs it: the diagonal is spread over all of them
        for (int b = threadIdx.x; b < B; b += blockDim.x) t += static_cast<double>(__ldg(gl + b));
    const double gl_sum = block_sum(t, red) * static_cast<double>(HW);  // also the barrier that publishes perm[]
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= C * C) return;
    const int i = idx / C, j = idx - i * C;
    if (i > j) {
        // gL[i,j] = sum_c M[i,c] U'[j,c],  U'[j,c] != 0 only for c >= j
        const float* Mi = gW + perm[i] * C;
        float acc = __ldg(Mi + j) * __fmul_rn(__ldg(sign_s + j), expf(__ldg(log_s + j)));
        for (int c = j + 1; c < C; ++c) acc = fmaf(__ldg(Mi + c), __ldg(U + j * C + c), acc);
        gL[idx] = acc;
        gU[idx] = 0.f;
    } else {
        // (A^T M)[i,j] = sum_k A[k,i] M[k,j],  A[k,i] != 0 only for k >= i (A[i,i] = 1)
        float acc = __ldg(gW + perm[i] * C + j);
        for (int k = i + 1; k < C; ++k) acc = fmaf(__ldg(L + k * C + i), __ldg(gW + perm[k] * C + j), acc);
        if (i == j) {
            const float d = __fmul_rn(__ldg(sign_s + i), expf(__ldg(log_s + i)));
            g_log_s[i] = static_cast<float>(static_cast<double>(acc * d) + gl_sum);
            gL[idx] = 0.f;
            gU[idx] = 0.f;
        } else {
            gU[idx] = acc;
            gL[idx] = 0.f;
        }
    }
}

// =====================================================================================================================
// Logit (modules.py:146-150): x' = clamp(x, lo, hi); y = log x' - log(1-x'); ldj += sum -(log x' + log(1-x'))
//   inside the clamp range: gx = gy/(x(1-x)) + gl[b]*(1/(1-x) - 1/x);  outside: 0 (clamp has zero gradient)
// NLL (main.py:85): nll_b = 0.5||z_b||^2 + const - ldj_b:  gz = z * g_b,  g_ldj = -g_b
// =====================================================================================================================
__global__ void __launch_bounds__(256) logit_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gy,
                                                       const float* __restrict__ gl, float* __restrict__ gx, float lo,
                                                       float hi, int B, int D) {
    const long long total = static_cast<long long>(B) * D;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float xv = __ldg(x + i);
        float o = 0.f;
        if (xv >= lo && xv <= hi) {
            const float rx = __fdiv_rn(1.f, xv), r1 = __fdiv_rn(1.f, __fsub_rn(1.f, xv));
            o = __ldg(gy + i) * rx * r1;
            if (gl) o = fmaf(__ldg(gl + i / D), r1 - rx, o);
        }
        gx[i] = o;
    }
}

__global__ void __launch_bounds__(256) nll_bwd_kernel(const float* __restrict__ z, const float* __restrict__ grows,
                                                     float* __restrict__ gz, float* __restrict__ gldj, int B, int D) {
    const long long total = static_cast<long long>(B) * D;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        gz[i] = __ldg(z + i) * __ldg(grows + i / D);
    if (gldj)
        for (long long b = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; b < B;
             b += static_cast<long long>(gridDim.x) * blockDim.x)
            gldj[b] = -__ldg(grows + b);
}


// =====================================================================================================================
// MixLogAttnCoupling (coupling.py:172-190): per transformed element x with parameters (ar, b, lp_k, mu_k, s_k):
//   a = tanh(ar) A + Bb;  pi = softmax(lp);  u_k = (x - mu_k) e^{-s_k};  sg_k = sigmoid(u_k)
//   F = sum pi_k sg_k;  log f = logsumexp_k(log pi_k + u_k - s_k - 2 softplus(u_k));  y = clamp(F, eps, 1-eps)
//   out = logit(y) e^a + b;  ldj += log f - log y - log(1-y) + a
// Backward (G_* = dLoss/d*, gl = dLoss/dldj of the row), with r_k = softmax_k of the log-pdf terms (responsibilities,
// computed in the log domain so that far tails cannot divide by an underflowed density):
//   G_b = gy;  G_a = gy logit(y) e^a + gl;  G_y = gy e^a/(y(1-y)) + gl (1/(1-y) - 1/y);  G_F = G_y inside the clamp
//   G_u_k = G_F pi_k sg_k (1 - sg_k) + gl r_k (1 - 2 sg_k)
//   G_x = sum G_u_k e^{-s_k};  G_mu_k = -G_u_k e^{-s_k};  G_s_k = -G_u_k u_k - gl r_k
//   G_logpi_k = G_F pi_k sg_k + gl r_k;  G_lp_k = G_logpi_k - pi_k (G_F F + gl)      (log_softmax backward)
//   G_ar = G_a A (1 - tanh^2 ar);  G_A = sum G_a tanh(ar);  G_Bb = sum G_a
// =====================================================================================================================
constexpr float kLogitEpsBwd = 1.0e-5f;  // Logit() default inside MixLogAttnCoupling (coupling.py:169)

template <int MODE, int KT>
__global__ void __launch_bounds__(128) mixlog_bwd_kernel(const float* __restrict__ z, const float* __restrict__ params,
                                                        const float* __restrict__ gy, const float* __restrict__ gl,
                                                        float* __restrict__ gz, float* __restrict__ gparams,
                                                        double* __restrict__ gab, const float* __restrict__ pa,
                                                        const float* __restrict__ pb, SplitGeom g, int K) {
    __shared__ double red[33];
    const int kk = KT ? KT : K;
    const float A = __ldg(pa), Bb = __ldg(pb);
    float acc_a = 0.f, acc_b = 0.f;
    const long long total = static_cast<long long>(g.B) * g.n0;
    const size_t n0 = g.n0;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int row = static_cast<int>(i / g.n0);
        const int j = static_cast<int>(i - static_cast<long long>(row) * g.n0);
        const size_t zbase = static_cast<size_t>(row) * g.D;
        const size_t pbase = static_cast<size_t>(row) * (2 + 3 * kk) * n0 + j;
        const float* prow = params + pbase;
        float* gprow = gparams + pbase;
        const int e = half_offset<MODE>(g, j, 0), e1 = half_offset<MODE>(g, j, 1);
        gz[zbase + e1] = __ldg(gy + zbase + e1);  // pass-through half
        const float x = __ldg(z + zbase + e), gyv = __ldg(gy + zbase + e), glv = gl ? __ldg(gl + row) : 0.f;
        const float th = tanhf(__ldg(prow));
        const float a = __fadd_rn(__fmul_rn(th, A), Bb);
        const float ea = expf(a);
        float lp[KT ? KT : NFB_MAX_MIXTURES], mu[KT ? KT : NFB_MAX_MIXTURES], sv[KT ? KT : NFB_MAX_MIXTURES];
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < kk; ++k) {
            lp[k] = __ldg(prow + (2 + k) * n0);
            mu[k] = __ldg(prow + (2 + kk + k) * n0);
            sv[k] = __ldg(prow + (2 + 2 * kk + k) * n0);
            mx = fmaxf(mx, lp[k]);
        }
        float wsum = 0.f;
#pragma unroll
        for (int k = 0; k < kk; ++k) wsum += expf(lp[k] - mx);
        const float lse = mx + logf(wsum);
        // pass 1: F and the log-pdf terms
        float pt[KT ? KT : NFB_MAX_MIXTURES];
        float F = 0.f, pmax = -INFINITY;
#pragma unroll
        for (int k = 0; k < kk; ++k) {
            const float lpi = lp[k] - lse;
            const float inv = expf(-sv[k]);
            const float u = __fmul_rn(__fsub_rn(x, mu[k]), inv);
            const float ek = expf(-fabsf(u));
            const float r = __fdiv_rn(1.f, 1.f + ek);
            const float sg = u >= 0.f ? r : ek * r;
            F = fmaf(expf(lpi), sg, F);
            const float sp = fmaxf(u, 0.f) + log1pf(ek);  // softplus(u)
            pt[k] = lpi + ((u - sv[k]) - 2.f * sp);
            pmax = fmaxf(pmax, pt[k]);
        }
        float psum = 0.f;
#pragma unroll
        for (int k = 0; k < kk; ++k) psum += expf(pt[k] - pmax);
        const float rps = __fdiv_rn(1.f, psum);
        const float y = fminf(fmaxf(F, kLogitEpsBwd), 1.f - kLogitEpsBwd);
        const float ry = __fdiv_rn(1.f, y), r1 = __fdiv_rn(1.f, __fsub_rn(1.f, y));
        const float lg = logf(y) - logf(__fsub_rn(1.f, y));
        const float Ga = fmaf(gyv * lg, ea, glv);
        const float Gy = fmaf(gyv * ea, ry * r1, glv * (r1 - ry));
        const float GF = (F >= kLogitEpsBwd && F <= 1.f - kLogitEpsBwd) ? Gy : 0.f;
        const float lsm_corr = fmaf(GF, F, glv);  // sum_k G_logpi_k
        // pass 2: per-component gradients
        float Gx = 0.f;
#pragma unroll
        for (int k = 0; k < kk; ++k) {
            const float pik = expf(lp[k] - lse);
            const float inv = expf(-sv[k]);
            const float u = __fmul_rn(__fsub_rn(x, mu[k]), inv);
            const float ek = expf(-fabsf(u));
            const float r = __fdiv_rn(1.f, 1.f + ek);
            const float sg = u >= 0.f ? r : ek * r;
            const float rk = expf(pt[k] - pmax) * rps;
            const float Gu = fmaf(GF * pik, sg * (1.f - sg), glv * rk * (1.f - 2.f * sg));
            Gx = fmaf(Gu, inv, Gx);
            const float Glpi = fmaf(GF * pik, sg, glv * rk);
            gprow[(2 + k) * n0] = Glpi - pik * lsm_corr;
            gprow[(2 + kk + k) * n0] = -Gu * inv;
            gprow[(2 + 2 * kk + k) * n0] = -fmaf(Gu, u, glv * rk);
        }
        gz[zbase + e] = Gx;
        gprow[0] = Ga * A * (1.f - th * th);
        gprow[n0] = gyv;
        acc_a = fmaf(Ga, th, acc_a);
        acc_b += Ga;
    }
    const double sa = block_sum_d(acc_a, red), sb = block_sum_d(acc_b, red);
    if (threadIdx.x == 0) { atomicAdd(gab, sa); atomicAdd(gab + 1, sb); }
}

template <int MODE, int KT>
static int mixlog_bwd_launch(const float* z, const float* params, const float* gy, const float* gl, float* gz,
                             float* gparams, double* gab, const float* a, const float* b, const SplitGeom& g, int K,
                             cudaStream_t st) {
    const long long total = static_cast<long long>(g.B) * g.n0;
    long long blocks = (total + 127) / 128;
    if (blocks > kSMs * 16) blocks = kSMs * 16;
    mixlog_bwd_kernel<MODE, KT><<<static_cast<int>(blocks), 128, 0, st>>>(z, params, gy, gl, gz, gparams, gab, a, b, g, K);
    return launch_status();
}

// =====================================================================================================================
// Rational-quadratic spline coupling (oracle/flow_oracle.py:rqs_transform, forward direction).  Per element:
//   xk, yk = knots(softmax widths / heights), bin, (x0,x1,y0,y1,d0,d1), w, h, sk = h/w, xi = (x-x0)/w, om = xi(1-xi),
//   t = d1+d0-2sk, den = sk + t om, num = sk xi^2 + d0 om, out = y0 + h num/den,
//   ld = 2 log sk + log(d1 xi^2 + 2 sk om + d0 (1-xi)^2) - 2 log den.
// The backward walks that expression tree in reverse; only the two knots of the active bin and its two derivatives
// receive gradient, which is then spread over the softmax logits (cumsum + softmax backward).
// =====================================================================================================================
constexpr float kMinWb = 1.0e-3f, kMinHb = 1.0e-3f, kMinDb = 1.0e-3f;

// softmax probabilities sm[0..K) and knots[0..K] exactly as rqs_knots (coupling_rqs.cu)
template <int KT>
__device__ __forceinline__ void rqs_knots_sm(float* knots, float* sm, const float* __restrict__ base, size_t stride, int K,
                                             float mn, float bound) {
    const int kk = KT ? KT : K;
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < kk; ++k) { sm[k] = __ldg(base + k * stride); mx = fmaxf(mx, sm[k]); }
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < kk; ++k) { sm[k] = expf(sm[k] - mx); sum += sm[k]; }
    const float scale = 1.f - mn * static_cast<float>(kk);
    float c = 0.f;
    knots[0] = -bound;
#pragma unroll
    for (int k = 0; k < kk; ++k) {
        sm[k] = __fdiv_rn(sm[k], sum);
        c = __fadd_rn(c, __fadd_rn(mn, __fmul_rn(scale, sm[k])));
        knots[k + 1] = __fsub_rn(__fmul_rn(2.f * bound, c), bound);
    }
    knots[kk] = bound;
}

// gradient wrt the K logits given the gradients of the two active knots (lower knot index `bin`, upper `bin`+1)
template <int KT>
__device__ __forceinline__ void rqs_logit_grads(float* __restrict__ gout, size_t stride, const float* sm, int K, int bin,
                                                float g_lo, float g_hi, float mn, float bound) {
    const int kk = KT ? KT : K;
    const float scale = (1.f - mn * static_cast<float>(kk)) * 2.f * bound;
    if (bin == 0) g_lo = 0.f;            // knot 0 is the constant -bound
    if (bin + 1 == kk) g_hi = 0.f;       // knot K is the constant +bound
    // G_sm_i = scale * (g_lo [i < bin] + g_hi [i < bin+1]);  G_u_i = sm_i (G_sm_i - sum_j sm_j G_sm_j)
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < kk; ++k) {
        const float gs = scale * ((k < bin ? g_lo : 0.f) + (k < bin + 1 ? g_hi : 0.f));
        dot = fmaf(sm[k], gs, dot);
    }
#pragma unroll
    for (int k = 0; k < kk; ++k) {
        const float gs = scale * ((k < bin ? g_lo : 0.f) + (k < bin + 1 ? g_hi : 0.f));
        gout[k * stride] = sm[k] * (gs - dot);
    }
}

template <int MODE, int KT>
__global__ void __launch_bounds__(128) rqs_bwd_kernel(const float* __restrict__ z, const float* __restrict__ params,
                                                     const float* __restrict__ gy, const float* __restrict__ gl,
                                                     float* __restrict__ gz, float* __restrict__ gparams, SplitGeom g,
                                                     int K, float bound) {
    const int kk = KT ? KT : K;
    const long long total = static_cast<long long>(g.B) * g.n0;
    const size_t n0 = g.n0;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int row = static_cast<int>(i / g.n0);
        const int j = static_cast<int>(i - static_cast<long long>(row) * g.n0);
        const size_t zbase = static_cast<size_t>(row) * g.D;
        const size_t pbase = static_cast<size_t>(row) * (3 * kk - 1) * n0 + j;
        const float* prow = params + pbase;
        float* gprow = gparams + pbase;
        const int e = half_offset<MODE>(g, j, 0), e1 = half_offset<MODE>(g, j, 1);
        gz[zbase + e1] = __ldg(gy + zbase + e1);
        const float x = __ldg(z + zbase + e), Gout = __ldg(gy + zbase + e), Gld = gl ? __ldg(gl + row) : 0.f;
        if (!(x >= -bound && x <= bound)) {  // identity tails
            gz[zbase + e] = Gout;
            for (int k = 0; k < 3 * kk - 1; ++k) gprow[k * n0] = 0.f;
            continue;
        }
        float xk[(KT ? KT : NFB_MAX_BINS) + 1], yk[(KT ? KT : NFB_MAX_BINS) + 1];
        float smw[KT ? KT : NFB_MAX_BINS], smh[KT ? KT : NFB_MAX_BINS];
        rqs_knots_sm<KT>(xk, smw, prow, n0, K, kMinWb, bound);
        rqs_knots_sm<KT>(yk, smh, prow + static_cast<size_t>(kk) * n0, n0, K, kMinHb, bound);
        int bin = 0;
#pragma unroll
        for (int k = 1; k < kk; ++k) bin += (x >= xk[k]) ? 1 : 0;
        float x0 = 0.f, x1 = 0.f, y0 = 0.f, y1 = 0.f;
#pragma unroll
        for (int k = 0; k < kk; ++k)
            if (k == bin) { x0 = xk[k]; x1 = xk[k + 1]; y0 = yk[k]; y1 = yk[k + 1]; }
        const float* dbase = prow + static_cast<size_t>(2 * kk) * n0;
        const float ud0 = bin == 0 ? 0.f : __ldg(dbase + static_cast<size_t>(bin - 1) * n0);
        const float ud1 = bin == kk - 1 ? 0.f : __ldg(dbase + static_cast<size_t>(bin) * n0);
        const float d0 = bin == 0 ? 1.f : kMinDb + softplus_f(ud0);
        const float d1 = bin == kk - 1 ? 1.f : kMinDb + softplus_f(ud1);
        const float w = x1 - x0, h = y1 - y0;
        const float rw = __fdiv_rn(1.f, w);
        const float sk = h * rw;
        const float xi = (x - x0) * rw;
        const float om = xi * (1.f - xi);
        const float t = d1 + d0 - 2.f * sk;
        const float den = sk + t * om;
        const float num = sk * xi * xi + d0 * om;
        const float q = d1 * xi * xi + 2.f * sk * om + d0 * (1.f - xi) * (1.f - xi);
        const float rden = __fdiv_rn(1.f, den);
        // reverse sweep
        float Gy0 = Gout;
        float Gh = Gout * num * rden;
        const float Gnum = Gout * h * rden;
        float Gden = -Gout * h * num * rden * rden - 2.f * Gld * rden;
        float Gsk = 2.f * Gld / sk;
        const float Gq = Gld / q;
        float Gd1 = Gq * xi * xi;
        float Gd0 = Gq * (1.f - xi) * (1.f - xi);
        float Gxi = Gq * (2.f * d1 * xi - 2.f * d0 * (1.f - xi));
        float Gom = Gq * 2.f * sk;
        Gsk += Gq * 2.f * om;
        Gsk += Gnum * xi * xi;
        Gxi += Gnum * 2.f * sk * xi;
        Gd0 += Gnum * om;
        Gom += Gnum * d0;
        Gsk += Gden;
        const float Gt = Gden * om;
        Gom += Gden * t;
        Gd1 += Gt;
        Gd0 += Gt;
        Gsk -= 2.f * Gt;
        Gxi += Gom * (1.f - 2.f * xi);
        const float Gx = Gxi * rw;
        float Gx0 = -Gxi * rw;
        float Gw = -Gxi * xi * rw;
        Gh += Gsk * rw;
        Gw -= Gsk * sk * rw;
        const float Gx1 = Gw;
        Gx0 -= Gw;
        const float Gy1 = Gh;
        Gy0 -= Gh;
        gz[zbase + e] = Gx;
        rqs_logit_grads<KT>(gprow, n0, smw, K, bin, Gx0, Gx1, kMinWb, bound);
        rqs_logit_grads<KT>(gprow + static_cast<size_t>(kk) * n0, n0, smh, K, bin, Gy0, Gy1, kMinHb, bound);
        float* gd = gprow + static_cast<size_t>(2 * kk) * n0;
        for (int k = 0; k < kk - 1; ++k) {
            float v = 0.f;
            // d softplus(u)/du = sigmoid(u) (1 beyond the threshold-20 linear branch, where sigmoid rounds to 1 anyway)
            if (k == bin - 1) v = Gd0 * __fdiv_rn(1.f, 1.f + expf(-ud0));
            if (k == bin) v = Gd1 * __fdiv_rn(1.f, 1.f + expf(-ud1));
            gd[k * n0] = v;
        }
    }
}

template <int MODE, int KT>
static int rqs_bwd_launch(const float* z, const float* params, const float* gy, const float* gl, float* gz, float* gparams,
                          const SplitGeom& g, int K, float bound, cudaStream_t st) {
    const long long total = static_cast<long long>(g.B) * g.n0;
    long long blocks = (total + 127) / 128;
    if (blocks > kSMs * 16) blocks = kSMs * 16;
    rqs_bwd_kernel<MODE, KT><<<static_cast<int>(blocks), 128, 0, st>>>(z, params, gy, gl, gz, gparams, g, K, bound);
    return launch_status();
}

}  // namespace nfb

using namespace nfb;

extern "C" int nfb_affine_coupling_bwd(const float* z_in, const float* params, const float* gy, const float* gldj,
                                       float* gz, float* gparams, float* g_s_log_scale, float* g_s_bias, double* scratch,
                                       const float* s_log_scale, const float* s_bias, int B, int C, int H, int W,
                                       int mode, int odd, nfb_stream_t stream) {
    if (!z_in || !params || !gy || !gz || !gparams || !scratch || !s_log_scale || !s_bias) return NFB_ERR_NULL;
    SplitGeom g;
    int rc = make_geom(g, B, C, H, W, mode, odd);
    if (rc != NFB_OK) return rc;
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(scratch, 0, 2 * sizeof(double), st);
    if (mode == NFB_SPLIT_1D) rc = launch_affine_bwd<NFB_SPLIT_1D>(z_in, params, gy, gldj, gz, gparams, scratch, s_log_scale, s_bias, g, st);
    else if (mode == NFB_SPLIT_CHECKER) rc = launch_affine_bwd<NFB_SPLIT_CHECKER>(z_in, params, gy, gldj, gz, gparams, scratch, s_log_scale, s_bias, g, st);
    else rc = launch_affine_bwd<NFB_SPLIT_CHANNEL>(z_in, params, gy, gldj, gz, gparams, scratch, s_log_scale, s_bias, g, st);
    if (rc != NFB_OK || (!g_s_log_scale && !g_s_bias)) return rc;
    round_to_float_kernel<<<1, 32, 0, st>>>(scratch, g_s_log_scale, g_s_bias);
    return launch_status();
}

extern "C" int nfb_actnorm_bwd(const float* gy, const float* z_in, const float* gldj, const float* log_scale,
                               const float* bias, float* gz, float* g_log_scale, float* g_bias, double* scratch, int B,
                               int C, int HW, nfb_stream_t stream) {
    return launch_chan_bwd<CH_ACTNORM>(gy, z_in, gldj, log_scale, nullptr, bias, gz, g_log_scale, g_bias, scratch, B, C,
                                       HW, stream);
}

extern "C" int nfb_bnflow_bwd(const float* gy, const float* x_in, const float* gldj, const float* mean, const float* var,
                              const float* log_gamma, float* gx, float* g_log_gamma, float* g_beta, double* scratch,
                              int B, int C, int HW, nfb_stream_t stream) {
    return launch_chan_bwd<CH_BN>(gy, x_in, gldj, log_gamma, var, mean, gx, g_log_gamma, g_beta, scratch, B, C, HW,
                                  stream);
}

static void invconv_wgrad_plan(int B, int C, int HW, int& T, int& chunk, int& chunks) {
    T = C <= 16 ? 16 : (C <= 32 ? 32 : 64);
    const int nt = (C + T - 1) / T;
    const long long npos = static_cast<long long>(B) * HW;
    long long want = (kSMs * 2 + nt * nt - 1) / (nt * nt);   // about two CTAs per SM in total
    long long ck = (npos + want - 1) / want;
    ck = ((ck + 127) / 128) * 128;                           // whole staging steps
    if (ck < 128) ck = 128;
    chunk = static_cast<int>(ck);
    chunks = static_cast<int>((npos + ck - 1) / ck);
}

extern "C" long long nfb_invconv1x1_wgrad_scratch(int B, int C, int HW) {
    if (B <= 0 || C <= 0 || HW <= 0) return NFB_ERR_SHAPE;
    int T, chunk, chunks;
    invconv_wgrad_plan(B, C, HW, T, chunk, chunks);
    return static_cast<long long>(chunks) * C * C;
}

extern "C" int nfb_invconv1x1_wgrad(const float* gy, const float* z_in, float* gW, float* scratch, int B, int C, int HW,
                                    nfb_stream_t stream) {
    if (!gy || !z_in || !gW || !scratch) return NFB_ERR_NULL;
    if (B <= 0 || C <= 0 || HW <= 0) return NFB_ERR_SHAPE;
    if (static_cast<long long>(B) * HW > (1LL << 30)) return NFB_ERR_SHAPE;
    cudaStream_t st = as_stream(stream);
    int T, chunk, chunks;
    invconv_wgrad_plan(B, C, HW, T, chunk, chunks);
    const int nt = (C + T - 1) / T;
    dim3 grid(chunks, nt * nt);
    if (T == 16) invconv_wgrad_kernel<16><<<grid, 256, 0, st>>>(gy, z_in, scratch, B, C, HW, chunk);
    else if (T == 32) invconv_wgrad_kernel<32><<<grid, 256, 0, st>>>(gy, z_in, scratch, B, C, HW, chunk);
    else invconv_wgrad_kernel<64><<<grid, 256, 0, st>>>(gy, z_in, scratch, B, C, HW, chunk);
    const int rc = launch_status();
    if (rc != NFB_OK) return rc;
    invconv_wgrad_reduce<<<(C * C + 255) / 256, 256, 0, st>>>(scratch, gW, C * C, chunks);
    return launch_status();
}

extern "C" int nfb_invconv1x1_weight_bwd(const float* gW, const float* P, const float* L, const float* U,
                                         const float* log_s, const float* sign_s, const float* gldj, float* gL, float* gU,
                                         float* g_log_s, int B, int C, int HW, nfb_stream_t stream) {
    if (!gW || !P || !L || !U || !log_s || !sign_s || !gL || !gU || !g_log_s) return NFB_ERR_NULL;
    if (B <= 0 || C <= 0 || HW <= 0) return NFB_ERR_SHAPE;
    if (C > 4096) return NFB_ERR_UNSUPPORTED;
    invconv_weight_bwd_kernel<<<(C * C + 255) / 256, 256, sizeof(int) * C, as_stream(stream)>>>(
        gW, P, L, U, log_s, sign_s, gldj, gL, gU, g_log_s, B, C, HW);
    return launch_status();
}

extern "C" int nfb_logit_bwd(const float* x_in, const float* gy, const float* gldj, float* gx, float lo, float hi, int B,
                             int D, nfb_stream_t stream) {
    if (!x_in || !gy || !gx) return NFB_ERR_NULL;
    if (B <= 0 || D <= 0) return NFB_ERR_SHAPE;
    logit_bwd_kernel<<<flat_grid(static_cast<long long>(B) * D), 256, 0, as_stream(stream)>>>(x_in, gy, gldj, gx, lo, hi, B, D);
    return launch_status();
}

extern "C" int nfb_gauss_nll_bwd(const float* z, const float* g_rows, float* gz, float* gldj, int B, int D,
                                 nfb_stream_t stream) {
    if (!z || !g_rows || !gz) return NFB_ERR_NULL;
    if (B <= 0 || D <= 0) return NFB_ERR_SHAPE;
    nll_bwd_kernel<<<flat_grid(static_cast<long long>(B) * D), 256, 0, as_stream(stream)>>>(z, g_rows, gz, gldj, B, D);
    return launch_status();
}

extern "C" int nfb_mixlog_coupling_bwd(const float* z_in, const float* params, const float* gy, const float* gldj,
                                       float* gz, float* gparams, float* g_a_log_scale, float* g_a_bias, double* scratch,
                                       const float* a_log_scale, const float* a_bias, int B, int C, int H, int W, int mode,
                                       int odd, int K, nfb_stream_t stream) {
    if (!z_in || !params || !gy || !gz || !gparams || !scratch || !a_log_scale || !a_bias) return NFB_ERR_NULL;
    if (K <= 0) return NFB_ERR_SHAPE;
    if (K > NFB_MAX_MIXTURES) return NFB_ERR_UNSUPPORTED;
    SplitGeom g;
    int rc = make_geom(g, B, C, H, W, mode, odd);
    if (rc != NFB_OK) return rc;
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(scratch, 0, 2 * sizeof(double), st);
#define NFB_MIXBWD(M)                                                                                                  \
    (K == 4 ? mixlog_bwd_launch<M, 4>(z_in, params, gy, gldj, gz, gparams, scratch, a_log_scale, a_bias, g, K, st)     \
            : K == 8 ? mixlog_bwd_launch<M, 8>(z_in, params, gy, gldj, gz, gparams, scratch, a_log_scale, a_bias, g, K, st) \
                     : mixlog_bwd_launch<M, 0>(z_in, params, gy, gldj, gz, gparams, scratch, a_log_scale, a_bias, g, K, st))
    if (mode == NFB_SPLIT_1D) rc = NFB_MIXBWD(NFB_SPLIT_1D);
    else if (mode == NFB_SPLIT_CHECKER) rc = NFB_MIXBWD(NFB_SPLIT_CHECKER);
    else rc = NFB_MIXBWD(NFB_SPLIT_CHANNEL);
#undef NFB_MIXBWD
    if (rc != NFB_OK || (!g_a_log_scale && !g_a_bias)) return rc;
    round_to_float_kernel<<<1, 32, 0, st>>>(scratch, g_a_log_scale, g_a_bias);
    return launch_status();
}

extern "C" int nfb_rqs_coupling_bwd(const float* z_in, const float* params, const float* gy, const float* gldj, float* gz,
                                    float* gparams, int B, int C, int H, int W, int mode, int odd, int K, float bound,
                                    nfb_stream_t stream) {
    if (!z_in || !params || !gy || !gz || !gparams) return NFB_ERR_NULL;
    if (K < 2 || !(bound > 0.f)) return NFB_ERR_SHAPE;
    if (K > NFB_MAX_BINS) return NFB_ERR_UNSUPPORTED;
    SplitGeom g;
    const int rc = make_geom(g, B, C, H, W, mode, odd);
    if (rc != NFB_OK) return rc;
    cudaStream_t st = as_stream(stream);
#define NFB_RQSBWD(M)                                                                              \
    (K == 8 ? rqs_bwd_launch<M, 8>(z_in, params, gy, gldj, gz, gparams, g, K, bound, st)           \
            : rqs_bwd_launch<M, 0>(z_in, params, gy, gldj, gz, gparams, g, K, bound, st))
    if (mode == NFB_SPLIT_1D) return NFB_RQSBWD(NFB_SPLIT_1D);
    if (mode == NFB_SPLIT_CHECKER) return NFB_RQSBWD(NFB_SPLIT_CHECKER);
    return NFB_RQSBWD(NFB_SPLIT_CHANNEL);
#undef NFB_RQSBWD
}
