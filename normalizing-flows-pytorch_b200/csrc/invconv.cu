// invconv.cu -- InvertibleConv1x1 (modules.py:441-497).
//   nfb_invconv1x1_weight : W = P (L o tril + I)(U o triu + diag(sign_s exp(log_s)))  and (optionally) W^-1
//   nfb_invconv1x1_apply  : out[b,:,p] = M z[b,:,p]  (+ the sample-independent log-det sum(log_s) * HW)
// The contraction is tiny (C in {2,3,12,48,192}) and needs fp32-exact products (bits/dim parity 1e-5 rules out
// single-pass TF32/BF16 tensor-core math), so it is register-tiled FP32 FFMA with the C x C matrix in shared memory.
#include "common.cuh"

namespace nfb {

// ---- weight assembly: one CTA; masks and the diagonal are applied on the fly (no staging) ----------------
constexpr int kMaxInvconvC = 256;

__device__ __forceinline__ float lower_at(const float* __restrict__ L, int C, int r, int c) {  // L o tril(-1) + I
    return c < r ? __ldg(L + r * C + c) : (c == r ? 1.f : 0.f);
}
__device__ __forceinline__ float upper_at(const float* __restrict__ U, const float* diag, int C, int r, int c) {
    return c > r ? __ldg(U + r * C + c) : (c == r ? diag[r] : 0.f);  // U o triu(1) + diag(sign_s exp(log_s))
}

__global__ void __launch_bounds__(256) invconv_weight_kernel(const float* __restrict__ P, const float* __restrict__ L,
                                                            const float* __restrict__ U, const float* __restrict__ log_s,
                                                            const float* __restrict__ sign_s, float* __restrict__ W,
                                                            float* __restrict__ Winv, int C) {
    __shared__ float diag[kMaxInvconvC];
    for (int i = threadIdx.x; i < C; i += blockDim.x) diag[i] = __fmul_rn(__ldg(sign_s + i), expf(__ldg(log_s + i)));
    __syncthreads();
    // W[r,c] = sum_k P[r,k] (L' U')[k,c]   (modules.py:473).  P is a permutation, so only one k contributes.
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < C * C; i += gridDim.x * blockDim.x) {  // all CTAs share W
        const int r = i / C, c = i - r * C;
        float acc = 0.f;
        for (int k = 0; k < C; ++k) {
            const float p = __ldg(P + r * C + k);
            if (p != 0.f) {
                float t = 0.f;
                const int jmax = k < c ? k : c;
                for (int j = 0; j <= jmax; ++j) t = fmaf(lower_at(L, C, k, j), upper_at(U, diag, C, j, c), t);
                acc = fmaf(p, t, acc);
            }
        }
        W[i] = acc;
    }
    if (Winv == nullptr || blockIdx.x != 0) return;  // the inverse (fp64 substitution, thread per column) stays on CTA 0
    // W^-1 = U'^-1 L'^-1 P^T.  Thread j solves W x = e_j: forward then backward substitution, all in fp64,
    // rounded to fp32 once.  Replaces the per-pixel lu_solve of modules.py:490 by one matrix apply.
    for (int j = threadIdx.x; j < C; j += blockDim.x) {
        double y[kMaxInvconvC];
        for (int i = 0; i < C; ++i) {
            double acc = static_cast<double>(__ldg(P + j * C + i));  // (P^T e_j)_i = P[j,i]
            for (int k = 0; k < i; ++k) acc -= static_cast<double>(__ldg(L + i * C + k)) * y[k];
            y[i] = acc;
        }
        for (int i = C - 1; i >= 0; --i) {
            double acc = y[i];
            for (int k = i + 1; k < C; ++k) acc -= static_cast<double>(__ldg(U + i * C + k)) * y[k];
            y[i] = acc / static_cast<double>(diag[i]);
        }
        for (int i = 0; i < C; ++i) Winv[i * C + j] = static_cast<float>(y[i]);
    }
}

// ---- apply: generic scalar kernel (any C, any HW; used for the 1-D case HW == 1 and odd shapes) ----------
__global__ void __launch_bounds__(256) invconv_apply_scalar(const float* __restrict__ zin, float* __restrict__ zout,
                                                           const float* ldj_in, float* ldj_out,
                                                           const float* __restrict__ M, const float* __restrict__ log_s,
                                                           float sign, int B, int C, int HW) {
    if (ldj_out && static_cast<long long>(blockIdx.x) * blockDim.x < B) {
        float part = 0.f;
        for (int c = threadIdx.x & 31; c < C; c += 32) part += __ldg(log_s + c);
        part = warp_sum(part);
        const int b = blockIdx.x * blockDim.x + threadIdx.x;
        if (b < B) ldj_out[b] = __fadd_rn(ldj_in[b], __fmul_rn(sign, __fmul_rn(part, static_cast<float>(HW))));
    }
    const long long total = static_cast<long long>(B) * C * HW;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long b = i / (static_cast<long long>(C) * HW);
        const int r = static_cast<int>(i - b * C * HW);
        const int co = r / HW, p = r - co * HW;
        const float* zb = zin + b * C * HW + p;
        float acc = 0.f;
        for (int ci = 0; ci < C; ++ci) acc = fmaf(__ldg(M + co * C + ci), __ldg(zb + static_cast<size_t>(ci) * HW), acc);
        zout[i] = acc;
    }
}

// ---- apply: tiled kernel.  CTA = one sample x TP pixels; thread = 4 pixels x OCG output channels ----------
// shared: Mt[ci][co] (transposed so OCG consecutive co are one vector load), zs[ci][TP].
// ACTNORM: the preceding ActNorm layer (modules.py:246-249) is applied while z is staged into shared memory,
// (z - bias)/exp(log_scale) and its log-det term first -- bit-identical to running the two layers back to back.
template <int OCG>
__device__ __forceinline__ void load_m(float (&m)[OCG], const float* p) {
    if (OCG == 8) {
        const float4 a = ld4(p), b = ld4(p + 4);
        m[0] = a.x; m[1 % OCG] = a.y; m[2 % OCG] = a.z; m[3 % OCG] = a.w;
        m[4 % OCG] = b.x; m[5 % OCG] = b.y; m[6 % OCG] = b.z; m[7 % OCG] = b.w;
    } else if (OCG == 4) {
        const float4 a = ld4(p);
        m[0] = a.x; m[1 % OCG] = a.y; m[2 % OCG] = a.z; m[3 % OCG] = a.w;
    } else if (OCG == 2) {
        const float2 a = *reinterpret_cast<const float2*>(p);
        m[0] = a.x; m[1 % OCG] = a.y;
    } else {
#pragma unroll
        for (int o = 0; o < OCG; ++o) m[o] = p[o];
    }
}

template <int OCG, int ACTNORM>
__global__ void __launch_bounds__(512) invconv_apply_tiled(const float* __restrict__ zin, float* __restrict__ zout,
                                                          const float* ldj_in, float* ldj_out,
                                                          const float* __restrict__ M, const float* __restrict__ log_s,
                                                          const float* __restrict__ an_log_scale,
                                                          const float* __restrict__ an_bias, float sign, int B, int C,
                                                          int HW, int TP) {
    extern __shared__ __align__(16) float sm[];
    float* Mt = sm;                          // C*C
    float* zs = sm + ((C * C + 3) & ~3);     // C*TP, 16-byte aligned

    // Mt[ci][co] = M[co][ci]: contiguous (conflict-free) shared stores, strided reads of the small L1-resident matrix;
    // staged ONCE per CTA -- the CTA then walks over (sample, pixel tile) units (at streaming batch sizes re-staging a
    // 48 x 48 matrix per 48 x 64 tile had doubled the CTA's loads)
    for (int i = threadIdx.x; i < C * C; i += blockDim.x) {
        const int ci = i / C, co = i - ci * C;
        Mt[i] = __ldg(M + co * C + ci);
    }
    const int tiles = (HW + TP - 1) / TP;
    const long long units = static_cast<long long>(B) * tiles;
    for (long long unit = blockIdx.x; unit < units; unit += gridDim.x) {
    const int b = static_cast<int>(unit / tiles);
    const int p0 = static_cast<int>(unit - static_cast<long long>(b) * tiles) * TP;
    const int tp = (HW - p0) < TP ? (HW - p0) : TP;  // multiple of 4
    if (unit != blockIdx.x) __syncthreads();  // everyone finished reading zs of the previous unit
    const float* zb = zin + (static_cast<size_t>(b) * C) * HW + p0;
    const int tpv = tp >> 2;
    for (int i = threadIdx.x; i < C * tpv; i += blockDim.x) {
        const int ci = i / tpv, pv = i - ci * tpv;
        float4 v = ldg4(zb + static_cast<size_t>(ci) * HW + 4 * pv);
        if (ACTNORM == 1) {
            const float e = expf(__ldg(an_log_scale + ci)), bi = __ldg(an_bias + ci);
            v.x = __fdiv_rn(__fsub_rn(v.x, bi), e);
            v.y = __fdiv_rn(__fsub_rn(v.y, bi), e);
            v.z = __fdiv_rn(__fsub_rn(v.z, bi), e);
            v.w = __fdiv_rn(__fsub_rn(v.w, bi), e);
        }
        st4(zs + ci * TP + 4 * pv, v);
    }
    if (ldj_out && p0 == 0 && threadIdx.x < 32) {
        float part = 0.f, an = 0.f;
        for (int c = threadIdx.x; c < C; c += 32) {
            part += __ldg(log_s + c);
            if (ACTNORM == 1) an -= __ldg(an_log_scale + c);
            if (ACTNORM == 2) an += __ldg(an_log_scale + c);
        }
        part = warp_sum(part);
        if (ACTNORM) an = warp_sum(an);
        if (threadIdx.x == 0) {
            float l = ldj_in[b];
            if (ACTNORM == 1) l = __fadd_rn(l, __fmul_rn(an, static_cast<float>(HW)));     // ActNorm.forward first
            l = __fadd_rn(l, __fmul_rn(sign, __fmul_rn(part, static_cast<float>(HW))));
            if (ACTNORM == 2) l = __fadd_rn(l, __fmul_rn(an, static_cast<float>(HW)));     // ActNorm.backward after the conv
            ldj_out[b] = l;
        }
    }
    __syncthreads();

    const int nog = C / OCG;
    for (int w = threadIdx.x; w < nog * tpv; w += blockDim.x) {
        const int og = w / tpv, pv = w - og * tpv;  // consecutive threads -> consecutive pixels, same og (broadcast)
        float acc[OCG][4];
#pragma unroll
        for (int o = 0; o < OCG; ++o) acc[o][0] = acc[o][1] = acc[o][2] = acc[o][3] = 0.f;
        const float* mrow = Mt + og * OCG;
        const float* zcol = zs + 4 * pv;
#pragma unroll 4
        for (int ci = 0; ci < C; ++ci) {
            const float4 z = ld4(zcol + ci * TP);
            float m[OCG];
            load_m<OCG>(m, mrow + ci * C);
#pragma unroll
            for (int o = 0; o < OCG; ++o) {
                acc[o][0] = fmaf(m[o], z.x, acc[o][0]);
                acc[o][1] = fmaf(m[o], z.y, acc[o][1]);
                acc[o][2] = fmaf(m[o], z.z, acc[o][2]);
                acc[o][3] = fmaf(m[o], z.w, acc[o][3]);
            }
        }
        float* ob = zout + (static_cast<size_t>(b) * C + og * OCG) * HW + p0 + 4 * pv;
#pragma unroll
        for (int o = 0; o < OCG; ++o) {
            if (ACTNORM == 2) {  // ActNorm.backward (modules.py:253): y * exp(log_scale) + bias
                const float e = expf(__ldg(an_log_scale + og * OCG + o)), bi = __ldg(an_bias + og * OCG + o);
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[o][q] = __fadd_rn(__fmul_rn(acc[o][q], e), bi);
            }
            st4(ob + static_cast<size_t>(o) * HW, make_float4(acc[o][0], acc[o][1], acc[o][2], acc[o][3]));
        }
    }
    }  // units
}

// ---- apply, tiny C (3 and 12: the first two levels of Glow on 32x32): no shared-memory staging of z and no barrier in
// the loop.  thread = 4 pixels x ALL channels: C float4 loads in flight, C x C x 4 fmaf in the same ascending-ci chains as
// invconv_apply_tiled (bit-identical), C float4 stores; the matrix is read from shared memory as warp-wide broadcasts.
template <int C, int ACTNORM>
__global__ void __launch_bounds__(256) invconv_apply_reg(const float* __restrict__ zin, float* __restrict__ zout,
                                                        const float* ldj_in, float* ldj_out,
                                                        const float* __restrict__ M, const float* __restrict__ log_s,
                                                        const float* __restrict__ an_log_scale,
                                                        const float* __restrict__ an_bias, float sign, int B, int HW) {
    constexpr int CP = (C + 3) & ~3;  // matrix rows padded to whole float4s
    __shared__ __align__(16) float Ms[C * CP];
    __shared__ float an_e[C], an_b[C];
    for (int i = threadIdx.x; i < C * CP; i += blockDim.x) {
        const int co = i / CP, ci = i - co * CP;
        Ms[i] = ci < C ? __ldg(M + co * C + ci) : 0.f;
    }
    if (ACTNORM)
        for (int c = threadIdx.x; c < C; c += blockDim.x) { an_e[c] = expf(__ldg(an_log_scale + c)); an_b[c] = __ldg(an_bias + c); }
    if (ldj_out && static_cast<long long>(blockIdx.x) * blockDim.x < B) {  // the leading CTAs own one sample per thread
        float part = 0.f, an = 0.f;
        for (int c = threadIdx.x & 31; c < C; c += 32) {
            part += __ldg(log_s + c);
            if (ACTNORM == 1) an -= __ldg(an_log_scale + c);
            if (ACTNORM == 2) an += __ldg(an_log_scale + c);
        }
        part = warp_sum(part);
        if (ACTNORM) an = warp_sum(an);
        const int b = blockIdx.x * blockDim.x + threadIdx.x;
        if (b < B) {
            float l = ldj_in[b];
            if (ACTNORM == 1) l = __fadd_rn(l, __fmul_rn(an, static_cast<float>(HW)));
            l = __fadd_rn(l, __fmul_rn(sign, __fmul_rn(part, static_cast<float>(HW))));
            if (ACTNORM == 2) l = __fadd_rn(l, __fmul_rn(an, static_cast<float>(HW)));
            ldj_out[b] = l;
        }
    }
    __syncthreads();
    const int qpr = HW >> 2;
    const long long total = static_cast<long long>(B) * qpr;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long b = i / qpr;
        const int q = static_cast<int>(i - b * qpr);
        const float* zb = zin + static_cast<size_t>(b) * C * HW + 4 * q;
        float4 v[C];
#pragma unroll
        for (int ci = 0; ci < C; ++ci) v[ci] = ldg4(zb + static_cast<size_t>(ci) * HW);
        if (ACTNORM == 1) {
#pragma unroll
            for (int ci = 0; ci < C; ++ci) {
                const float e = an_e[ci], bi = an_b[ci];
                v[ci].x = __fdiv_rn(__fsub_rn(v[ci].x, bi), e);
                v[ci].y = __fdiv_rn(__fsub_rn(v[ci].y, bi), e);
                v[ci].z = __fdiv_rn(__fsub_rn(v[ci].z, bi), e);
                v[ci].w = __fdiv_rn(__fsub_rn(v[ci].w, bi), e);
            }
        }
        float* ob = zout + static_cast<size_t>(b) * C * HW + 4 * q;
#pragma unroll
        for (int co = 0; co < C; ++co) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int c4 = 0; c4 < CP / 4; ++c4) {
                const float4 m = ld4(Ms + co * CP + 4 * c4);
                const float mm[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int ci = 4 * c4 + k;
                    if (ci < C) {
                        acc.x = fmaf(mm[k], v[ci].x, acc.x);
                        acc.y = fmaf(mm[k], v[ci].y, acc.y);
                        acc.z = fmaf(mm[k], v[ci].z, acc.z);
                        acc.w = fmaf(mm[k], v[ci].w, acc.w);
                    }
                }
            }
            if (ACTNORM == 2) {
                const float e = an_e[co], bi = an_b[co];
                acc.x = __fadd_rn(__fmul_rn(acc.x, e), bi);
                acc.y = __fadd_rn(__fmul_rn(acc.y, e), bi);
                acc.z = __fadd_rn(__fmul_rn(acc.z, e), bi);
                acc.w = __fadd_rn(__fmul_rn(acc.w, e), bi);
            }
            st4(ob + static_cast<size_t>(co) * HW, acc);
        }
    }
}

template <int C, int ACTNORM>
static int launch_reg(const float* zin, float* zout, const float* ldj_in, float* ldj_out, const float* M,
                      const float* log_s, const float* an_ls, const float* an_b, float sign, int B, int HW,
                      cudaStream_t st) {
    const long long total = static_cast<long long>(B) * (HW / 4);
    const int threads = total < static_cast<long long>(kSMs) * 256 ? 128 : 256;  // small batches: spread over more SMs
    long long blocks = (total + threads - 1) / threads;
    const long long need = (B + threads - 1) / threads;
    if (blocks > kSMs * 8) blocks = kSMs * 8;
    if (blocks < need) blocks = need;
    invconv_apply_reg<C, ACTNORM><<<static_cast<int>(blocks), threads, 0, st>>>(zin, zout, ldj_in, ldj_out, M, log_s, an_ls, an_b,
                                                                          sign, B, HW);
    return launch_status();
}

// ---- apply, large C (multiples of 64; C = 192 is the last level of a 64x64 Glow): the C x C matrix no longer fits next to a
// useful pixel tile, and re-staging it per (sample, pixel tile) made the launch L2-bound by ~100x.  Here a CTA owns a tile of
// CT = 64 OUTPUT channels, keeps that slice of the matrix (C x 64) in shared memory for its whole life and walks over
// (sample, 64-pixel tile) units, persistent; thread = 4 output channels x 4 pixels, the same ascending-ci fmaf chain as
// invconv_apply_tiled (bit-identical results).
template <int ACTNORM>
__global__ void __launch_bounds__(256, 2) invconv_apply_cotile(const float* __restrict__ zin, float* __restrict__ zout,
                                                              const float* ldj_in, float* ldj_out,
                                                              const float* __restrict__ M, const float* __restrict__ log_s,
                                                              const float* __restrict__ an_log_scale,
                                                              const float* __restrict__ an_bias, float sign, int B, int C,
                                                              int HW, int TP) {
    constexpr int CT = 64;
    extern __shared__ __align__(16) float sm[];
    __shared__ float ldj_term[2];
    float* Mt = sm;             // [C][CT]: Mt[ci][co - co0] = M[co][ci]
    float* zs = sm + C * CT;    // [C][TP]
    float* ane = zs + C * TP;   // ACTNORM == 1: exp(log_scale)[C], bias[C]
    const int co0 = blockIdx.y * CT;
    for (int i = threadIdx.x; i < C * CT; i += blockDim.x) {
        const int co = i / C, ci = i - co * C;  // coalesced reads of M's rows
        Mt[ci * CT + co] = __ldg(M + static_cast<size_t>(co0 + co) * C + ci);
    }
    if (ACTNORM == 1)
        for (int c = threadIdx.x; c < C; c += blockDim.x) { ane[c] = expf(__ldg(an_log_scale + c)); ane[C + c] = __ldg(an_bias + c); }
    // the sample-independent log-det terms, once per CTA; the co-tile-0 CTAs then update ldj for a grid-strided set of samples
    if (ldj_out && blockIdx.y == 0 && threadIdx.x < 32) {
        float part = 0.f, an = 0.f;
        for (int c = threadIdx.x; c < C; c += 32) {
            part += __ldg(log_s + c);
            if (ACTNORM == 1) an -= __ldg(an_log_scale + c);
            if (ACTNORM == 2) an += __ldg(an_log_scale + c);
        }
        part = warp_sum(part);
        if (ACTNORM) an = warp_sum(an);
        if (threadIdx.x == 0) { ldj_term[0] = part; ldj_term[1] = an; }
    }
    __syncthreads();
    if (ldj_out && blockIdx.y == 0) {
        const float part = ldj_term[0], an = ldj_term[1];
        for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
            float l = ldj_in[b];
            if (ACTNORM == 1) l = __fadd_rn(l, __fmul_rn(an, static_cast<float>(HW)));
            l = __fadd_rn(l, __fmul_rn(sign, __fmul_rn(part, static_cast<float>(HW))));
            if (ACTNORM == 2) l = __fadd_rn(l, __fmul_rn(an, static_cast<float>(HW)));
            ldj_out[b] = l;
        }
    }
    // a unit = TP consecutive positions of the flat (sample, pixel) axis: small maps (4x4: 16 pixels) put several samples
    // into one tile so that every thread has work; a pixel quad never straddles a sample (HW % 4 == 0)
    const long long npos = static_cast<long long>(B) * HW;
    const long long units = (npos + TP - 1) / TP;
    const int tpv = TP >> 2;
    const int cg = threadIdx.x / tpv, pv = threadIdx.x - cg * tpv;  // 256 threads = (CT / 4 = 16 channel groups) x (TP / 4 quads)
    for (long long unit = blockIdx.x; unit < units; unit += gridDim.x) {
        const long long pos0 = unit * TP;
        if (unit != blockIdx.x) __syncthreads();  // everyone finished reading zs of the previous unit
        for (int i = threadIdx.x; i < C * tpv; i += blockDim.x) {
            const int ci = i / tpv, q = i - ci * tpv;
            const long long pos = pos0 + 4 * q;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pos < npos) {
                const long long b = pos / HW;
                const int pp = static_cast<int>(pos - b * HW);
                v = ldg4(zin + (static_cast<size_t>(b) * C + ci) * HW + pp);
                if (ACTNORM == 1) {
                    const float e = ane[ci], bi = ane[C + ci];
                    v.x = __fdiv_rn(__fsub_rn(v.x, bi), e);
                    v.y = __fdiv_rn(__fsub_rn(v.y, bi), e);
                    v.z = __fdiv_rn(__fsub_rn(v.z, bi), e);
                    v.w = __fdiv_rn(__fsub_rn(v.w, bi), e);
                }
            }
            st4(zs + ci * TP + 4 * q, v);
        }
        __syncthreads();
        const long long pos = pos0 + 4 * pv;
        if (pos < npos) {
            float acc[4][4];
#pragma unroll
            for (int o = 0; o < 4; ++o) acc[o][0] = acc[o][1] = acc[o][2] = acc[o][3] = 0.f;
            const float* mrow = Mt + cg * 4;
            const float* zcol = zs + 4 * pv;
#pragma unroll 8
            for (int ci = 0; ci < C; ++ci) {
                const float4 zv = ld4(zcol + ci * TP);
                const float4 m = ld4(mrow + ci * CT);
                const float mm[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    acc[o][0] = fmaf(mm[o], zv.x, acc[o][0]);
                    acc[o][1] = fmaf(mm[o], zv.y, acc[o][1]);
                    acc[o][2] = fmaf(mm[o], zv.z, acc[o][2]);
                    acc[o][3] = fmaf(mm[o], zv.w, acc[o][3]);
                }
            }
            const long long b = pos / HW;
            const int pp = static_cast<int>(pos - b * HW);
            float* ob = zout + (static_cast<size_t>(b) * C + co0 + cg * 4) * HW + pp;
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                if (ACTNORM == 2) {
                    const float e = expf(__ldg(an_log_scale + co0 + cg * 4 + o)), bi = __ldg(an_bias + co0 + cg * 4 + o);
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[o][q] = __fadd_rn(__fmul_rn(acc[o][q], e), bi);
                }
                st4(ob + static_cast<size_t>(o) * HW, make_float4(acc[o][0], acc[o][1], acc[o][2], acc[o][3]));
            }
        }
    }
}

template <int ACTNORM>
static int launch_cotile(const float* zin, float* zout, const float* ldj_in, float* ldj_out, const float* M,
                         const float* log_s, const float* an_ls, const float* an_b, float sign, int B, int C, int HW,
                         cudaStream_t st) {
    constexpr int CT = 64, TP = 64;  // 16 channel groups x 16 pixel quads = 256 threads
    const size_t smem = (static_cast<size_t>(C) * CT + static_cast<size_t>(C) * TP + 2 * C) * 4;
    if (smem > 110 * 1024) return -100;  // two CTAs per SM
    auto kern = invconv_apply_cotile<ACTNORM>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    const int ncot = C / CT;
    const long long units = (static_cast<long long>(B) * HW + TP - 1) / TP;
    long long gx = (2LL * kSMs + ncot - 1) / ncot;  // ~2 CTAs per SM in total
    if (gx > units) gx = units;
    dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>(ncot));
    kern<<<grid, 256, smem, st>>>(zin, zout, ldj_in, ldj_out, M, log_s, an_ls, an_b, sign, B, C, HW, TP);
    return launch_status();
}

template <int OCG, int ACTNORM>
static int launch_tiled(const float* zin, float* zout, const float* ldj_in, float* ldj_out, const float* M,
                        const float* log_s, const float* an_ls, const float* an_b, float sign, int B, int C, int HW,
                        cudaStream_t st) {
    // pixel tile: whole sample if it fits, else the largest multiple of 4 that keeps shared memory <= ~96 KB
    int TP = HW;
    const size_t budget = 96 * 1024;
    const size_t mt = (static_cast<size_t>(C) * C + 3) & ~static_cast<size_t>(3);
    while ((mt + static_cast<size_t>(C) * TP) * 4 > budget && TP > 16) TP = ((TP / 2 + 3) / 4) * 4;
    const size_t smem = (mt + static_cast<size_t>(C) * TP) * 4;
    if (smem > 200 * 1024) return -100;  // caller falls back to the scalar kernel
    auto kern = invconv_apply_tiled<OCG, ACTNORM>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    const int work = (C / OCG) * (TP / 4);
    int threads = work >= 512 ? 512 : ((work + 31) / 32) * 32;
    if (threads < 64) threads = 64;
    const long long units = static_cast<long long>(B) * ((HW + TP - 1) / TP);
    int per_sm = static_cast<int>((200 * 1024) / (smem + 1024));  // resident CTAs by shared memory ...
    const int by_threads = 2048 / threads;                         // ... and by threads
    if (per_sm > by_threads) per_sm = by_threads;
    if (per_sm < 1) per_sm = 1;
    const long long cap = static_cast<long long>(kSMs) * per_sm;
    const int grid = static_cast<int>(units < cap ? units : cap);
    kern<<<grid, threads, smem, st>>>(zin, zout, ldj_in, ldj_out, M, log_s, an_ls, an_b, sign, B, C, HW, TP);
    return launch_status();
}

template <int ACTNORM>
static int apply_dispatch(const float* z_in, float* z_out, const float* ldj_in, float* ldj_out, const float* M,
                          const float* log_s, const float* an_ls, const float* an_b, float sign, int B, int C, int HW,
                          cudaStream_t st) {
    if (HW % 4 == 0 && aligned16(z_in) && aligned16(z_out) && C % 64 == 0 && C >= 128) {
        const int rc = launch_cotile<ACTNORM>(z_in, z_out, ldj_in, ldj_out, M, log_s, an_ls, an_b, sign, B, C, HW, st);
        if (rc != -100) return rc;
    }
    if (HW % 4 == 0 && aligned16(z_in) && aligned16(z_out)) {
        if (C == 3) return launch_reg<3, ACTNORM>(z_in, z_out, ldj_in, ldj_out, M, log_s, an_ls, an_b, sign, B, HW, st);
        if (C == 12) return launch_reg<12, ACTNORM>(z_in, z_out, ldj_in, ldj_out, M, log_s, an_ls, an_b, sign, B, HW, st);
        int rc;
        if (C % 8 == 0 && C >= 96) rc = launch_tiled<8, ACTNORM>(z_in, z_out, ldj_in, ldj_out, M, log_s, an_ls, an_b, sign, B, C, HW, st);
        else if (C % 4 == 0) rc = launch_tiled<4, ACTNORM>(z_in, z_out, ldj_in, ldj_out, M, log_s, an_ls, an_b, sign, B, C, HW, st);
        else if (C % 3 == 0) rc = launch_tiled<3, ACTNORM>(z_in, z_out, ldj_in, ldj_out, M, log_s, an_ls, an_b, sign, B, C, HW, st);
        else if (C % 2 == 0) rc = launch_tiled<2, ACTNORM>(z_in, z_out, ldj_in, ldj_out, M, log_s, an_ls, an_b, sign, B, C, HW, st);
        else rc = launch_tiled<1, ACTNORM>(z_in, z_out, ldj_in, ldj_out, M, log_s, an_ls, an_b, sign, B, C, HW, st);
        if (rc != -100) return rc;
    }
    return -100;
}

}  // namespace nfb

using namespace nfb;

extern "C" int nfb_invconv1x1_weight(const float* P, const float* L, const float* U, const float* log_s,
                                     const float* sign_s, float* W_out, float* Winv_out, int C, nfb_stream_t stream) {
    if (!P || !L || !U || !log_s || !sign_s || !W_out) return NFB_ERR_NULL;
    if (C <= 0) return NFB_ERR_SHAPE;
    if (C > kMaxInvconvC) return NFB_ERR_UNSUPPORTED;
    const int ctas = (C * C + 255) / 256 < 64 ? (C * C + 255) / 256 : 64;
    invconv_weight_kernel<<<ctas, 256, 0, as_stream(stream)>>>(P, L, U, log_s, sign_s, W_out, Winv_out, C);
    return launch_status();
}

extern "C" int nfb_invconv1x1_apply(const float* z_in, float* z_out, const float* ldj_in, float* ldj_out,
                                    const float* M, const float* log_s, float sign, int B, int C, int HW,
                                    nfb_stream_t stream) {
    if (!z_in || !z_out || !M) return NFB_ERR_NULL;
    if (ldj_out && (!ldj_in || !log_s)) return NFB_ERR_NULL;  // ldj_out == NULL: plain matrix apply (the gradient W^T gy)
    if (B <= 0 || C <= 0 || HW <= 0) return NFB_ERR_SHAPE;
    if (z_in == z_out) return NFB_ERR_UNSUPPORTED;
    cudaStream_t st = as_stream(stream);
    const int rc = apply_dispatch<0>(z_in, z_out, ldj_in, ldj_out, M, log_s, nullptr, nullptr, sign, B, C, HW, st);
    if (rc != -100) return rc;
    const long long total = static_cast<long long>(B) * C * HW;
    long long blocks = (total + 255) / 256;
    const long long need = (B + 255) / 256;
    if (blocks > kSMs * 16) blocks = kSMs * 16;
    if (blocks < need) blocks = need;
    invconv_apply_scalar<<<static_cast<int>(blocks), 256, 0, st>>>(z_in, z_out, ldj_in, ldj_out, M, log_s, sign, B, C, HW);
    return launch_status();
}

extern "C" int nfb_actnorm_invconv_fwd(const float* z_in, float* z_out, const float* ldj_in, float* ldj_out,
                                       const float* log_scale, const float* bias, const float* W, const float* log_s,
                                       int B, int C, int HW, nfb_stream_t stream) {
    if (!z_in || !z_out || !ldj_in || !ldj_out || !log_scale || !bias || !W || !log_s) return NFB_ERR_NULL;
    if (B <= 0 || C <= 0 || HW <= 0) return NFB_ERR_SHAPE;
    if (z_in == z_out) return NFB_ERR_UNSUPPORTED;
    const int rc = apply_dispatch<1>(z_in, z_out, ldj_in, ldj_out, W, log_s, log_scale, bias, 1.f, B, C, HW,
                                        as_stream(stream));
    return rc == -100 ? NFB_ERR_UNSUPPORTED : rc;  // caller runs the two layers separately for odd shapes
}

extern "C" int nfb_invconv_actnorm_inv(const float* y_in, float* y_out, const float* ldj_in, float* ldj_out,
                                       const float* Winv, const float* log_s, const float* log_scale, const float* bias,
                                       int B, int C, int HW, nfb_stream_t stream) {
    if (!y_in || !y_out || !ldj_in || !ldj_out || !Winv || !log_s || !log_scale || !bias) return NFB_ERR_NULL;
    if (B <= 0 || C <= 0 || HW <= 0) return NFB_ERR_SHAPE;
    if (y_in == y_out) return NFB_ERR_UNSUPPORTED;
    const int rc = apply_dispatch<2>(y_in, y_out, ldj_in, ldj_out, Winv, log_s, log_scale, bias, -1.f, B, C, HW,
                                     as_stream(stream));
    return rc == -100 ? NFB_ERR_UNSUPPORTED : rc;
}
