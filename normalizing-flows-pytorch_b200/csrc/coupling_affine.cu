// coupling_affine.cu -- AffineCoupling / AdditiveCoupling bijection fused with split-gather,
// merge-scatter and the per-sample log|det J| reduction (reference: coupling.py:52-122,
// squeeze.py:5-83).  One pass over z and params; nothing is materialised in between.
//
// Work decomposition: a sample ("row") is cut into items of 8 consecutive floats of z in its ORIGINAL
// layout.  Every item is two 16-byte loads of z, up to two (t, s_raw) float4 pairs from the conditioner
// output, and two 16-byte stores -- all fully coalesced.  Which of the 8 floats belong to the transformed
// half z0 and where their (t, s) live follows from the split mode (DESIGN.md "index formulas").
#include "common.cuh"

namespace nfb {

template <bool INV>
__device__ __forceinline__ float affine_elem(float& z, float t, float sraw, float a, float b) {
    // coupling.py:107: s = tanh(raw) * s_log_scale + s_bias   (two rounded ops in the reference, no FMA)
    const float s = __fadd_rn(__fmul_rn(tanhf(sraw), a), b);
    if (!INV)
        z = __fadd_rn(__fmul_rn(z, expf(s)), t);  // coupling.py:109
    else
        z = __fmul_rn(expf(-s), __fsub_rn(z, t));  // coupling.py:119
    return s;
}

template <bool INV>
__device__ __forceinline__ float affine_vec4(float& z0, float& z1, float& z2, float& z3, const float4& t,
                                             const float4& s, float a, float b) {
    float acc = affine_elem<INV>(z0, t.x, s.x, a, b);
    acc += affine_elem<INV>(z1, t.y, s.y, a, b);
    acc += affine_elem<INV>(z2, t.z, s.z, a, b);
    acc += affine_elem<INV>(z3, t.w, s.w, a, b);
    return acc;
}

// ---- vectorised item functors -------------------------------------------------------------------
template <int MODE, bool INV>
struct AffineVec {
    const float* zin;
    float* zout;
    const float* __restrict__ params;
    const float* __restrict__ pa;
    const float* __restrict__ pb;
    SplitGeom g;
    int items;
    bool inplace;
    bool pow2;          // checkerboard: h*w/4 and w/4 are powers of two -> shifts instead of divisions
    int sh_ipl, sh_wq;

    __device__ __forceinline__ float finish(float acc) const { return INV ? -acc : acc; }

    __device__ __forceinline__ float operator()(int row, int it) const {
        const float a = __ldg(pa), b = __ldg(pb);
        const size_t base = static_cast<size_t>(row) * g.D;
        const float* zr = zin + base;
        float* zo = zout + base;
        const float* pr = params + base;  // params row stride = 2*n0 = D
        float acc = 0.f;
        if (MODE == NFB_SPLIT_CHANNEL) {
            // item = 8 transformed floats + the 8 pass-through floats at the same offset of the other half
            const int o0 = (g.odd ? g.n0 : 0) + 8 * it;
            const int o1 = (g.odd ? 0 : g.n0) + 8 * it;
            float4 v0 = ld4(zr + o0), v1 = ld4(zr + o0 + 4);
            const float4 t0 = ldg4(pr + 8 * it), t1 = ldg4(pr + 8 * it + 4);
            const float4 s0 = ldg4(pr + g.n0 + 8 * it), s1 = ldg4(pr + g.n0 + 8 * it + 4);
            float4 c0, c1;
            if (!inplace) { c0 = ld4(zr + o1); c1 = ld4(zr + o1 + 4); }
            acc += affine_vec4<INV>(v0.x, v0.y, v0.z, v0.w, t0, s0, a, b);
            acc += affine_vec4<INV>(v1.x, v1.y, v1.z, v1.w, t1, s1, a, b);
            st4(zo + o0, v0); st4(zo + o0 + 4, v1);
            if (!inplace) { st4(zo + o1, c0); st4(zo + o1 + 4, c1); }
        } else if (MODE == NFB_SPLIT_1D) {
            // floats 8it..8it+7 = (a0 b0 a1 b1 a2 b2 a3 b3); the a's are entries 4it..4it+3 of the even half
            float4 v0 = ld4(zr + 8 * it), v1 = ld4(zr + 8 * it + 4);
            const float4 t = ldg4(pr + 4 * it), s = ldg4(pr + g.n0 + 4 * it);
            if (!g.odd) acc += affine_vec4<INV>(v0.x, v0.z, v1.x, v1.z, t, s, a, b);
            else        acc += affine_vec4<INV>(v0.y, v0.w, v1.y, v1.w, t, s, a, b);
            st4(zo + 8 * it, v0); st4(zo + 8 * it + 4, v1);
        } else {
            // checkerboard.  Items are ordered (c, dy, i, jb): item = 8 consecutive x (jb) of input line y = 2i+dy of
            // channel c, so all items of one (c, dy) -- h*w/4 of them, >= a warp for 32x32 images -- share the two
            // squeezed channels k = 4c+2dy (even x) and k+1 (odd x): the z0/z1 classification is warp-uniform and
            // costs three compares instead of integer divisions.
            const int wq = g.w >> 2;                 // items per line (W/8)
            const int ipl = g.h * wq;                // items per (c, dy)
            int cd, rem, i, jb;
            if (pow2) { cd = it >> sh_ipl; rem = it & (ipl - 1); i = rem >> sh_wq; jb = rem & (wq - 1); }
            else { cd = it / ipl; rem = it - cd * ipl; i = rem / wq; jb = rem - i * wq; }
            const int k = 2 * cd;                    // 4c + 2dy
            const int e0 = (cd >> 1) * g.HW + (2 * i + (cd & 1)) * g.W + 8 * jb;
            const int qe = (k >= g.C) + (k >= 2 * g.C) + (k >= 3 * g.C);
            const int qo = (k + 1 >= g.C) + (k + 1 >= 2 * g.C) + (k + 1 >= 3 * g.C);
            const bool te = ((qe == 0 || qe == 3) ? 1 : 0) != g.odd, to = ((qo == 0 || qo == 3) ? 1 : 0) != g.odd;
            const int me = (qe == 0) ? k : (qe == 3) ? k - 2 * g.C : k - g.C;
            const int mo = (qo == 0) ? k + 1 : (qo == 3) ? k + 1 - 2 * g.C : k + 1 - g.C;
            const int hw = g.h * g.w;
            const int sp = i * g.w + 4 * jb;
            float4 v0, v1, t_e, s_e, t_o, s_o;
            if (te || to || !inplace) { v0 = ld4(zr + e0); v1 = ld4(zr + e0 + 4); }
            if (te) { t_e = ldg4(pr + me * hw + sp); s_e = ldg4(pr + g.n0 + me * hw + sp); }
            if (to) { t_o = ldg4(pr + mo * hw + sp); s_o = ldg4(pr + g.n0 + mo * hw + sp); }
            if (te) acc += affine_vec4<INV>(v0.x, v0.z, v1.x, v1.z, t_e, s_e, a, b);
            if (to) acc += affine_vec4<INV>(v0.y, v0.w, v1.y, v1.w, t_o, s_o, a, b);
            if (te || to || !inplace) { st4(zo + e0, v0); st4(zo + e0 + 4, v1); }
        }
        return acc;
    }
};

// ---- scalar fallback: any shape the reference accepts (e.g. the 2-D moons model, W = 2 or 4) -------
template <int MODE, bool INV>
struct AffineScalar {
    const float* zin;
    float* zout;
    const float* __restrict__ params;
    const float* __restrict__ pa;
    const float* __restrict__ pb;
    SplitGeom g;
    int items;
    bool inplace;

    __device__ __forceinline__ float finish(float acc) const { return INV ? -acc : acc; }

    __device__ __forceinline__ float operator()(int row, int e) const {
        const size_t base = static_cast<size_t>(row) * g.D;
        int idx;
        const bool tr = classify<MODE>(g, e, idx);
        float z = zin[base + e];
        float s = 0.f;
        if (tr) {
            const float* pr = params + base;
            s = affine_elem<INV>(z, __ldg(pr + idx), __ldg(pr + g.n0 + idx), __ldg(pa), __ldg(pb));
        }
        if (tr || !inplace) zout[base + e] = z;
        return s;
    }
};

template <int MODE, bool INV>
static int launch_affine(const float* z_in, float* z_out, const float* params, const float* ldj_in, float* ldj_out,
                         const float* a, const float* b, const SplitGeom& g, cudaStream_t st) {
    const bool inplace = (z_in == z_out);
    bool vec = aligned16(z_in) && aligned16(z_out) && aligned16(params);
    int items = 0;
    if (MODE == NFB_SPLIT_CHANNEL) { vec = vec && (g.n0 % 8 == 0); items = g.n0 / 8; }
    else if (MODE == NFB_SPLIT_1D) { vec = vec && (g.D % 8 == 0); items = g.D / 8; }
    else                           { vec = vec && (g.W % 8 == 0); items = g.D / 8; }
    if (vec) {
        AffineVec<MODE, INV> f{z_in, z_out, params, a, b, g, items, inplace, false, 0, 0};
        if (MODE == NFB_SPLIT_CHECKER) {
            const int wq = g.w / 4, ipl = g.h * wq;
            auto is_pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
            auto lg = [](int v) { int s = 0; while ((1 << s) < v) ++s; return s; };
            f.pow2 = is_pow2(wq) && is_pow2(ipl);
            f.sh_ipl = lg(ipl);
            f.sh_wq = lg(wq);
        }
        return launch_rows(f, ldj_in, ldj_out, g.B, st);
    }
    AffineScalar<MODE, INV> f{z_in, z_out, params, a, b, g, g.D, inplace};
    return launch_rows(f, ldj_in, ldj_out, g.B, st);
}

template <bool INV>
static int affine_entry(const float* z_in, float* z_out, const float* params, const float* ldj_in, float* ldj_out,
                        const float* a, const float* b, int B, int C, int H, int W, int mode, int odd,
                        nfb_stream_t stream) {
    if (!z_in || !z_out || !params || !ldj_in || !ldj_out || !a || !b) return NFB_ERR_NULL;
    SplitGeom g;
    const int rc = make_geom(g, B, C, H, W, mode, odd);
    if (rc != NFB_OK) return rc;
    cudaStream_t st = as_stream(stream);
    switch (mode) {
        case NFB_SPLIT_1D: return launch_affine<NFB_SPLIT_1D, INV>(z_in, z_out, params, ldj_in, ldj_out, a, b, g, st);
        case NFB_SPLIT_CHECKER: return launch_affine<NFB_SPLIT_CHECKER, INV>(z_in, z_out, params, ldj_in, ldj_out, a, b, g, st);
        default: return launch_affine<NFB_SPLIT_CHANNEL, INV>(z_in, z_out, params, ldj_in, ldj_out, a, b, g, st);
    }
}

// ---- additive coupling (no log-det): flat elementwise over all samples ------------------------------
template <int MODE>
__global__ void additive_kernel(const float* zin, float* zout, const float* __restrict__ t, float sign, SplitGeom g,
                                bool inplace) {
    const long long total = static_cast<long long>(g.B) * g.D;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int row = static_cast<int>(i / g.D);
        const int e = static_cast<int>(i - static_cast<long long>(row) * g.D);
        int idx;
        const bool tr = classify<MODE>(g, e, idx);
        float z = zin[i];
        if (tr) z = __fadd_rn(z, __fmul_rn(sign, __ldg(t + static_cast<size_t>(row) * g.n0 + idx)));
        if (tr || !inplace) zout[i] = z;
    }
}

}  // namespace nfb

using namespace nfb;

extern "C" int nfb_affine_coupling_fwd(const float* z_in, float* z_out, const float* params, const float* ldj_in,
                                       float* ldj_out, const float* s_log_scale, const float* s_bias, int B, int C,
                                       int H, int W, int mode, int odd, nfb_stream_t stream) {
    return affine_entry<false>(z_in, z_out, params, ldj_in, ldj_out, s_log_scale, s_bias, B, C, H, W, mode, odd, stream);
}

extern "C" int nfb_affine_coupling_inv(const float* z_in, float* z_out, const float* params, const float* ldj_in,
                                       float* ldj_out, const float* s_log_scale, const float* s_bias, int B, int C,
                                       int H, int W, int mode, int odd, nfb_stream_t stream) {
    return affine_entry<true>(z_in, z_out, params, ldj_in, ldj_out, s_log_scale, s_bias, B, C, H, W, mode, odd, stream);
}

extern "C" int nfb_additive_coupling(const float* z_in, float* z_out, const float* params, float sign, int B, int C,
                                     int H, int W, int mode, int odd, nfb_stream_t stream) {
    if (!z_in || !z_out || !params) return NFB_ERR_NULL;
    SplitGeom g;
    const int rc = make_geom(g, B, C, H, W, mode, odd);
    if (rc != NFB_OK) return rc;
    const long long total = static_cast<long long>(B) * g.D;
    const int grid = static_cast<int>(total / 256 + 1 < kSMs * 8 ? total / 256 + 1 : kSMs * 8);
    const bool inplace = z_in == z_out;
    cudaStream_t st = as_stream(stream);
    if (mode == NFB_SPLIT_1D) additive_kernel<NFB_SPLIT_1D><<<grid, 256, 0, st>>>(z_in, z_out, params, sign, g, inplace);
    else if (mode == NFB_SPLIT_CHECKER) additive_kernel<NFB_SPLIT_CHECKER><<<grid, 256, 0, st>>>(z_in, z_out, params, sign, g, inplace);
    else additive_kernel<NFB_SPLIT_CHANNEL><<<grid, 256, 0, st>>>(z_in, z_out, params, sign, g, inplace);
    return launch_status();
}
