// conditioner_tc.cu -- the ConvNet conditioner (modules.py:416-438) on the 5th-generation tensor cores.
//
// Every 3x3 / 1x1 layer is an implicit GEMM issued with tcgen05.mma (kind::tf32, M=128, N=32, K=8) by ONE thread,
// accumulators in TMEM, operands in shared memory.  Precision: single-pass TF32 misses the 1e-5 bits/dim bar
// (SURVEY.md F8), so every product is error-compensated ("3xTF32"): x = x_hi + x_lo with x_hi the top 19 bits,
//   a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo      (fp32 accumulation in TMEM; the dropped a_lo*b_lo is ~2^-22 relative).
//
// "Flat shift" implicit GEMM: the CTA keeps the zero-padded activations of its S samples as ONE flat sequence of
// positions (padded image (H+2)x(W+2), samples back to back, G = W+3 guard positions at both ends), stored K-major
// WITHOUT swizzle as [ci/4][position][4 ci] -- the canonical UMMA layout ((8,m),2):((1,SBO),LBO) in 16-byte units with
// SBO = 8 positions and LBO = one channel-chunk plane.  A 3x3 tap is then just a row offset (ky-1)*(W+2)+(kx-1) added
// to the A descriptor's start address: the same buffer feeds all nine taps, no im2col, no shifted copies.  Outputs are
// computed for every flat position (padding positions included; 128 per tile) and the epilogue writes zeros back to
// the padding positions, which keeps the halo intact for the next layer.
//
// Accumulation: the tensor core adds into its fp32 accumulator with truncation, so a long chain of MMAs into one
// accumulator drifts by ~(number of MMAs) x 2^-24 (measured: 1.2e-5 with all 108 MMAs of a layer in one chain).  Each
// layer therefore uses FOUR TMEM accumulators per tile: one per kernel row for the a_hi*b_hi products (12 MMAs each) and
// one for the small compensation terms; the epilogue sums them in registers with round-to-nearest.  The residual stream
// lives in registers (the thread that owns TMEM lane m owns flat position 128t+m in every layer; one CTA per SM leaves
// 255 registers per thread).
// Per stage: [thread 0 issues 108*T MMAs] -> tcgen05.commit -> mbarrier -> [all 4 warps: tcgen05.ld their 32 TMEM lanes,
// bias/BN/ReLU, hi/lo split, st.shared into the activation planes] while cp.async streams the next stage's weights.
#include "conditioner.cuh"

namespace nfb {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    // bounded: a lost arrival must trap (cudaErrorLaunchFailure), never hang the GPU
    uint32_t done = 0;
    for (int it = 0; it < (1 << 22) && !done; ++it) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
    if (!done) __trap();
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 consecutive TMEM columns of this thread's lane -> registers
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, no swizzle: start address, leading (K-chunk) and stride (8-row group) byte offsets, all in 16-byte units
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFFu);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= static_cast<uint64_t>(1) << 46;  // descriptor version (Blackwell)
    return d;                              // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// instruction descriptor: D = F32, A = B = TF32, both K-major, N = 32, M = 128 (mma_sm100_desc.hpp bit layout)
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);

template <int H, int W, int S>
struct TcGeom {
    static constexpr int Wp = W + 2, Hp = H + 2, Lp = Hp * Wp, PTOT = S * Lp;
    static constexpr int T = (PTOT + 127) / 128;  // M tiles
    static constexpr int G = Wp + 1;              // guard positions on both ends
    static constexpr int PBUF = T * 128 + 2 * G;  // positions per channel-chunk plane
    static constexpr int CS = PBUF * 4;           // floats per plane
    static constexpr int TMEM_COLS = (4 * T * 32 <= 128) ? 128 : (4 * T * 32 <= 256) ? 256 : 512;  // 4 accumulators x T tiles
    static_assert(4 * T * 32 <= 512, "tile count exceeds TMEM");
    static constexpr size_t SMEM = (static_cast<size_t>(16) * CS + 2 * kTcStage + 1152 + 8) * sizeof(float);
};

template <int H, int W, int S, int MODE>
__global__ void __launch_bounds__(128, 1) convnet_tc_kernel(const float* __restrict__ zsrc, float* __restrict__ out,
                                                           const float* __restrict__ pk, SplitGeom g, int Cin, int Cout,
                                                           int B, int dbg) {
    using GM = TcGeom<H, W, S>;
    constexpr int Wp = GM::Wp, Lp = GM::Lp, PTOT = GM::PTOT, T = GM::T, G = GM::G, PBUF = GM::PBUF, CS = GM::CS;
    extern __shared__ __align__(128) float smem[];
    float* actH = smem;                 // [8][PBUF][4]
    float* actL = actH + 8 * CS;
    float* wH = actL + 8 * CS;          // one stage of weights, hi then lo
    float* wL = wH + kTcStage;
    float* cst = wL + kTcStage;         // per-channel constants (352 + CoutPad <= 1152 floats)
    uint64_t* mbar = reinterpret_cast<uint64_t*>(cst + 1152);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);

    const TcLayout TL = tc_layout(Cin, Cout);
    const int tid = threadIdx.x, warp = tid >> 5;

    auto load_weights = [&](const float* hi, const float* lo, int nfloats) {
        for (int i = tid * 4; i < nfloats; i += 128 * 4) { cp_async16(wH + i, hi + i); cp_async16(wL + i, lo + i); }
        cp_async_commit();
    };
    auto stage_ptr = [&](int s) { return pk + TL.stage0 + static_cast<size_t>(s) * 2 * kTcStage; };

    // ---- prologue: first weights in flight, zero the activation planes, constants, barrier, TMEM -----------------
    load_weights(stage_ptr(0), stage_ptr(0) + kTcStage, kTcStage);
    for (int i = tid * 4; i < 16 * CS; i += 128 * 4) st4(smem + i, make_float4(0.f, 0.f, 0.f, 0.f));
    for (int i = tid; i < 352 + TL.cout_pad && i < 1152; i += 128) cst[i] = __ldg(pk + TL.consts + i);
    if (tid == 0) mbar_init(mbar, 1);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(static_cast<uint32_t>(GM::TMEM_COLS)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    uint32_t phase = 0;

    const uint64_t dAH = make_smem_desc(smem_u32(actH), PBUF * 16, 128), dAL = make_smem_desc(smem_u32(actL), PBUF * 16, 128);
    const uint64_t dBH = make_smem_desc(smem_u32(wH), 512, 128), dBL = make_smem_desc(smem_u32(wL), 512, 128);

    // TMEM accumulators of tile t: region r in {0,1,2} = kernel row ky (a_hi*b_hi), region 3 = compensation terms
    auto region = [&](int r, int t) -> uint32_t { return tmem + static_cast<uint32_t>((r * T + t) * 32); };

    // thread 0: all MMAs of one 3x3 (ntaps = 9) or 1x1 (ntaps = 1, centre tap) layer
    auto issue_layer = [&](bool accumulate, int ksteps, int ntaps, uint32_t b_chunk_off) {
        tc_fence_after();
#pragma unroll 1
        for (int t = 0; t < T; ++t) {
            uint32_t acc_c = accumulate ? 1u : 0u;
#pragma unroll 1
            for (int tap = 0; tap < ntaps; ++tap) {
                const int ky = (ntaps == 9) ? tap / 3 : 0, kx = (ntaps == 9) ? tap % 3 : 0;
                const int delta = (ntaps == 9) ? ((ky - 1) * Wp + (kx - 1)) : 0;
                const uint32_t d_main = region(ky, t), d_comp = region(3, t);
                uint32_t acc_m = (accumulate || kx != 0) ? 1u : 0u;
#pragma unroll 1
                for (int j = 0; j < ksteps; ++j) {
                    const uint32_t a_off = static_cast<uint32_t>((2 * j) * PBUF + t * 128 + G + delta);  // 16-byte units
                    const uint32_t b_off = b_chunk_off + static_cast<uint32_t>((tap * 8 + 2 * j) * 32);
                    if (dbg & 1) continue;  // profiling knob: no tensor work
                    mma_tf32(d_main, dAH + a_off, dBH + b_off, kIdesc, acc_m);
                    if (dbg & 4) continue;  // profiling knob: single-pass TF32
                    mma_tf32(d_comp, dAL + a_off, dBH + b_off, kIdesc, acc_c);
                    mma_tf32(d_comp, dAH + a_off, dBL + b_off, kIdesc, 1u);
                    acc_m = 1u;
                    acc_c = 1u;
                }
            }
        }
        mma_commit(mbar);
    };
    // everyone: shared-memory writes -> visible to the tensor core, then one thread issues, everyone waits
    auto run_layer = [&](bool accumulate, int ksteps, int ntaps, uint32_t b_chunk_off) {
        cp_async_wait_all();
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) issue_layer(accumulate, ksteps, ntaps, b_chunk_off);
        mbar_wait(mbar, phase);
        phase ^= 1u;
        tc_fence_after();
    };

    // flat position -> (sample, y, x); true for an interior pixel of a valid sample
    auto locate = [&](int p, int& b, int& y, int& x) -> bool {
        if (p >= PTOT) return false;
        const int s = p / Lp, r = p - s * Lp;
        const int yy = r / Wp, xx = r - yy * Wp;
        b = blockIdx.x * S + s;
        y = yy - 1;
        x = xx - 1;
        return b < B && yy >= 1 && yy <= H && xx >= 1 && xx <= W;
    };

    // sum of the layer's accumulators for this thread's row of tile t (round-to-nearest adds in registers)
    auto gather_acc = [&](int t, int nmain, float (&v)[32]) {
        if (dbg & 2) nmain = 0;  // profiling knob: one TMEM load instead of four
        const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
        tmem_ld32(lane_base + region(3, t), v);
        for (int r = 0; r < nmain; ++r) {
            float u[32];
            tmem_ld32(lane_base + region(r, t), u);
#pragma unroll
            for (int c = 0; c < 32; ++c) v[c] += u[c];
        }
    };
    // a[32] (already activated; zero outside the image) -> hi/lo planes at flat position p
    auto store_act = [&](int p, const float (&a)[32]) {
        float* ph = actH + (p + G) * 4;
        float* pl = actL + (p + G) * 4;
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
            float hi[4], lo[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { hi[q] = tf32_hi(a[c4 * 4 + q]); lo[q] = a[c4 * 4 + q] - hi[q]; }
            st4(ph + c4 * CS, make_float4(hi[0], hi[1], hi[2], hi[3]));
            st4(pl + c4 * CS, make_float4(lo[0], lo[1], lo[2], lo[3]));
        }
    };

    // constants: b0 | blk0: sA tA b1' b2 | blk1: sA tA b1' b2 | sO tO | bout
    const float* c_b0 = cst;
    auto c_blk = [&](int blk, int k) { return cst + 32 + blk * 128 + k * 32; };
    const float* c_sO = cst + 288;
    const float* c_tO = cst + 320;
    const float* c_bout = cst + 352;

    float xres[T][32];  // residual stream of this thread's rows

    // ---- in conv: Cin -> 32 in passes of 32 input channels ---------------------------------------------------------
    int stage = 0;
    for (int c = 0; c < TL.n_in; ++c) {
        const int CI = (Cin - c * kF) < kF ? (Cin - c * kF) : kF;
        const int CI8 = (CI + 7) & ~7;
        for (int p = tid; p < PTOT; p += 128) {
            int b, y, x;
            const bool in = locate(p, b, y, x);
            for (int ci = 0; ci < CI8; ++ci) {
                float v = 0.f;
                if (in && ci < CI) {
                    const int j = (c * kF + ci) * (H * W) + y * W + x;
                    if (MODE < 0) v = __ldg(zsrc + static_cast<size_t>(b) * Cin * (H * W) + j);
                    else v = __ldg(zsrc + static_cast<size_t>(b) * g.D + half_offset<(MODE < 0 ? NFB_SPLIT_CHANNEL : MODE)>(g, j, 1));
                }
                const float hi = tf32_hi(v);
                const int o = (ci >> 2) * CS + (p + G) * 4 + (ci & 3);
                actH[o] = hi;
                actL[o] = v - hi;
            }
        }
        run_layer(c > 0, CI8 / 8, 9, 0);
        ++stage;
        load_weights(stage_ptr(stage), stage_ptr(stage) + kTcStage, kTcStage);  // a next stage always exists here
    }
    // x = conv0 + b0;  a = relu(BN_a0(x))
#pragma unroll
    for (int t = 0; t < T; ++t) {
        float v[32], a[32];
        gather_acc(t, 3, v);
        int b, y, x;
        const bool in = locate(t * 128 + tid, b, y, x);
#pragma unroll
        for (int ch = 0; ch < 32; ++ch) {
            xres[t][ch] = v[ch] + c_b0[ch];
            a[ch] = in ? fmaxf(fmaf(xres[t][ch], c_blk(0, 0)[ch], c_blk(0, 1)[ch]), 0.f) : 0.f;
        }
        store_act(t * 128 + tid, a);
    }
    tc_fence_before();

    // ---- two residual blocks ------------------------------------------------------------------------------------
#pragma unroll 1
    for (int blk = 0; blk < 2; ++blk) {
        run_layer(false, 4, 9, 0);  // conv1 (BN_b folded)
        ++stage;
        load_weights(stage_ptr(stage), stage_ptr(stage) + kTcStage, kTcStage);
#pragma unroll
        for (int t = 0; t < T; ++t) {
            float v[32], a[32];
            gather_acc(t, 3, v);
            int b, y, x;
            const bool in = locate(t * 128 + tid, b, y, x);
#pragma unroll
            for (int ch = 0; ch < 32; ++ch) a[ch] = in ? fmaxf(v[ch] + c_blk(blk, 2)[ch], 0.f) : 0.f;
            store_act(t * 128 + tid, a);
        }
        tc_fence_before();
        run_layer(false, 4, 9, 0);  // conv2
        ++stage;
        if (blk == 0) load_weights(stage_ptr(stage), stage_ptr(stage) + kTcStage, kTcStage);
        const float* sN = blk == 0 ? c_blk(1, 0) : c_sO;  // the BatchNorm that consumes the updated residual stream
        const float* tN = blk == 0 ? c_blk(1, 1) : c_tO;
#pragma unroll
        for (int t = 0; t < T; ++t) {
            float v[32], a[32];
            gather_acc(t, 3, v);
            int b, y, x;
            const bool in = locate(t * 128 + tid, b, y, x);
#pragma unroll
            for (int ch = 0; ch < 32; ++ch) {
                xres[t][ch] += v[ch] + c_blk(blk, 3)[ch];
                a[ch] = in ? fmaxf(fmaf(xres[t][ch], sN[ch], tN[ch]), 0.f) : 0.f;
            }
            store_act(t * 128 + tid, a);
        }
        tc_fence_before();
    }

    // ---- out block: conv1x1 32 -> Cout, 8 chunks of 32 output channels per weight stage ------------------------------
    for (int c0 = 0; c0 < TL.n_chunks; c0 += 8) {
        const int nc = (TL.n_chunks - c0) < 8 ? (TL.n_chunks - c0) : 8;
        load_weights(pk + TL.out_hi + c0 * 1024, pk + TL.out_lo + c0 * 1024, nc * 1024);
        for (int c = 0; c < nc; ++c) {
            run_layer(false, 4, 1, static_cast<uint32_t>(c * 256));  // 1024 floats = 256 x 16 B per chunk
#pragma unroll 1
            for (int t = 0; t < T; ++t) {
                float v[32];
                gather_acc(t, 1, v);
                int b, y, x;
                if (locate(t * 128 + tid, b, y, x)) {
#pragma unroll
                    for (int o = 0; o < 32; ++o) {
                        const int oc = (c0 + c) * 32 + o;
                        if (oc < Cout) out[((static_cast<size_t>(b) * Cout + oc) * H + y) * W + x] = v[o] + c_bout[oc];
                    }
                }
            }
            tc_fence_before();
        }
        __syncthreads();  // all MMAs of this weight stage are complete (mbar) before the next stage overwrites wH/wL
    }

    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(static_cast<uint32_t>(GM::TMEM_COLS)));
}

// ---- packing of the tensor-core section from the FFMA section -------------------------------------------------------
__global__ void __launch_bounds__(256) pack_tc_kernel(const float* __restrict__ pk, float* __restrict__ tc, int Cin, int Cout) {
    const PackLayout L = pack_layout(Cin, Cout, 9);
    const TcLayout T = tc_layout(Cin, Cout);
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < T.total; idx += gridDim.x * blockDim.x) {
        float val;
        bool lo = false;
        if (idx < T.stage0) {
            const int k = idx - T.consts;
            const int c = k & 31;
            if (k < 32) val = pk[L.b0 + c];
            else if (k < 288) {
                const int blk = (k - 32) / 128, which = ((k - 32) % 128) / 32;
                val = which == 0 ? pk[L.bnA[blk] + c] : which == 1 ? pk[L.bnA[blk] + kF + c]
                      : which == 2 ? pk[L.b1[blk] + c] : pk[L.b2[blk] + c];
            } else if (k < 320) val = pk[L.bnO + c];
            else if (k < 352) val = pk[L.bnO + kF + c];
            else val = pk[L.bout + (k - 352)];
            tc[idx] = val;
            continue;
        }
        if (idx < T.out_hi) {
            const int e0 = idx - T.stage0;
            const int s = e0 / (2 * kTcStage);
            int r = e0 - s * 2 * kTcStage;
            lo = r >= kTcStage;
            if (lo) r -= kTcStage;
            const int tap = r / 1024, r2 = r - tap * 1024;
            const int c4 = r2 / 128, n = (r2 % 128) / 4, q = r2 % 4;
            const int ci = 4 * c4 + q;
            if (s < T.n_in) {
                const int cig = s * kF + ci;
                val = cig < Cin ? pk[L.w0 + (cig * 9 + tap) * kF + n] : 0.f;
            } else {
                const int i = s - T.n_in;
                const int base = (i & 1) ? L.w2[i >> 1] : L.w1[i >> 1];
                val = pk[base + (ci * 9 + tap) * kF + n];
            }
        } else {
            int e0 = idx - T.out_hi;
            lo = e0 >= T.n_chunks * 1024;
            if (lo) e0 -= T.n_chunks * 1024;
            const int chunk = e0 / 1024, r2 = e0 - chunk * 1024;
            const int c4 = r2 / 128, n = (r2 % 128) / 4, q = r2 % 4;
            const int ci = 4 * c4 + q;
            val = pk[L.wout + (chunk * kF + ci) * kF + n];  // pack_wn layout: ((o>>5)*J + j)*32 + (o&31), J = 32
        }
        const float hi = __uint_as_float(__float_as_uint(val) & 0xFFFFE000u);
        tc[idx] = lo ? val - hi : hi;
    }
}

int pack_tc_launch(const float* pk_ffma, float* pk_tc_section, int Cin, int Cout, cudaStream_t st) {
    const TcLayout T = tc_layout(Cin, Cout);
    int blocks = (T.total + 255) / 256;
    if (blocks > kSMs * 8) blocks = kSMs * 8;
    pack_tc_kernel<<<blocks, 256, 0, st>>>(pk_ffma, pk_tc_section, Cin, Cout);
    return launch_status();
}

template <int H, int W, int S, int MODE>
static int launch_tc(const float* zsrc, float* out, const float* pk_tc, const SplitGeom& g, int Cin, int Cout, int B,
                     cudaStream_t st) {
    using GM = TcGeom<H, W, S>;
    auto kern = convnet_tc_kernel<H, W, S, MODE>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(GM::SMEM));
        attr_set = true;
    }
    kern<<<(B + S - 1) / S, 128, GM::SMEM, st>>>(zsrc, out, pk_tc, g, Cin, Cout, B, g_tune[4]);
    return launch_status();
}

template <int MODE>
static int tc_by_size(const float* zsrc, float* out, const float* pk_tc, const SplitGeom& g, int Cin, int Cout, int B, int h,
                      int w, cudaStream_t st) {
    if (Cout > 768) return NFB_ERR_UNSUPPORTED;  // constants buffer
    if (h == 16 && w == 16) return launch_tc<16, 16, 1, MODE>(zsrc, out, pk_tc, g, Cin, Cout, B, st);
    if (h == 8 && w == 8) return launch_tc<8, 8, 2, MODE>(zsrc, out, pk_tc, g, Cin, Cout, B, st);
    if (h == 4 && w == 4) return launch_tc<4, 4, 3, MODE>(zsrc, out, pk_tc, g, Cin, Cout, B, st);
    return NFB_ERR_UNSUPPORTED;
}

int convnet_tc_dispatch(const float* zsrc, float* out, const float* pk_tc, const SplitGeom& g, int mode, int Cin, int Cout,
                        int B, int h, int w, cudaStream_t st) {
    if (mode == NFB_SPLIT_CHECKER) return tc_by_size<NFB_SPLIT_CHECKER>(zsrc, out, pk_tc, g, Cin, Cout, B, h, w, st);
    if (mode == NFB_SPLIT_CHANNEL) return tc_by_size<NFB_SPLIT_CHANNEL>(zsrc, out, pk_tc, g, Cin, Cout, B, h, w, st);
    return tc_by_size<-1>(zsrc, out, pk_tc, g, Cin, Cout, B, h, w, st);
}

}  // namespace nfb
