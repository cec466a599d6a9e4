// conditioner_tc.cuh -- the ConvNet conditioner (modules.py:416-438, weight_norm.py:35-45) on the 5th-generation tensor
// cores, optionally fused with AffineCoupling._transform (coupling.py:104-112) so that (t, s) never leave the SM.
// (Kernel + launcher; included by conditioner_tc.cu and conditioner_tc_step.cu.)
//
// Arithmetic: every 3x3 / 1x1 layer is an implicit GEMM issued with tcgen05.mma (M = 128 positions) by ONE thread,
// accumulators in TMEM.  Single-pass TF32 / FP16 misses the 1e-5 bits/dim bar, so every product is
// error-compensated: x = x_hi + x_lo with 11 significant bits each (TF32: x_hi = tf32(x), x_lo = tf32(x - x_hi); the default
// FP16 split is described further down),
//     a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi            (the dropped a_lo*b_lo is ~2^-24 relative).
// The GEMM is skinny (N = 32 output channels) and in SS mode every MMA re-reads its 4 KB A tile from shared memory,
// so the operand B is N-CONCATENATED: rows [w_hi | w_lo] give a_hi*b_hi and a_hi*b_lo from ONE A read (N = 64), and
// a_lo*b_hi is a second MMA (N = 32) into the compensation columns: 2 A reads per k-step instead of 3.
//
// Implicit GEMM without padding or im2col ("flat shift + lane masks"): a tile is 128 consecutive positions of the flat
// (sample, y, x) sequence, stored K-major without swizzle as [ci/4][position][4 ci]; a 3x3 tap is a row offset
// (dy*W + dx) in the A descriptor's start address, and the rows whose tap falls outside the image are switched off
// with the instruction's disable-output-lane mask.  M utilisation is 100 % (16x16: 2 tiles per sample; 8x8: 2 samples
// per tile; 4x4: 8 samples per tile).
//
// Operand format (default, round 2b): the FP16 split -- x = hi + 2^-10 lo' with hi = fp16(x), lo' = fp16(2^10 (x - hi)); the
// same 11 + 11 significant bits and the same exact products as the TF32 split (the error against the fp64 oracle is the
// same or smaller), but 2 bytes per operand element and K = 16 per instruction: half the shared-memory operand traffic
// the issue loop was bound by, half the instructions, half the weight stages (36 KB), half the activation planes.
// kind::tf32 operands (NFB_CONV_TF32) remain for data outside the fp16 range (|x| >= 65504 gives NaN outputs here).
//
// Accumulation: the tensor core adds into the fp32 accumulator with truncation, so a long MMA chain drifts by
// ~(chain length) x 2^-24.  The k-steps of a layer are therefore spread over G accumulator groups (default 3: <= 13
// chained MMAs), each [main 32 | compensation 32] columns, started by an unmasked centre-tap MMA with accumulate = 0
// and summed by the epilogue with round-to-nearest adds.
//
// Roles (320 threads, one persistent CTA per SM, units = groups of samples round-robin over the grid):
//   warp 9   TMA producer: streams the weight stages (<= 72 KB each: [tap][k-step][w_hi|w_lo], the exact shared-memory
//            image, packed once per weight update) with cp.async.bulk into a 2-slot ring, mbarrier complete_tx.
//   warp 8   MMA issuer (one lane): waits weights-full + activations-ready, issues tcgen05.mma, tcgen05.commit ->
//            accumulators-done / weights-empty / halo-free mbarriers.
//   warps 0-7 epilogue: tcgen05.ld -> bias / BatchNorm / ReLU / residual (registers) -> hi/lo split -> st.shared into
//            the activation planes of the NEXT layer (in place) -> activations-ready; gather of z1 from z with the
//            coupling's split addressing; final layer: affine coupling on z0 (in place on z) + per-sample log-det, or
//            the plain params store.
// 16x16: the two tiles of a sample overlap (epilogue of one under the MMAs of the other); the in-place activation
// update is ordered by a "halo-free" commit after the second tile's dy = -1 taps.
#pragma once
#include <cstdio>
#include <utility>

#include "conditioner.cuh"

namespace nfb {

namespace {

template <int V>
using IC = std::integral_constant<int, V>;
template <class F, int... Is>
__device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, Is...>) {
    (f(IC<Is>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
    static_for_impl(f, std::make_integer_sequence<int, N>{});
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}
// site: who waits (for the time-out report): 1 producer, 2.. MMA lane, 30.. epilogue (which waits 2 s longer, so that
// the role that is actually stuck reports first)
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int site = 0) {
    // bounded (4 s of wall clock): a lost arrival must trap (cudaErrorLaunchFailure), never hang the GPU
    if (mbar_try(bar, parity)) return;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
#pragma unroll 1
        for (int it = 0; it < 64; ++it)
            if (mbar_try(bar, parity)) return;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > (site >= 30 ? 6000000000ull : 4000000000ull)) {
            printf("nfb200 convnet_tc_kernel: mbarrier %u (parity %u) timed out at site %d: block %d thread %d\n", bar, parity,
                   site, static_cast<int>(blockIdx.x), static_cast<int>(threadIdx.x));
            __trap();
        }
    }
}
// TMA (bulk async copy engine): global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// true in exactly one lane of the (converged) warp; the compiler treats the guarded region as single-lane uniform code
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
// D[tmem] (+)= A[smem] * B[smem]; rows whose bit is set in the 128-bit mask keep their old accumulator value
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                         uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(m0), "r"(m1), "r"(m2), "r"(m3)
        : "memory");
}
// same instruction with half-precision operands (K = 16 per instruction)
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                        uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(m0), "r"(m1), "r"(m2), "r"(m3)
        : "memory");
}
template <bool F16>
__device__ __forceinline__ void mma_any(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                        uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3) {
    if (F16) mma_f16(d_tmem, adesc, bdesc, idesc, accumulate, m0, m1, m2, m3);
    else mma_tf32(d_tmem, adesc, bdesc, idesc, accumulate, m0, m1, m2, m3);
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// N consecutive TMEM columns of this thread's lane -> registers (load + wait in one statement: the registers are
// defined when the statement retires)
template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[N]);
template <>
__device__ __forceinline__ void tmem_ld<4>(uint32_t taddr, float (&v)[4]) {
    uint32_t r[4];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n\t"
        "tcgen05.wait::ld.sync.aligned;\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
template <>
__device__ __forceinline__ void tmem_ld<8>(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n\t"
        "tcgen05.wait::ld.sync.aligned;\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
template <>
__device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
template <>
__device__ __forceinline__ void tmem_ld<32>(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
        "tcgen05.wait::ld.sync.aligned;\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// two 16-column loads in flight, one wait
__device__ __forceinline__ void tmem_ld16x2(uint32_t ta, uint32_t tb, float (&a)[16], float (&b)[16]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%32];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%33];\n\t"
        "tcgen05.wait::ld.sync.aligned;\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(ta), "r"(tb)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        a[i] = __uint_as_float(r[i]);
        b[i] = __uint_as_float(r[16 + i]);
    }
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

__device__ __forceinline__ float tf32_rn(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = tf32_rn(x);
    lo = tf32_rn(x - hi);
}

// FP16 split ("fp16x3"): x = hi + lo, hi = fp16_rn(x), lo' = fp16_rn((x - hi) * 2^kTcLoShift).  hi carries 11 significant
// bits like TF32; the residual x - hi is exact in fp32 and at most 2^-11 |x|, so scaled by 2^10 it sits in the normal fp16
// range for every |x| < 65504 (and for tiny x, where hi is a subnormal, the residual still holds the rest of x).  Products of
// two fp16 numbers are exact in the fp32 accumulator exactly like TF32 ones; the compensation columns accumulate
// 2^10 (a_hi w_lo + a_lo w_hi) and the epilogue scales them back.  Same error as 3xTF32 (oracle comparison in tests/), half
// the operand bytes and half the instructions per input channel.  Saturating conversions: |x| >= 65504 degrades the
// precision of that element, it never produces an infinity.
constexpr int kTcLoShift = 10;
constexpr float kLoScale = 1024.f, kLoInv = 1.f / 1024.f;
__device__ __forceinline__ void split_f16_pair(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));  // low half = x0 (the lower address)
    float h0, h1;
    asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}\n" : "=f"(h0), "=f"(h1) : "r"(hi));
    const float l0 = (x0 - h0) * kLoScale, l1 = (x1 - h1) * kLoScale;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(l1), "f"(l0));
}

// instruction descriptor: D = F32, A = B = TF32 (F16: half precision), both K-major, M = 128, N as given (mma_sm100_desc.hpp
// bit layout: c_format bits 4-5, a_format 7-9, b_format 10-12; F16 = 0, TF32 = 2)
template <bool F16 = false>
__host__ __device__ constexpr uint32_t idesc_n(uint32_t n) {
    return (1u << 4) | ((F16 ? 0u : 2u) << 7) | ((F16 ? 0u : 2u) << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

// bit l of word wd set <=> lane 32*wd + l of tile t is a position whose tap (dy, dx) falls outside its image
template <int H, int W>
__host__ __device__ constexpr uint32_t edge_mask(int t, int dy, int dx, int wd) {
    uint32_t m = 0;
    for (int l = 0; l < 32; ++l) {
        const int p = t * 128 + wd * 32 + l;
        const int pix = p % (H * W);
        const int y = pix / W, x = pix % W;
        const bool off = (dy < 0 && y == 0) || (dy > 0 && y == H - 1) || (dx < 0 && x == 0) || (dx > 0 && x == W - 1);
        if (off) m |= 1u << l;
    }
    return m;
}

// offset inside one sample (original layout of z) of channel m, pixel (i, j) of half `second` (0 = z0, 1 = z1)
template <int MODE>
__device__ __forceinline__ int half_elem_offset(const SplitGeom& g, int m, int second, int i, int j) {
    const bool outer = (second ^ g.odd) == 0;
    if (MODE == NFB_SPLIT_CHANNEL) return (m + (outer ? 0 : g.c0)) * g.HW + i * g.W + j;
    const int k = outer ? (m < g.C ? m : m + 2 * g.C) : m + g.C;
    return (k >> 2) * g.HW + (2 * i + ((k >> 1) & 1)) * g.W + 2 * j + (k & 1);
}

// developer timeline (builds with -DNFB_TC_TIMELINE only: `make EXTRA=-DNFB_TC_TIMELINE`; profiles/tc2_timeline.py): when set
// (nfb_debug_timeline), CTA 0 records clock64() stamps: [role 0 = epilogue thread 0, 1 = epilogue thread 128, 2 = MMA lane 0,
// 3 = kernel entry / prologue done / exit][event index] = (tag << 48) | (clock & 0xffffffffffff).  In the normal build the
// stamps compile to nothing: the MMA lane pays ~5 cycles for every instruction it executes, also for a not-taken branch.
static __device__ unsigned long long* g_tl_buf = nullptr;
constexpr int kTlEvents = 512;
struct Timeline {
#ifdef NFB_TC_TIMELINE
    unsigned long long* p;
    int n;
    __device__ __forceinline__ void init(int role, bool on) {
        p = (on && g_tl_buf) ? g_tl_buf + role * kTlEvents : nullptr;
        n = 0;
    }
    __device__ __forceinline__ void stamp(int tag) {
        if (p && n < kTlEvents) {
            p[n++] = (static_cast<unsigned long long>(tag) << 48) | (static_cast<unsigned long long>(clock64()) & 0xffffffffffffull);
        }
    }
#else
    __device__ __forceinline__ void init(int, bool) {}
    __device__ __forceinline__ void stamp(int) {}
#endif
};

// tap issue order: centre first (its first Ge k-steps start the accumulator groups with accumulate = 0, unmasked);
// tile 0 of a linked pair: dy = +1 taps last (they read tile 1's rows); tile 1: dy = -1 taps right after the centre
// (they read tile 0's rows, which tile 0's epilogue overwrites once they are done)
constexpr uint64_t kOrder0 = 0x876210534ull, kOrder1 = 0x876532104ull;
constexpr int kSlotBytes = 9 * 4 * 2048;  // one 32->32 3x3 stage: [tap][k-step][2 x (64 rows x 16 B)] (TF32; FP16 split: half)
__host__ __device__ constexpr int slot_bytes(bool f16) { return f16 ? kSlotBytes / 2 : kSlotBytes; }
constexpr int kEpiThreads = 256;
constexpr int kThreads = 320;
constexpr int kTileCols = 256;            // TMEM columns reserved per tile

// PAIR (maps of <= 128 pixels only): a unit is TWO independent tiles that take turns on the tensor core exactly like the
// two tiles of a 16x16 sample (epilogue of one under the MMAs of the other, weights streamed once for both): ~1.5x the
// work per SM-second of the single-tile unit, which leaves the tensor core idle during every epilogue, at the price of
// half as many CTAs -- the choice when the batch (or several batches in flight) fills the machine anyway.
// DUAL (16x16 maps, FP16 split): TWO units (samples) are in flight per CTA -- separate activation planes and accumulators,
// one shared weight ring.  The MMA lane issues [A tile 0, A tile 1, B tile 0, B tile 1] per layer, each epilogue warp quad
// serves "its" tile of A and then of B, so the tensor pipe always has the other sample's layer to run while one sample is
// in its epilogue, and the serial head and tail of a sample (gather, input layer, output layer, coupling) hide behind the
// other sample's 3x3 layers.  The residual stream of the 4 tiles lives in shared memory instead of registers.
template <int H, int W, bool PAIR, bool F16 = false, bool DUAL = false>
struct TcGeom {
    static constexpr int HW = H * W;
    static constexpr bool LINKED = HW > 128 || PAIR;    // two tiles per unit, ping-pong (16x16: the two halves of a sample)
    static constexpr int T = LINKED ? 2 : 1;            // tiles per unit
    static constexpr int SPT = HW > 128 ? 0 : 128 / HW; // samples per tile (0: a sample spans both tiles)
    static constexpr int SPU = HW > 128 ? 1 : T * SPT;  // samples per unit
    static constexpr int CS = LINKED ? 1 : 2;           // epilogue warps per TMEM lane quarter of one tile
    static constexpr int NCH = 32 / CS;                 // channels per epilogue thread
    static constexpr int GUARD = W + 1;                 // positions before / after the tiles (tap offsets reach there)
    static constexpr int PB = 2 * GUARD + T * 128;      // positions per channel-chunk plane
    static constexpr int PS = PB * 16;                  // bytes per plane
    static constexpr int NPL = F16 ? 4 : 8;             // planes of 16 B per position: 4 TF32 / 8 FP16 channels each
    static constexpr int NJ = F16 ? 2 : 4;              // k-steps per tap of a 32-channel layer
    static constexpr int ACT_BYTES = 2 * NPL * PS;      // hi planes + lo planes
    static constexpr int NU = DUAL ? 2 : 1;             // units in flight per CTA
    static constexpr int TC = DUAL ? 128 : kTileCols;   // TMEM columns per tile (DUAL: 2 accumulator groups, out layer <= 128)
    static constexpr int XS_BYTES = DUAL ? NU * T * 32 * 128 * 4 : 0;  // residual stream [unit][tile][channel][row]
    static_assert(T <= 2 && NU * T * TC <= 512, "units exceed TMEM");
    static_assert(!DUAL || (LINKED && F16), "DUAL: two-tile units with FP16-split operands");
    static_assert(HW == 256 || 128 % HW == 0, "tile must hold whole samples");
    static_assert(!(PAIR && HW > 128), "PAIR is for maps of at most 128 pixels");
};

}  // namespace

// =====================================================================================================================
// the kernel
// =====================================================================================================================
// The NEXT flow step's ActNorm + invertible 1x1 convolution (modules.py:246-250, 470-480), run by the same CTA on the
// samples it has just finished (CP = channels of z, compile time: the per-pixel matrix-vector product lives in registers).
struct PostOp {
    const float* an_log_scale;  // (C)
    const float* an_bias;       // (C)
    const float* W;             // (C, C) row-major, from nfb_invconv1x1_weight
    const float* log_s;         // (C)
};

template <int H, int W, int MODE, bool FUSED, bool PAIR, int CP, bool F16, bool DUAL>
__global__ void __launch_bounds__(kThreads, 1)
convnet_tc_kernel(const float* zsrc, float* zdst, float* ldj, const float* __restrict__ pk, SplitGeom g, int Cin, int Cout,
                  int B, const float* __restrict__ p_sa, const float* __restrict__ p_sb, PostOp post, int G, int dbg_arg) {
    // profiling knobs (NFB_CONV_DEBUG: skip MMAs / TMEM loads / activation stores / weight traffic -- WRONG results) exist in
    // builds with -DNFB_TC_DEBUG_KNOBS only; in the normal build they fold to constants (the MMA lane pays for every branch)
#ifdef NFB_TC_DEBUG_KNOBS
    const int dbg = dbg_arg;
#else
    constexpr int dbg = 0;
    (void)dbg_arg;
#endif
    using GM = TcGeom<H, W, PAIR, F16, DUAL>;
    constexpr int HW = GM::HW, T = GM::T, SPT = GM::SPT, SPU = GM::SPU, CS = GM::CS, NCH = GM::NCH, GUARD = GM::GUARD,
                  PB = GM::PB, PS = GM::PS, NPL = GM::NPL, NJ = GM::NJ, NU = GM::NU, TC = GM::TC;
    constexpr bool LINKED = GM::LINKED;
    constexpr int SLOT = slot_bytes(F16);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* act = smem_raw;                              // [NU units][2 NPL planes][PB][16 B]
    unsigned char* ring = smem_raw + NU * GM::ACT_BYTES;        // 2 x SLOT
    float* xs = reinterpret_cast<float*>(ring + 2 * SLOT);      // DUAL: residual stream [unit][tile][32 channels][128 rows]
    float* cst = reinterpret_cast<float*>(ring + 2 * SLOT + GM::XS_BYTES);

    const TcPlan P = tc_plan(Cin, Cout, F16 ? 1 : 0);
    const int n_cst = 352 + (FUSED ? P.nqf * P.NWf : P.nqg * P.NWg);
    uint64_t* bars = reinterpret_cast<uint64_t*>(cst + ((n_cst + 3) & ~3));
    // barrier indices
    // ([unit][tile] for ACT_READY / ACC_DONE; an odd count keeps the 16-byte alignment of the tables behind the barriers)
    constexpr int W_FULL = 0, W_EMPTY = 2, ACT_READY = 4, ACC_DONE = 4 + 2 * NU, WAR0 = 4 + 4 * NU, N_BARS = (4 + 5 * NU) | 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + N_BARS);
    float* red_all = reinterpret_cast<float*>(tmem_slot + 2);   // [NU units][8 warps][2]
    uint4* mask_tab = reinterpret_cast<uint4*>(red_all + 16 * NU);        // [T][9 taps]: rows of the tile whose tap leaves the image
    constexpr int ROWW = 1 + NJ;                                 // uint4 per row of the issue program
    constexpr int N_ROWS = 2 * NU * T * 9;                       // [weight slot][unit][tile][tap in issue order]
    uint4* prog = mask_tab + 2 * 9;                              // N_ROWS x [lane mask, NJ k-steps]: see below
    uint4* prog_in = prog + N_ROWS * ROWW;                       // the same for the last pass of the input layer (nj_last k-steps)
    // CP > 0: Wt[ci][co], then exp(log_scale)[c], bias[c], 2 sums
    float* wpost = reinterpret_cast<float*>(prog_in + N_ROWS * ROWW);
    const uint32_t bar0 = smem_u32(bars);
    auto bar = [&](int i) -> uint32_t { return bar0 + 8u * static_cast<uint32_t>(i); };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    Timeline tl_k;  // kernel entry / prologue done / exit (globaltimer ns in the low bits for tags >= 60)
    tl_k.init(3, blockIdx.x == 0 && tid == 0);
    tl_k.stamp(0);
    const int NW = FUSED ? P.NWf : P.NWg, nq = FUSED ? P.nqf : P.nqg;
    const int out_chunk_bytes = 64 * NJ * NW;
    const int qps = SLOT / out_chunk_bytes;                     // out chunks per weight stage
    const int n_out_stage = (nq + qps - 1) / qps;
    const int n_stage = P.n_in + 4 + n_out_stage;
    const int n_units = ((B + SPU - 1) / SPU + NU - 1) / NU;    // loop iterations: NU units each

    // ---- prologue: barriers, constants, TMEM ----------------------------------------------------------------------
    if (tid == 0) {
        mbar_init(bar(W_FULL), 1); mbar_init(bar(W_FULL + 1), 1);
        mbar_init(bar(W_EMPTY), 1); mbar_init(bar(W_EMPTY + 1), 1);
        for (int i = 0; i < 2 * NU; ++i) { mbar_init(bar(ACT_READY + i), 128 * CS); mbar_init(bar(ACC_DONE + i), 1); }
        for (int i = 0; i < NU; ++i) mbar_init(bar(WAR0 + i), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < T * 9 * 4) {
        const int t = tid / 36, tap = (tid % 36) / 4, wd = tid % 4;
        reinterpret_cast<uint32_t*>(mask_tab)[tid] = edge_mask<H, W>(t, tap / 3 - 1, tap % 3 - 1, wd);
    }
    if (CP > 0) {
        for (int i = tid; i < CP * CP; i += kThreads) {
            const int ci = i / CP, co = i - ci * CP;
            wpost[i] = __ldg(post.W + co * CP + ci);  // transposed Wt[ci][co]: consecutive outputs of one input are contiguous
        }
        if (tid < CP) {
            wpost[CP * CP + tid] = expf(__ldg(post.an_log_scale + tid));
            wpost[CP * CP + CP + tid] = __ldg(post.an_bias + tid);
        }
        if (tid == 0) {
            float an = 0.f, cv = 0.f;
            for (int c = 0; c < CP; ++c) { an -= __ldg(post.an_log_scale + c); cv += __ldg(post.log_s + c); }
            wpost[CP * CP + 2 * CP] = an;
            wpost[CP * CP + 2 * CP + 1] = cv;
        }
    }
    {
        const float* src = pk + P.consts;
        for (int i = tid; i < 352; i += kThreads) cst[i] = __ldg(src + i);
        const float* ob = pk + (FUSED ? P.obias_f : P.obias_g);
        for (int i = tid; i < nq * NW; i += kThreads) cst[352 + i] = __ldg(ob + i);
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    tl_k.stamp(1);

    if (warp == 9) {
        // =============================== TMA producer ================================================================
        if (elect_one()) {
            const unsigned char* pkb = reinterpret_cast<const unsigned char*>(pk);
            uint32_t cnt = 0;
            for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
                for (int s = 0; s < n_stage; ++s, ++cnt) {
                    const uint32_t slot = cnt & 1u;
                    mbar_wait(bar(W_EMPTY + slot), ((cnt >> 1) & 1u) ^ 1u, 1);
                    size_t off;
                    uint32_t bytes;
                    if (s < P.n_in) {
                        off = static_cast<size_t>(P.in0) * 4 + static_cast<size_t>(s) * SLOT;
                        bytes = (s == P.n_in - 1) ? 9u * P.nj_last * 2048u : static_cast<uint32_t>(SLOT);
                    } else if (s < P.n_in + 4) {
                        off = static_cast<size_t>(P.mid0) * 4 + static_cast<size_t>(s - P.n_in) * SLOT;
                        bytes = SLOT;
                    } else {
                        const int q0 = (s - P.n_in - 4) * qps;
                        const int nqs = (nq - q0) < qps ? (nq - q0) : qps;
                        off = static_cast<size_t>(FUSED ? P.outf : P.outg) * 4 + static_cast<size_t>(q0) * out_chunk_bytes;
                        bytes = static_cast<uint32_t>(nqs * out_chunk_bytes);
                    }
                    if (dbg & 32) { mbar_arrive(bar(W_FULL + slot)); continue; }  // profiling knob: no weight traffic
                    mbar_expect_tx(bar(W_FULL + slot), bytes);
                    const uint32_t dst = smem_u32(ring + slot * SLOT);
                    for (uint32_t o = 0; o < bytes; o += 18432u) {
                        const uint32_t n = (bytes - o) < 18432u ? (bytes - o) : 18432u;
                        bulk_g2s(dst + o, pkb + off + o, n, bar(W_FULL + slot));
                    }
                }
            }
        }
    } else if (warp == 8) {
        // =============================== MMA issuer ==================================================================
        // One lane chosen with elect.sync runs the whole role: the compiler then keeps descriptors, masks and counters in
        // uniform registers and emits back-to-back UTCHMMA.  (Measured: `if (lane == 0)` or a predicated asm wraps every
        // MMA in an ELECT loop with R2UR moves, ~200 cycles per k-step; fully unrolled issue code is instruction-fetch
        // bound, ~380 cycles per k-step.)
        // Issue program of a 32->32 3x3 layer: the operands of its k-steps are the same in every such layer, so the complete
        // low words of the three descriptors and the accumulator address are tabulated once per (weight slot, unit, tile,
        // tap in issue order): [lane mask | NJ x (A_hi descriptor, A_lo descriptor, B descriptor, TMEM address)].  The issuing
        // lane is a single thread whose instructions cost their full latency (~5 cycles each, measured): with these rows a
        // tap is 1 + NJ shared-memory loads, the register -> uniform-register moves and its MMAs, nothing else.
        // A second table holds the last pass of the input layer (nj_last <= NJ k-steps per tap, its own weight offsets).
        for (int r = lane; r < 2 * N_ROWS; r += 32) {
            const bool in_layer = r >= N_ROWS;
            const int rr = in_layer ? r - N_ROWS : r;
            const int nj = in_layer ? P.nj_last : NJ;
            const int Ge = G < nj ? G : nj;
            const int i = rr % 9, t = (rr / 9) % T, u = (rr / (9 * T)) % NU, slot = rr / (9 * T * NU);
            const uint64_t order = (LINKED && t == 1) ? kOrder1 : kOrder0;
            const int tap = static_cast<int>((order >> (4 * i)) & 15u);
            const int dy = tap / 3 - 1, dx = tap % 3 - 1;
            uint4* row = (in_layer ? prog_in : prog) + rr * ROWW;
            row[0] = mask_tab[t * 9 + tap];
            for (int j = 0; j < NJ; ++j) {
                const int kc = nj * i + j;
                const uint32_t a_off = static_cast<uint32_t>(u * (GM::ACT_BYTES >> 4) + 2 * j * PB + GUARD + t * 128 + dy * W + dx);
                uint4 e;
                e.x = (((smem_u32(act) >> 4) + a_off) & 0x3FFFu) | (static_cast<uint32_t>(PS >> 4) << 16);
                e.y = (((smem_u32(act + NPL * PS) >> 4) + a_off) & 0x3FFFu) | (static_cast<uint32_t>(PS >> 4) << 16);
                e.z = (((smem_u32(ring + slot * SLOT) >> 4) + static_cast<uint32_t>((tap * nj + j) * 128)) & 0x3FFFu) | ((1024u >> 4) << 16);
                e.w = tmem + static_cast<uint32_t>(u * T * TC + t * TC + (kc % Ge) * 64);
                row[1 + j] = e;
            }
        }
        __syncwarp();
        if (elect_one()) {
            const uint64_t dAH = make_smem_desc(smem_u32(act), PS, 128);
            const uint64_t dAL = make_smem_desc(smem_u32(act + NPL * PS), PS, 128);
            const bool no_mma = (dbg & 1) != 0, no_lo = (dbg & 4) != 0;
            uint32_t cnt = 0;
            uint32_t ph_act = 0;  // bit (2u + t): phase of ACT_READY[u][t]
            Timeline tl;
            tl.init(2, blockIdx.x == 0);
            auto wait_act = [&](int u, int t) {
                tl.stamp(10 + t);
                const int i = 2 * u + t;
                mbar_wait(bar(ACT_READY + i), (ph_act >> i) & 1u, 10 + i);
                ph_act ^= 1u << i;
                tc_fence_after();
                tl.stamp(12 + t);
            };
            auto commit = [&](int b) {
                mma_commit(bar(b));
                tl.stamp(20 + b);
            };
            for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
                for (int s = 0; s < n_stage; ++s, ++cnt) {
                    const uint32_t slot = cnt & 1u;
                    tl.stamp(1);
                    mbar_wait(bar(W_FULL + slot), (cnt >> 1) & 1u, 2);
                    tc_fence_after();
                    tl.stamp(2);
                    const uint32_t wbase = smem_u32(ring + slot * SLOT);
                    if (s < P.n_in + 4) {
                        // ---- 3x3 layer ----
                        const int nj = (s == P.n_in - 1) ? P.nj_last : NJ;
                        const int Ge = G < nj ? G : nj;
                        const uint64_t dB = make_smem_desc(wbase, 1024, 128);
                        int kc = 0, rr = 0;  // k-steps issued into this tile's accumulators; round-robin group
                        uint32_t uA = 0, uT = 0;  // current unit: offset of its activation planes (16-byte units) / TMEM columns
                        int cur_u = 0;
                        // one k-step = 8 input channels of one tap: [main | comp] += a_hi * [w_hi | w_lo]; comp += a_lo * w_hi
                        auto kstep = [&](uint32_t tbase, uint64_t a_hi, uint64_t a_lo, uint64_t b, const uint4& m) {
                            const uint32_t d = tbase + static_cast<uint32_t>(rr * 64);
                            if (!no_mma) {
                                mma_any<F16>(d, a_hi, b, idesc_n<F16>(64), kc >= Ge ? 1u : 0u, m.x, m.y, m.z, m.w);
                                if (!no_lo) mma_any<F16>(d + 32, a_lo, b, idesc_n<F16>(32), 1u, m.x, m.y, m.z, m.w);
                            }
                            ++kc;
                            rr = (rr + 1 == Ge) ? 0 : rr + 1;
                        };
                        // taps order[i_lo .. i_hi) of tile t
                        auto issue = [&](int t, uint64_t order, int i_lo, int i_hi) {
                            if (!no_lo && (nj == NJ || s == P.n_in - 1)) {
                                // tabulated operands (see the prologue of this warp)
                                if (no_mma) return;
                                constexpr uint32_t hi32 = (128u >> 4) | (1u << 14);  // SBO = 128 B, descriptor version 1
                                const bool full = nj == NJ;
                                const uint4* pt = (full ? prog : prog_in) + (((slot * NU + cur_u) * T + t) * 9 + i_lo) * ROWW;
                                auto tap_row = [&](const uint4* row, bool first, bool all) {
                                    const uint4 m = row[0];
                                    uint4 e[NJ];
#pragma unroll
                                    for (int j = 0; j < NJ; ++j)
                                        if (all || j < nj) e[j] = row[1 + j];
#pragma unroll
                                    for (int j = 0; j < NJ; ++j) {
                                        if (all || j < nj) {
                                            const uint64_t b = (static_cast<uint64_t>(hi32) << 32) | e[j].z;
                                            // the centre tap comes first: its first Ge k-steps start the accumulator groups
                                            const uint32_t acc = (first && j < Ge) ? 0u : 1u;
                                            mma_any<F16>(e[j].w, (static_cast<uint64_t>(hi32) << 32) | e[j].x, b, idesc_n<F16>(64), acc, m.x, m.y, m.z, m.w);
                                            mma_any<F16>(e[j].w + 32, (static_cast<uint64_t>(hi32) << 32) | e[j].y, b, idesc_n<F16>(32), 1u, m.x, m.y, m.z, m.w);
                                        }
                                    }
                                };
                                int i = i_lo;
                                if (full) {
                                    if (i == 0) { tap_row(pt, true, true); ++i; pt += ROWW; }
#pragma unroll 1
                                    for (; i < i_hi; ++i, pt += ROWW) tap_row(pt, false, true);
                                } else {
                                    if (i == 0) { tap_row(pt, true, false); ++i; pt += ROWW; }
#pragma unroll 1
                                    for (; i < i_hi; ++i, pt += ROWW) tap_row(pt, false, false);
                                }
                                return;
                            }
                            const uint32_t tbase = tmem + uT + static_cast<uint32_t>(t * TC);
#pragma unroll 1
                            for (int i = i_lo; i < i_hi; ++i) {
                                const int tap = static_cast<int>((order >> (4 * i)) & 15u);
                                const int ty = (tap * 11) >> 5, dy = ty - 1, dx = tap - 3 * ty - 1;
                                const uint4 m = mask_tab[t * 9 + tap];
                                const uint32_t a_off = uA + static_cast<uint32_t>(GUARD + t * 128 + dy * W + dx);
                                const uint64_t a_hi = dAH + a_off, a_lo = dAL + a_off, b = dB + static_cast<uint32_t>(tap * nj * 128);
#pragma unroll 1
                                for (int j = 0; j < nj; ++j) kstep(tbase, a_hi + 2 * j * PB, a_lo + 2 * j * PB, b + j * 128, m);
                            }
                        };
#pragma unroll 1
                        for (int u = 0; u < NU; ++u) {
                            uA = static_cast<uint32_t>(u * (GM::ACT_BYTES >> 4));
                            uT = static_cast<uint32_t>(u * T * TC);
                            cur_u = u;
                            kc = 0; rr = 0;
                            wait_act(u, 0);
                            if (LINKED) {
                                issue(0, kOrder0, 0, 6);
                                wait_act(u, 1);
                                issue(0, kOrder0, 6, 9);
                                commit(ACC_DONE + 2 * u);
                                kc = 0; rr = 0;
                                issue(1, kOrder1, 0, 4);
                                commit(WAR0 + u);  // tile 0's rows are no longer read: its epilogue may overwrite them
                                issue(1, kOrder1, 4, 9);
                                commit(ACC_DONE + 2 * u + 1);
                            } else {
                                issue(0, kOrder0, 0, 9);
                                commit(ACC_DONE + 2 * u);
                            }
                        }
                    } else {
                        // ---- 1x1 output layer, chunks of NW columns: [main NW | comp NW] ----
                        const int q0 = (s - P.n_in - 4) * qps;
                        const int nqs = (nq - q0) < qps ? (nq - q0) : qps;
                        const uint32_t i_main = idesc_n<F16>(static_cast<uint32_t>(2 * NW)), i_comp = idesc_n<F16>(static_cast<uint32_t>(NW));
#pragma unroll 1
                        for (int qi = 0; qi < nqs; ++qi) {
                            const uint64_t dB = make_smem_desc(wbase + static_cast<uint32_t>(qi * out_chunk_bytes),
                                                               static_cast<uint32_t>(2 * NW * 16), 128);
#pragma unroll 1
                            for (int ut = 0; ut < NU * T; ++ut) {
                                const int u = ut / T, t = ut % T;
                                wait_act(u, t);
                                const uint32_t d = tmem + static_cast<uint32_t>(u * T * TC + t * TC);
#pragma unroll
                                for (int j = 0; j < NJ; ++j) {
                                    const uint32_t a_off = static_cast<uint32_t>(u * (GM::ACT_BYTES >> 4) + 2 * j * PB + GUARD + t * 128);
                                    const uint32_t b_off = static_cast<uint32_t>(j * 4 * NW);  // 2 blocks of 2NW rows x 16 B
                                    if (!no_mma) {
                                        mma_any<F16>(d, dAH + a_off, dB + b_off, i_main, j > 0 ? 1u : 0u, 0u, 0u, 0u, 0u);
                                        if (!no_lo) mma_any<F16>(d + NW, dAL + a_off, dB + b_off, i_comp, 1u, 0u, 0u, 0u, 0u);
                                    }
                                }
                                commit(ACC_DONE + 2 * u + t);
                            }
                        }
                    }
                    commit(W_EMPTY + slot);
                }
            }
        }
    } else {
        // =============================== epilogue warps ===============================================================
        const int q4 = warp & 3, grp = warp >> 2;
        const int tile = LINKED ? grp : 0;
        const int ch0 = LINKED ? 0 : grp * NCH;          // first of this thread's NCH channels
        const int row = q4 * 32 + lane;                  // TMEM lane = position inside the tile
        const int pos = tile * 128 + row;                // position inside the unit
        const int pix = HW > 128 ? pos : row % HW;
        const int yy = pix / W, xx = pix % W;
        const uint32_t t_lane0 = tmem + (static_cast<uint32_t>(q4 * 32) << 16) + static_cast<uint32_t>(tile * TC);
        unsigned char* my_act0 = act + (GUARD + pos) * 16;
        uint32_t ph_acc = 0, ph_war = 0;  // bit u: phase of ACC_DONE[u][tile] / WAR0[u]
        Timeline tl;
        tl.init(grp, blockIdx.x == 0 && (tid & 127) == 0);

        // ---- the unit this thread is working on (DUAL: the stages below alternate between the two units in flight) --------
        int cu = 0, unit = 0, b = 0;
        bool valid = false;
        const float* zb = zsrc;
        uint32_t t_lane = t_lane0, b_acc = bar(ACC_DONE + tile), b_act = bar(ACT_READY + tile), b_war = bar(WAR0);
        unsigned char* my_act = my_act0;
        float* xs_u = xs;  // DUAL: this thread's column of the residual stream, [channel][128 rows]
        // FP16 split: largest |activation| this thread has converted for the unit.  At >= 65504 the conversion saturates and
        // the sums that consumed it are wrong: the thread then returns NaN for its outputs of that unit (a loud failure; such
        // data needs NFB_CONV_TF32), see the output layer.
        float amax = 0.f, amax_u0 = 0.f, amax_u1 = 0.f;
        auto set_unit = [&](int u) {
            cu = u;
            const int un = unit * NU + u;
            b = HW > 128 ? un : un * SPU + tile * SPT + row / HW;
            valid = b < B;
            zb = zsrc + static_cast<size_t>(b) * (MODE < 0 ? static_cast<size_t>(Cin) * HW : static_cast<size_t>(g.D));
            t_lane = t_lane0 + static_cast<uint32_t>(u * T * TC);
            my_act = my_act0 + u * GM::ACT_BYTES;
            b_acc = bar(ACC_DONE + 2 * u + tile);
            b_act = bar(ACT_READY + 2 * u + tile);
            b_war = bar(WAR0 + u);
            if (DUAL) xs_u = xs + ((u * T + tile) * 32) * 128 + row;
            amax = u ? amax_u1 : amax_u0;
        };
        auto end_unit = [&]() {
            if (cu) amax_u1 = amax;
            else amax_u0 = amax;
        };
        auto wait_acc = [&]() {
            tl.stamp(30);
            mbar_wait(b_acc, (ph_acc >> cu) & 1u, 30 + cu);
            ph_acc ^= 1u << cu;
            tc_fence_after();
            tl.stamp(31);
        };
        auto wait_war = [&]() {
            if (LINKED && grp == 0) {
                tl.stamp(32);
                mbar_wait(b_war, (ph_war >> cu) & 1u, 32 + cu);
                ph_war ^= 1u << cu;
                tl.stamp(33);
            }
        };
        auto signal_act = [&]() {
            tl.stamp(34);
            fence_proxy_async();  // generic-proxy st.shared -> visible to the tensor core's async-proxy reads
            tc_fence_before();
            mbar_arrive(b_act);
            tl.stamp(35);
        };
        // The epilogue walks its NCH channels in sub-passes of EC = 16 (one for a column-split tile, two for a linked tile):
        // 32 channels at once need ~130 live registers next to the 32 of the residual stream and spill (measured).
        constexpr int EC = 16, NP = NCH / EC;
        // v[c] = sum over the Ge accumulator groups of (main + compensation) for channels ch0 + sub*EC ... + EC - 1
        auto load_acc = [&](int Ge, int sub, float (&v)[EC]) {
            if (dbg & 2) {  // profiling knob: no TMEM reads
#pragma unroll
                for (int i = 0; i < EC; ++i) v[i] = 0.f;
                return;
            }
            const uint32_t col = static_cast<uint32_t>(ch0 + sub * EC);
#pragma unroll 1
            for (int gi = 0; gi < Ge; ++gi) {
                float m[EC], cp[EC];
                tmem_ld16x2(t_lane + static_cast<uint32_t>(gi * 64) + col, t_lane + static_cast<uint32_t>(gi * 64 + 32) + col, m, cp);
                if (gi == 0) {
#pragma unroll
                    for (int i = 0; i < EC; ++i) v[i] = F16 ? fmaf(cp[i], kLoInv, m[i]) : m[i] + cp[i];
                } else {
#pragma unroll
                    for (int i = 0; i < EC; ++i) v[i] += F16 ? fmaf(cp[i], kLoInv, m[i]) : m[i] + cp[i];
                }
            }
            tl.stamp(36);
        };
        // a[EC] (activated) -> hi / lo planes of this thread's position
        auto store_act = [&](int sub, const float (&a)[EC]) {
            if (dbg & 16) return;  // profiling knob: no activation stores
            if (F16) {
#pragma unroll
                for (int i = 0; i < EC; ++i) amax = fmaxf(amax, fabsf(a[i]));
#pragma unroll
                for (int c8 = 0; c8 < EC / 8; ++c8) {
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) split_f16_pair(a[c8 * 8 + 2 * q], a[c8 * 8 + 2 * q + 1], hi[q], lo[q]);
                    const int plane = (ch0 + sub * EC) / 8 + c8;
                    *reinterpret_cast<uint4*>(my_act + plane * PS) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint4*>(my_act + (NPL + plane) * PS) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
                return;
            }
#pragma unroll
            for (int c4 = 0; c4 < EC / 4; ++c4) {
                float hi[4], lo[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) split_tf32(a[c4 * 4 + q], hi[q], lo[q]);
                const int plane = (ch0 + sub * EC) / 4 + c4;
                st4(reinterpret_cast<float*>(my_act + plane * PS), make_float4(hi[0], hi[1], hi[2], hi[3]));
                st4(reinterpret_cast<float*>(my_act + (8 + plane) * PS), make_float4(lo[0], lo[1], lo[2], lo[3]));
            }
        };

        // residual stream of this thread's position: registers, or (DUAL) shared memory
        float xres[DUAL ? 1 : NCH];
        auto xr_load = [&](int sub, float (&x)[EC]) {
#pragma unroll
            for (int i = 0; i < EC; ++i) x[i] = DUAL ? xs_u[(ch0 + sub * EC + i) * 128] : xres[DUAL ? 0 : sub * EC + i];
        };
        auto xr_store = [&](int sub, const float (&x)[EC]) {
#pragma unroll
            for (int i = 0; i < EC; ++i) {
                if (DUAL) xs_u[(ch0 + sub * EC + i) * 128] = x[i];
                else xres[DUAL ? 0 : sub * EC + i] = x[i];
            }
        };

        const float* c_b0 = cst;
        auto c_blk = [&](int blk, int k) { return cst + 32 + blk * 128 + k * 32; };
        const float* c_sO = cst + 288;
        const float* c_tO = cst + 320;
        const float* c_ob = cst + 352;
        float sa = 0.f, sb = 0.f;
        if (FUSED) { sa = __ldg(p_sa); sb = __ldg(p_sb); }

        for (unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
            tl.stamp(40);
            amax_u0 = amax_u1 = 0.f;

            // ---- in conv: Cin -> 32 in passes of <= 32 input channels; the partial sums meet in the residual stream ----
#pragma unroll 1
            for (int c = 0; c < P.n_in; ++c) {
                const int CI = (Cin - c * kF) < kF ? (Cin - c * kF) : kF;
                const int n4 = F16 ? ((CI + 15) & ~15) / 4 : ((CI + 7) & ~7) / 4;
#pragma unroll 1
                for (int u = 0; u < NU; ++u) {
                set_unit(u);
                // this thread's share of the chunk: all loads first (they are independent: one round trip to L2), then the
                // hi/lo split and the stores
                constexpr int NG = 8 / CS;  // channel groups of 4 per thread
                float gv[NG][4];
#pragma unroll
                for (int k = 0; k < NG; ++k) {
                    const int c4 = k * CS + (CS == 2 ? grp : 0);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int ci = c4 * 4 + q;
                        float v = 0.f;
                        if (valid && ci < CI) {
                            const int cg = c * kF + ci;
                            if (MODE < 0) v = __ldg(zb + cg * HW + pix);
                            else v = __ldg(zb + half_elem_offset<(MODE < 0 ? NFB_SPLIT_CHANNEL : MODE)>(g, cg, 1, yy, xx));
                        }
                        gv[k][q] = v;
                    }
                }
#pragma unroll
                for (int k = 0; k < NG; ++k) {
                    const int c4 = k * CS + (CS == 2 ? grp : 0);
                    if (c4 < n4 && F16) {  // 4 channels = half of a 16-byte row of plane c4 / 2
                        amax = fmaxf(fmaxf(amax, fmaxf(fabsf(gv[k][0]), fabsf(gv[k][1]))), fmaxf(fabsf(gv[k][2]), fabsf(gv[k][3])));
                        uint32_t hi[2], lo[2];
                        split_f16_pair(gv[k][0], gv[k][1], hi[0], lo[0]);
                        split_f16_pair(gv[k][2], gv[k][3], hi[1], lo[1]);
                        unsigned char* dst = my_act + (c4 >> 1) * PS + (c4 & 1) * 8;
                        *reinterpret_cast<uint2*>(dst) = make_uint2(hi[0], hi[1]);
                        *reinterpret_cast<uint2*>(dst + NPL * PS) = make_uint2(lo[0], lo[1]);
                    } else if (c4 < n4) {
                        float hi[4], lo[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) split_tf32(gv[k][q], hi[q], lo[q]);
                        st4(reinterpret_cast<float*>(my_act + c4 * PS), make_float4(hi[0], hi[1], hi[2], hi[3]));
                        st4(reinterpret_cast<float*>(my_act + (8 + c4) * PS), make_float4(lo[0], lo[1], lo[2], lo[3]));
                    }
                }
                signal_act();
                end_unit();
                }
#pragma unroll 1
                for (int u = 0; u < NU; ++u) {
                set_unit(u);
                wait_acc();
                const int nj = (c == P.n_in - 1) ? P.nj_last : NJ;
#pragma unroll
                for (int sub = 0; sub < NP; ++sub) {
                    float v[EC];
                    load_acc(G < nj ? G : nj, sub, v);
                    if (c > 0) {
                        float x[EC];
                        xr_load(sub, x);
#pragma unroll
                        for (int i = 0; i < EC; ++i) v[i] += x[i];
                    }
                    xr_store(sub, v);
                }
                wait_war();
                end_unit();
                }
            }
#pragma unroll 1
            for (int u = 0; u < NU; ++u) {
            set_unit(u);
#pragma unroll
            for (int sub = 0; sub < NP; ++sub) {
                float a[EC], x[EC];
                xr_load(sub, x);
#pragma unroll
                for (int i = 0; i < EC; ++i) {
                    const int ch = ch0 + sub * EC + i;
                    x[i] += c_b0[ch];
                    a[i] = fmaxf(fmaf(x[i], c_blk(0, 0)[ch], c_blk(0, 1)[ch]), 0.f);
                }
                xr_store(sub, x);
                store_act(sub, a);
            }
            signal_act();
            end_unit();
            }
            // ---- two residual blocks ---------------------------------------------------------------------------------
#pragma unroll 1
            for (int blk = 0; blk < 2; ++blk) {
#pragma unroll 1
                for (int u = 0; u < NU; ++u) {
                set_unit(u);
                wait_acc();
#pragma unroll
                for (int sub = 0; sub < NP; ++sub) {
                    float v[EC], a[EC];
                    load_acc(G < NJ ? G : NJ, sub, v);  // conv1 (second BatchNorm of the block folded into weights and bias)
#pragma unroll
                    for (int i = 0; i < EC; ++i) a[i] = fmaxf(v[i] + c_blk(blk, 2)[ch0 + sub * EC + i], 0.f);
                    if (sub == 0) wait_war();
                    store_act(sub, a);
                }
                signal_act();
                end_unit();
                }
                const float* sN = blk == 0 ? c_blk(1, 0) : c_sO;  // the BatchNorm that consumes the updated stream
                const float* tN = blk == 0 ? c_blk(1, 1) : c_tO;
#pragma unroll 1
                for (int u = 0; u < NU; ++u) {
                set_unit(u);
                wait_acc();
#pragma unroll
                for (int sub = 0; sub < NP; ++sub) {
                    float v[EC], a[EC], x[EC];
                    load_acc(G < NJ ? G : NJ, sub, v);  // conv2 + skip
                    xr_load(sub, x);
#pragma unroll
                    for (int i = 0; i < EC; ++i) {
                        const int ch = ch0 + sub * EC + i;
                        x[i] += v[i] + c_blk(blk, 3)[ch];
                        a[i] = fmaxf(fmaf(x[i], sN[ch], tN[ch]), 0.f);
                    }
                    xr_store(sub, x);
                    if (sub == 0) wait_war();
                    store_act(sub, a);
                }
                signal_act();
                end_unit();
                }
            }
            // ---- output layer ----------------------------------------------------------------------------------------
            tl.stamp(41);
#pragma unroll 1
            for (int u = 0; u < NU; ++u) {
            set_unit(u);
            float ssum = 0.f;
            // z0 of the next chunk is fetched before its accumulators are awaited: the loads do not depend on the conditioner
            constexpr int ZP = HW > 128 ? 8 : 32;  // prefetched z0 values per thread and chunk (beyond that: loaded in place)
            float zpre[ZP];
            auto prefetch_z0 = [&](int q) {
                const int PC = NW / 2, m0 = q * PC;
                const int i0 = (CS == 2) ? grp * (PC / 2) : 0, n = PC / CS;
                const float* zo = zdst + static_cast<size_t>(b) * g.D;
#pragma unroll
                for (int k = 0; k < ZP; ++k) {
                    const int m = m0 + i0 + k;
                    zpre[k] = (valid && k < n && m < g.c0)
                                  ? zo[half_elem_offset<(MODE < 0 ? NFB_SPLIT_CHANNEL : MODE)>(g, m, 0, yy, xx)] : 0.f;
                }
            };
            if (FUSED) prefetch_z0(0);
            const bool saturated = F16 && !(amax < 65504.f);
#pragma unroll 1
            for (int q = 0; q < nq; ++q) {
                wait_acc();
                const float* ob = c_ob + q * NW;
                if (FUSED) {
                    // columns [0, PC) = t, [PC, 2PC) = s_raw of channels q*PC ...; AffineCoupling._transform in place on z0
                    const int PC = NW / 2, m0 = q * PC;
                    const int i0 = (CS == 2) ? grp * (PC / 2) : 0, i1 = i0 + PC / CS;
                    float* zo = zdst + static_cast<size_t>(b) * g.D;
#pragma unroll
                    for (int k4 = 0; k4 < ZP / 4; ++k4) {
                        const int i = i0 + 4 * k4;
                        if (i < i1) {
                            float tm[4], sm[4], tcp[4], scp[4];
                            tmem_ld<4>(t_lane + static_cast<uint32_t>(i), tm);
                            tmem_ld<4>(t_lane + static_cast<uint32_t>(PC + i), sm);
                            tmem_ld<4>(t_lane + static_cast<uint32_t>(NW + i), tcp);
                            tmem_ld<4>(t_lane + static_cast<uint32_t>(NW + PC + i), scp);
#pragma unroll
                            for (int r = 0; r < 4; ++r) {
                                const int m = m0 + i + r;
                                if (valid && m < g.c0) {
                                    float t = (F16 ? fmaf(tcp[r], kLoInv, tm[r]) : tm[r] + tcp[r]) + ob[i + r];
                                    float sraw = (F16 ? fmaf(scp[r], kLoInv, sm[r]) : sm[r] + scp[r]) + ob[PC + i + r];
                                    if (saturated) t = sraw = __int_as_float(0x7fc00000);
                                    const int off = half_elem_offset<(MODE < 0 ? NFB_SPLIT_CHANNEL : MODE)>(g, m, 0, yy, xx);
                                    // coupling.py:107-109: two rounded ops each, no FMA contraction (as in coupling_affine.cu)
                                    const float s = __fadd_rn(__fmul_rn(tanhf(sraw), sa), sb);
                                    zo[off] = __fadd_rn(__fmul_rn(zpre[4 * k4 + r], expf(s)), t);
                                    ssum += s;
                                }
                            }
                        }
                    }
                    for (int i = i0 + ZP; i < i1; i += 4) {  // more than ZP channels per thread (c0 > 64 at 16x16 only)
                        float tm[4], sm[4], tcp[4], scp[4];
                        tmem_ld<4>(t_lane + static_cast<uint32_t>(i), tm);
                        tmem_ld<4>(t_lane + static_cast<uint32_t>(PC + i), sm);
                        tmem_ld<4>(t_lane + static_cast<uint32_t>(NW + i), tcp);
                        tmem_ld<4>(t_lane + static_cast<uint32_t>(NW + PC + i), scp);
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            const int m = m0 + i + r;
                            if (valid && m < g.c0) {
                                float t = (F16 ? fmaf(tcp[r], kLoInv, tm[r]) : tm[r] + tcp[r]) + ob[i + r];
                                float sraw = (F16 ? fmaf(scp[r], kLoInv, sm[r]) : sm[r] + scp[r]) + ob[PC + i + r];
                                if (saturated) t = sraw = __int_as_float(0x7fc00000);
                                const int off = half_elem_offset<(MODE < 0 ? NFB_SPLIT_CHANNEL : MODE)>(g, m, 0, yy, xx);
                                const float s = __fadd_rn(__fmul_rn(tanhf(sraw), sa), sb);
                                zo[off] = __fadd_rn(__fmul_rn(zo[off], expf(s)), t);
                                ssum += s;
                            }
                        }
                    }
                    if (q + 1 < nq) prefetch_z0(q + 1);
                } else {
                    const int i0 = (CS == 2) ? grp * (NW / 2) : 0, i1 = i0 + NW / CS;
                    for (int i = i0; i < i1; i += 8) {
                        float mv[8], cv[8];
                        tmem_ld<8>(t_lane + static_cast<uint32_t>(i), mv);
                        tmem_ld<8>(t_lane + static_cast<uint32_t>(NW + i), cv);
#pragma unroll
                        for (int r = 0; r < 8; ++r) {
                            const int oc = q * NW + i + r;
                            if (valid && oc < Cout)
                                zdst[(static_cast<size_t>(b) * Cout + oc) * HW + pix] =
                                    saturated ? __int_as_float(0x7fc00000) : (F16 ? fmaf(cv[r], kLoInv, mv[r]) : mv[r] + cv[r]) + ob[i + r];
                        }
                    }
                }
                if (q + 1 < nq) {  // accumulator columns are free for the next chunk
                    tc_fence_before();
                    mbar_arrive(b_act);
                }
            }
            if (FUSED) {
                // per-sample log-det: fixed-order reduction (segmented shuffle -> shared memory -> one thread per sample)
                constexpr int SEG = HW < 32 ? HW : 32;
                float* red = red_all + 16 * cu;  // one array per unit in flight: no barrier is needed before its next use
                const int bb_l = (unit * NU + cu) * SPU + tid;
                float ldj_old = 0.f;
                if (tid < SPU && bb_l < B) ldj_old = ldj[bb_l];  // issued before the barrier: its latency hides behind it
#pragma unroll
                for (int o = SEG / 2; o > 0; o >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
                if ((lane & (SEG - 1)) == 0) red[warp * 2 + lane / SEG] = ssum;
                asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
                if (tid < SPU) {
                    const int bb = bb_l;
                    if (bb < B) {
                        float tot = 0.f;
                        if (HW > 128) {
#pragma unroll
                            for (int w8 = 0; w8 < 8; ++w8) tot += red[w8 * 2];
                        } else if (LINKED) {  // paired tiles: warps 4*tile .. 4*tile+3 hold tile `tile`, all 32 channels
                            const int tl_ = tid / SPT, si = tid % SPT;
                            if (HW == 16) tot = red[(tl_ * 4 + (si >> 1)) * 2 + (si & 1)];
                            else if (HW == 64) tot = red[(tl_ * 4 + 2 * si) * 2] + red[(tl_ * 4 + 2 * si + 1) * 2];
                            else tot = (red[(tl_ * 4) * 2] + red[(tl_ * 4 + 1) * 2]) + (red[(tl_ * 4 + 2) * 2] + red[(tl_ * 4 + 3) * 2]);
                        } else if (HW == 16) tot = red[(tid >> 1) * 2 + (tid & 1)] + red[((tid >> 1) + 4) * 2 + (tid & 1)];
                        else if (HW == 64) tot = (red[(2 * tid) * 2] + red[(2 * tid + 1) * 2]) + (red[(2 * tid + 4) * 2] + red[(2 * tid + 5) * 2]);
                        else tot = ((red[0] + red[2]) + (red[4] + red[6])) + ((red[8] + red[10]) + (red[12] + red[14]));
                        float l = __fadd_rn(ldj_old, tot);  // coupling.py:110
                        if (CP > 0) {
                            const float hw_full = static_cast<float>(g.HW);
                            l = __fadd_rn(l, __fmul_rn(wpost[CP * CP + 2 * CP], hw_full));      // ActNorm, modules.py:249
                            l = __fadd_rn(l, __fmul_rn(wpost[CP * CP + 2 * CP + 1], hw_full));  // 1x1 conv, modules.py:480
                        }
                        ldj[bb] = l;
                    }
                }
                // (the next write to this unit's `red` is a whole unit of mbarrier handshakes away, all of which this thread's
                // warp takes part in: no second barrier unless the post-op below needs this CTA's z stores)
                if (CP > 0) {
                    __threadfence_block();  // this CTA's z0 stores are visible to all its threads after the barrier
                    asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
                }
                if (CP > 0) {
                    // next step's ActNorm + 1x1 conv, in place: thread = pixel; the CP normalised inputs stay in registers, the
                    // outputs are produced one channel at a time (rolled loop over co, unrolled dot product over ci: the same
                    // ascending-ci fmaf chain as invconv_apply_tiled) and stored at once -- the originals are no longer needed
                    const float* es = wpost + CP * CP;
                    const float* bs = es + CP;
                    const int npix = SPU * g.HW;
                    for (int pp = tid; pp < npix; pp += kEpiThreads) {
                        const int sl = pp / g.HW, px = pp - sl * g.HW;
                        const int bb = (unit * NU + cu) * SPU + sl;
                        if (bb >= B) continue;
                        float* zp = zdst + static_cast<size_t>(bb) * g.D + px;
                        float vn[CP > 0 ? CP : 1];
#pragma unroll
                        for (int c = 0; c < CP; ++c) vn[c] = zp[static_cast<size_t>(c) * g.HW];
#pragma unroll
                        for (int c = 0; c < CP; ++c) vn[c] = __fdiv_rn(__fsub_rn(vn[c], bs[c]), es[c]);  // modules.py:246
                        // COB output channels at a time: COB independent fmaf chains (each still ascending in ci)
                        constexpr int COB = CP >= 48 ? 8 : (CP % 4 == 0 ? 4 : (CP > 0 ? CP : 1));
#pragma unroll 1
                        for (int co = 0; co < CP; co += COB) {
                            const float* wr = wpost + co;
                            float acc[COB];
#pragma unroll
                            for (int r = 0; r < COB; ++r) acc[r] = 0.f;
#pragma unroll
                            for (int ci = 0; ci < CP; ++ci) {
#pragma unroll
                                for (int r = 0; r < COB; ++r) acc[r] = fmaf(wr[ci * CP + r], vn[ci], acc[r]);  // modules.py:477
                            }
#pragma unroll
                            for (int r = 0; r < COB; ++r) zp[static_cast<size_t>(co + r) * g.HW] = acc[r];
                        }
                    }
                }
            }
            end_unit();
            }  // units of this iteration
        }
        tl.stamp(42);
        tc_fence_before();
    }

    __syncthreads();
    tl_k.stamp(50);
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
    }
}

// =====================================================================================================================
// launch
// =====================================================================================================================
// `packed` = the whole buffer of nfb_resnet_pack: FFMA section | TF32 section | FP16-split section
template <int H, int W, int MODE, bool FUSED, bool PAIR, int CP = 0, bool F16 = false, bool DUAL = false>
static int launch_tc(const float* zsrc, float* zdst, float* ldj, const float* packed, const SplitGeom& g, int Cin, int Cout, int B,
                     const float* sa, const float* sb, int flags, cudaStream_t st, const PostOp& post = PostOp{}) {
    using GM = TcGeom<H, W, PAIR, F16, DUAL>;
    const int dbg = (flags >> NFB_CONV_DEBUG_SHIFT) & 0xff;
    const int gq = (flags >> NFB_CONV_GROUPS_SHIFT) & 7;
    const int G = gq >= 1 && gq <= 4 ? gq : 3;  // accumulator groups per layer (at most the k-steps of a tap: 4 TF32 / 2 FP16)
    const TcPlan P = tc_plan(Cin, Cout, F16 ? 1 : 0);
    const float* pk_tc = packed + P.base;
    const int n_cst = 352 + (FUSED ? P.nqf * P.NWf : P.nqg * P.NWg);
    const size_t smem = static_cast<size_t>(GM::NU * GM::ACT_BYTES) + 2 * slot_bytes(F16) + GM::XS_BYTES + static_cast<size_t>((n_cst + 3) & ~3) * 4 +
                        ((4 + 5 * GM::NU) | 1) * 8 + 8 + 64 * GM::NU + 2 * 9 * 16 + 2 * (2 * GM::NU * GM::T * 9 * (1 + GM::NJ) * 16) +
                        (CP > 0 ? static_cast<size_t>(CP * CP + 2 * CP + 4) * 4 : 0);
    if (smem > 227 * 1024) return NFB_ERR_UNSUPPORTED;
    if (DUAL && 2 * (FUSED ? P.NWf : P.NWg) > GM::TC) return NFB_ERR_UNSUPPORTED;  // output chunk wider than a tile's columns
    auto kern = convnet_tc_kernel<H, W, MODE, FUSED, PAIR, CP, F16, DUAL>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    const int n_units = ((B + GM::SPU - 1) / GM::SPU + GM::NU - 1) / GM::NU;
    int grid = n_units < kSMs ? n_units : kSMs;
    // throughput mode: a CTA that runs several iterations pays the kernel prologue and the cold instruction cache once --
    // less SM-time per sample on fewer SMs (the other batches in flight use the rest), at the price of this launch's latency
    const int ipc = (flags >> NFB_CONV_ITERS_SHIFT) & 3;
    if ((flags & NFB_CONV_PAIR) && ipc > 0 && n_units > 1) {
        const int per = ipc + 1;
        const int gq2 = (n_units + per - 1) / per;
        if (gq2 < grid) grid = gq2;
    }
    kern<<<grid, kThreads, smem, st>>>(zsrc, zdst, ldj, pk_tc, g, Cin, Cout, B, sa, sb, post, G, dbg);
    return launch_status();
}


}  // namespace nfb
