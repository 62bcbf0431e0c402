// logmel_tc_core.cuh -- the 2048-point STFT frame transform on the sm_100a tensor cores (tcgen05 + TMEM).
//
// Replaces torch.stft's FFT (TA/functional/functional.py:123-145 -> torch.stft) for n_fft = 2048.
// A tile is 8 frames, owned by one 8-warp group of the log-mel kernel.  The real FFT of a frame is the
// 1024-point complex FFT of z[m] = x[2m] + i x[2m+1], m = 32 m1 + m2, k = q1 + 32 q2, in two radix-32
// stages; each stage is a real [128 rows] x [K = 64] x [N = 64] GEMM issued by ONE thread with
// tcgen05.mma (kind::f16, fp32 accumulators in TMEM), in split precision: operands are fp16 head +
// fp16 residual, three passes (head.head + residual.head + head.residual), K order 32 c + idx, N order
// 32 c' + q (logmel_tc_tables.h).  Between the MMAs the CUDA cores only move and convert:
//
//   B0  warp = frame, lane = m2: samples x Hann (hann[n+1024] = 1 - hann[n]), per-frame power-of-two scale
//       (frame peak -> [2^9, 2^10): fp16 residuals stay normal), head/residual -> tcgen05.st: the stage-1
//       A operand lives in TMEM (lane = row (frame % 4, m2)), so its MMAs run at the N/2-cycle floor
//       (profiles/r2/ubench_tc_dft.txt: A from shared memory costs 48 instead of 32 cycles at N = 64).
//   S1  D1[(t, m2), (c', q1)] = A1 . G                                        24 x UTCHMMA (2 x 4 frames)
//   B1  warp = frame: tcgen05.ld D1, twiddle W_1024^(q1 m2), head/residual -> shared memory as the M-major
//       stage-2 A operand.  Stage 2 runs as TWO MMAs over the same 8 frames, rows (t, j), j = 0..15:
//         MMA-A row j: y'[q1 = j]                     -> D2A[p] = Z[j + 32 p]
//         MMA-B row j: conj(y[q1'])  W_1024^(j' m2)   -> D2B[p] = conj Z[1024 - (j + 32 p)]
//       (q1' = 32 - j, j' = j; row 0: q1' = 16, j' = 16), so the real-FFT untangle partner of a bin sits in
//       the same TMEM lane and the same column of the other accumulator: no shuffle, no transpose.
//   S2  12 x UTCHMMA each, A from shared memory (M-major, no swizzle), one 32 KB buffer used twice.
//   B2  tcgen05.ld D2A, D2B: E' = a + b, T' = -i W_2048^k (a - b); 4|X[k]|^2 = |E' + T'|^2,
//       4|X[1024-k]|^2 = |E' - T'|^2, times the inverse of the frame scale -> the power rows the mel
//       phase reads.  Lanes j = 0 hold the self-paired rows q1 = 0 and q1 = 16 (bins k = 0, 16 mod 32):
//       their raw values go through a 4 KB scratch and one thread per bin pair finishes them (fix-up).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "lm_f2.cuh"

namespace lmtc {

constexpr int kTileF = 8;                    // frames per tile
constexpr int kGroupWarps = 8;
constexpr int kGroupThreads = 256;
constexpr int kPPitch = 1040;                // floats per power row: 1024 bins + pad, == 16 (mod 32) for the mel loads
constexpr int kA2Half = 16384;               // stage-2 A operand: head (16 KB) then residual (16 KB)
constexpr int kA2Bytes = kTileF * kPPitch * 4;   // the buffer also holds the 8 power rows (33 280 B >= 32 768 B)
constexpr int kScratchFloats = kTileF * 128; // fix-up scratch: per frame ar[32] ai[32] br[32] bi[32]
constexpr int kGBytes = 8192;
constexpr int kTw1Rows = 17, kUtwRows = 17, kUtwPitch = 34;
constexpr int kTmemColsPerGroup = 256;

// ----------------------------------------------------------------------------- PTX
__device__ __forceinline__ uint32_t s32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ bool elect_one() {
    uint32_t p;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(p));
    return p != 0;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
#define LMTC_LD16(r, addr)                                                                                                    \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"      \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), \
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])                    \
                 : "r"(addr))
#define LMTC_ST32(addr, r)                                                                                                    \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,"    \
                 "%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"                                              \
                 ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),          \
                   "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),   \
                   "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]),             \
                   "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory")
__device__ __forceinline__ void ldtm_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void sttm_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor (UMMA, Blackwell version bit set); layout 0 = no swizzle
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return static_cast<uint64_t>((saddr >> 4) & 0x3FFF) | (static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16) |
           (static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor: fp16 x fp16 -> fp32, M = 128, N = 64; bit 15 = A is MN-major
constexpr uint32_t kIdescTS = (1u << 4) | (uint32_t(64 >> 3) << 17) | (uint32_t(128 >> 4) << 24);
constexpr uint32_t kIdescSS = kIdescTS | (1u << 15);

// head / residual of two values -> two fp16x2 words (low half = first value)
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// What one group needs to know about its buffers (all warp-uniform).
struct Ctx {
    uint32_t tm;          // TMEM base of the group (column offset included): A1 / D2 at +0 .. +127, D1 at +128 .. +255
    uint32_t bar;         // shared address of the group's MMA-completion mbarrier
    uint32_t a2;          // shared address of the stage-2 operand buffer (also the power rows)
    uint64_t g_hi, g_lo;  // descriptors of the DFT matrices (K-major, LBO 128, SBO 1024)
    float* prow;          // generic pointer to the same buffer as fp32 power rows [8][kPPitch]
    float* scratch;       // [8][128] fix-up scratch
    float* pscale;        // [8] inverse squared frame scale
    const float2* tw1;    // [17][32]
    const float* utw_c;   // [17][34]
    const float* utw_s;
};

// ----------------------------------------------------------------------------- B0
// sbf: the frame's 2048 staged samples; s_win: first half of the Hann window.  warp = frame gw, lane = m2.
__device__ __forceinline__ void b0_frame(const Ctx& cx, const float* __restrict__ sbf, const float* __restrict__ s_win, int gw, int lane) {
    const float2* __restrict__ s2 = reinterpret_cast<const float2*>(sbf);
    const float2* __restrict__ w2 = reinterpret_cast<const float2*>(s_win);
    float xr[32], xi[32];   // x*[m1] = windowed sample 64 m1 + 2 lane (+1): real / imaginary part of z[32 m1 + lane]
    float mx = 0.f;
#pragma unroll
    for (int m1 = 0; m1 < 16; ++m1) {
        const float2 v1 = s2[32 * m1 + lane], v2 = s2[32 * (m1 + 16) + lane], w = w2[32 * m1 + lane];
        xr[m1] = v1.x * w.x;
        xi[m1] = v1.y * w.y;
        xr[m1 + 16] = fmaf(-v2.x, w.x, v2.x);
        xi[m1 + 16] = fmaf(-v2.y, w.y, v2.y);
        mx = fmaxf(mx, fmaxf(fabsf(xr[m1]), fabsf(xi[m1])));
        mx = fmaxf(mx, fmaxf(fabsf(xr[m1 + 16]), fabsf(xi[m1 + 16])));
    }
    // frame peak -> [2^9, 2^10): exponent arithmetic only, so the scale is an exact power of two
    uint32_t e = __reduce_max_sync(0xffffffffu, __float_as_uint(mx)) >> 23;
    e = e < 76u ? 76u : (e > 196u ? 196u : e);
    const float scale = __uint_as_float((263u - e) << 23);
    if (lane == 0) cx.pscale[gw] = __uint_as_float((2u * e - 145u) << 23);   // 2^(-2 log2 scale)
    uint32_t hi[32], lo[32];   // TMEM word 16 c + i = K indices 32 c + 2i (low half), 32 c + 2i + 1
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        split2(xr[2 * i] * scale, xr[2 * i + 1] * scale, hi[i], lo[i]);
        split2(xi[2 * i] * scale, xi[2 * i + 1] * scale, hi[16 + i], lo[16 + i]);
    }
    const uint32_t ta = cx.tm + (static_cast<uint32_t>((gw & 3) * 32) << 16) + static_cast<uint32_t>((gw >> 2) * 64);
    LMTC_ST32(ta, hi);
    LMTC_ST32(ta + 32, lo);
    sttm_wait();
}

// ----------------------------------------------------------------------------- MMA issue (one warp, uniform)
__device__ __forceinline__ void issue_stage1(const Ctx& cx, bool leader) {
    tc_fence_after();
    if (leader) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const uint32_t d = cx.tm + 128 + 64 * g, ah = cx.tm + 64 * g, al = ah + 32;
#pragma unroll
            for (int s = 0; s < 4; ++s) mma_ts(d, ah + 8 * s, cx.g_hi + 16 * s, kIdescTS, s > 0);
#pragma unroll
            for (int s = 0; s < 4; ++s) mma_ts(d, al + 8 * s, cx.g_hi + 16 * s, kIdescTS, 1);
#pragma unroll
            for (int s = 0; s < 4; ++s) mma_ts(d, ah + 8 * s, cx.g_lo + 16 * s, kIdescTS, 1);
        }
        tc_commit(cx.bar);
    }
    __syncwarp();
}
// which = 0: MMA-A -> D2A at +0; 1: MMA-B -> D2B at +64
__device__ __forceinline__ void issue_stage2(const Ctx& cx, bool leader, int which) {
    tc_fence_after();
    if (leader) {
        const uint32_t d = cx.tm + 64 * which;
        const uint64_t ah = make_desc(cx.a2, 128, 1024), al = make_desc(cx.a2 + kA2Half, 128, 1024);
#pragma unroll
        for (int s = 0; s < 4; ++s) mma_ss(d, ah + 16 * s, cx.g_hi + 16 * s, kIdescSS, s > 0);
#pragma unroll
        for (int s = 0; s < 4; ++s) mma_ss(d, al + 16 * s, cx.g_hi + 16 * s, kIdescSS, 1);
#pragma unroll
        for (int s = 0; s < 4; ++s) mma_ss(d, ah + 16 * s, cx.g_lo + 16 * s, kIdescSS, 1);
        tc_commit(cx.bar);
    }
    __syncwarp();
}

// ----------------------------------------------------------------------------- B1
// One half of the stage-2 operand for frame gw (lane = m2).  which = 0: rows j <-> q1 = j (MMA-A);
// which = 1: rows j <-> conj(y[q1']), q1' = 16 (j = 0) or 32 - j (MMA-B).  Returns the 16 packed words
// (8 head + 8 residual words: [part][chunk][4 words]) in `w` so that the caller can delay the stores.
//   w[0..7]   head:     re chunk 0 (4 words), re chunk 1, then (8..15) im chunk 0, im chunk 1
//   w[16..31] residual, same order
__device__ __forceinline__ void b1_half(const Ctx& cx, int gw, int lane, int which, uint32_t (&w)[32]) {
    const uint32_t td = cx.tm + (static_cast<uint32_t>((gw & 3) * 32) << 16) + 128 + static_cast<uint32_t>((gw >> 2) * 64);
    uint32_t ur[16], ui[16];
    LMTC_LD16(ur, td + 16 * which);
    LMTC_LD16(ui, td + 32 + 16 * which);
    ldtm_wait();
    float re[16], im[16];   // row j of this half
    if (which == 0) {
        re[0] = __uint_as_float(ur[0]);
        im[0] = __uint_as_float(ui[0]);
#pragma unroll
        for (int j = 1; j < 16; ++j) {   // y[j] W^(j m2): (yr + i yi)(wx + i wy)
            const float2 t = cx.tw1[j * 32 + lane];
            const float yr = __uint_as_float(ur[j]), yi = __uint_as_float(ui[j]);
            re[j] = fmaf(yr, t.x, -yi * t.y);
            im[j] = fmaf(yi, t.x, yr * t.y);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {   // conj(y[q1']) W^(j' m2): (yr - i yi)(wx + i wy); slot of q1' in this half = q1' - 16
            const int q = (j == 0) ? 0 : 16 - j, jp = (j == 0) ? 16 : j;
            const float2 t = cx.tw1[jp * 32 + lane];
            const float yr = __uint_as_float(ur[q]), yi = __uint_as_float(ui[q]);
            re[j] = fmaf(yr, t.x, yi * t.y);
            im[j] = fmaf(yr, t.y, -yi * t.x);
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        split2(re[2 * i], re[2 * i + 1], w[i], w[16 + i]);
        split2(im[2 * i], im[2 * i + 1], w[8 + i], w[24 + i]);
    }
}
// stores the words of b1_half: M-major, row m = 16 gw + j, chunk = m / 8 at SBO 1024, K row kappa = 32 c + m2 at 16 bytes
__device__ __forceinline__ void b1_store(const Ctx& cx, int gw, int lane, const uint32_t (&w)[32]) {
#pragma unroll
    for (int part = 0; part < 2; ++part)       // head, residual
#pragma unroll
        for (int c = 0; c < 2; ++c)            // real, imaginary
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {   // rows j = 0..7, 8..15
                const uint32_t addr = cx.a2 + part * kA2Half + static_cast<uint32_t>(2 * gw + ch) * 1024u +
                                      static_cast<uint32_t>(32 * c + lane) * 16u;
                const int o = 16 * part + 8 * c + 4 * ch;
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(w[o]), "r"(w[o + 1]), "r"(w[o + 2]), "r"(w[o + 3]) : "memory");
            }
    proxy_fence();   // the MMA reads these through the async proxy
}

// ----------------------------------------------------------------------------- B2
// warp gw reads TMEM lanes 32 (gw % 4) ..: rows (t, j) = (2 (gw % 4) + lane / 16, lane % 16), columns p0 .. p0 + 15,
// p0 = 16 (gw / 4); writes 4|X|^2 / scale^2 for k = j + 32 p and 1024 - k into the power rows.
// MUST be called after the stage-2 MMAs completed (the power rows overwrite the stage-2 operand).
__device__ __forceinline__ void b2_rows(const Ctx& cx, int gw, int lane) {
    const uint32_t td = cx.tm + (static_cast<uint32_t>((gw & 3) * 32) << 16);
    const int t = 2 * (gw & 3) + (lane >> 4), j = lane & 15, p0 = 16 * (gw >> 2);
    uint32_t ar[16], ai[16], br[16], bi[16];
    LMTC_LD16(ar, td + p0);
    LMTC_LD16(ai, td + 32 + p0);
    LMTC_LD16(br, td + 64 + p0);
    LMTC_LD16(bi, td + 96 + p0);
    ldtm_wait();
    float* __restrict__ row = cx.prow + t * kPPitch;
    if (j == 0) {   // self-paired rows q1 = 0 (A) and q1 = 16 (B): raw values to the scratch, finished by fixup()
        float* __restrict__ sc = cx.scratch + t * 128 + p0;
#pragma unroll
        for (int q = 0; q < 16; q += 4) {
            *reinterpret_cast<uint4*>(sc + q) = make_uint4(ar[q], ar[q + 1], ar[q + 2], ar[q + 3]);
            *reinterpret_cast<uint4*>(sc + 32 + q) = make_uint4(ai[q], ai[q + 1], ai[q + 2], ai[q + 3]);
            *reinterpret_cast<uint4*>(sc + 64 + q) = make_uint4(br[q], br[q + 1], br[q + 2], br[q + 3]);
            *reinterpret_cast<uint4*>(sc + 96 + q) = make_uint4(bi[q], bi[q + 1], bi[q + 2], bi[q + 3]);
        }
    }
    const lm_f2 ps = lm_bcast(cx.pscale[t]);
    const float* __restrict__ tc = cx.utw_c + j * kUtwPitch + p0;
    const float* __restrict__ ts = cx.utw_s + j * kUtwPitch + p0;
#pragma unroll
    for (int q = 0; q < 16; q += 2) {   // two columns per packed operation
        const lm_f2 Ar = lm_pack(__uint_as_float(ar[q]), __uint_as_float(ar[q + 1])), Ai = lm_pack(__uint_as_float(ai[q]), __uint_as_float(ai[q + 1]));
        const lm_f2 Br = lm_pack(__uint_as_float(br[q]), __uint_as_float(br[q + 1])), Bi = lm_pack(__uint_as_float(bi[q]), __uint_as_float(bi[q + 1]));
        const float2 c2 = *reinterpret_cast<const float2*>(tc + q), s2 = *reinterpret_cast<const float2*>(ts + q);
        const lm_f2 C = lm_pack(c2.x, c2.y), S = lm_pack(s2.x, s2.y);
        const lm_f2 Er = lm_add2(Ar, Br), Ei = lm_add2(Ai, Bi), Dr = lm_sub2(Ar, Br), Di = lm_sub2(Ai, Bi);
        // T' = -i (c - i s)(Dr + i Di) = (c Di - s Dr) - i (c Dr + s Di)
        const lm_f2 Tr = lm_fma2(C, Di, lm_mul2(lm_pack(-s2.x, -s2.y), Dr));
        const lm_f2 Tn = lm_fma2(C, Dr, lm_mul2(S, Di));   // = -Ti
        const lm_f2 Ur = lm_add2(Er, Tr), Ui = lm_sub2(Ei, Tn), Vr = lm_sub2(Er, Tr), Vi = lm_add2(Ei, Tn);
        const lm_f2 PU = lm_mul2(lm_fma2(Ur, Ur, lm_mul2(Ui, Ui)), ps), PV = lm_mul2(lm_fma2(Vr, Vr, lm_mul2(Vi, Vi)), ps);
        if (j != 0) {
            const int k = j + 32 * (p0 + q);
            row[k] = lm_lo(PU);
            row[k + 32] = lm_hi(PU);
            row[1024 - k] = lm_lo(PV);
            row[1024 - k - 32] = lm_hi(PV);
        }
    }
}
// One thread per bin pair of the self-paired rows: frame gw, lane = pair.  Call after a group barrier
// following b2_rows (scratch complete) and before the barrier that releases the power rows to the mel phase.
__device__ __forceinline__ void fixup(const Ctx& cx, int gw, int lane) {
    const float* __restrict__ sc = cx.scratch + gw * 128;
    float* __restrict__ row = cx.prow + gw * kPPitch;
    float ar, ai, br, bi;
    int k;
    if (lane < 16) {          // q1 = 0: a = Z[32 p], b = conj Z[32 (32 - p)] = conj A[(32 - p) % 32]; lane 0 takes p = 16 (k = 512)
        const int p = lane == 0 ? 16 : lane, pm = (32 - p) & 31;
        ar = sc[p]; ai = sc[32 + p]; br = sc[pm]; bi = -sc[32 + pm];
        k = 32 * p;
    } else {                  // q1 = 16: B[p] = conj Z[16 + 32 (31 - p)]:  a = Z[16 + 32 p] = conj B[31 - p],  b = B[p]
        const int p = lane - 16;
        ar = sc[64 + 31 - p]; ai = -sc[96 + 31 - p]; br = sc[64 + p]; bi = sc[96 + p];
        k = 16 + 32 * p;
    }
    const int jj = k & 31, pp = k >> 5;   // rows 0 and 16 of the untangle table
    const float c = cx.utw_c[jj * kUtwPitch + pp], s = cx.utw_s[jj * kUtwPitch + pp];
    const float er = ar + br, ei = ai + bi, dr = ar - br, di = ai - bi;
    const float tr = fmaf(c, di, -s * dr), tn = fmaf(c, dr, s * di);
    const float ur = er + tr, ui = ei - tn, vr = er - tr, vi = ei + tn;
    const float ps = cx.pscale[gw];
    row[k] = fmaf(ur, ur, ui * ui) * ps;
    row[1024 - k] = fmaf(vr, vr, vi * vi) * ps;
    if (lane == 0) { row[0] = 0.f; row[1024] = 0.f; }   // DC / Nyquist: zero mel weight, but must stay finite
}

}  // namespace lmtc
