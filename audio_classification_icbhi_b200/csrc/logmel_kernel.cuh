// logmel_kernel.cuh -- the fused log-mel kernel for sm_100a (B200).
//
// One persistent CTA per SM (16 warps).  A CTA takes whole clips (blockIdx.x, += gridDim.x) and
// walks a flat list of work items (clip, tile); a tile is 16 frames (32 for n_fft = 1024).
// Two CTA barriers per item: X sits in the MIDDLE of the FFT (before the transpose first touches
// the power rows), Y after it.  A warp that finishes its share of the mel phase of item i therefore
// runs straight into the register-only first half of the FFT of item i+1 while slower warps are
// still on the tensor cores.  Per item:
//
//   stage   (done TWO items ahead, into the buffer just consumed) the tile's samples -> shared memory,
//           once per sample although every sample feeds 4 frames.  The contiguous interior of
//           a plain clip is one cp.async.bulk (TMA, 1-D) completing on an mbarrier; whatever
//           is left -- torch.stft's reflect padding, zero padding, and whole tiles of augmented
//           clips -- goes through the gather path that does pad/crop
//           (R/src/data/preprocessing.py:70-83), noise and roll (:85-93) as index arithmetic.
//   FFT     one warp = one frame.  x Hann, 2048 real -> 1024 complex points held as
//           32 registers/lane; register radix-32 FFT over the lane-local index, twiddle,
//           32x32 transpose through the warp's shared-memory row, second radix-32 FFT,
//           then the real-FFT untangle with the partner bin fetched by warp shuffle;
//           4|X|^2 overwrites the warp's row (never HBM).
//           (TA/functional/functional.py:123-145: torch.stft + abs().pow(2))
//   mel     [16 frames x bins] . [bins x 8 mels] per warp on the tensor cores:
//           mma.sync m16n8k8 TF32 with both operands split hi+lo (3 MMAs per step, error
//           ~2^-22), walking only the band of bins the 8 filters touch
//           (TA/transforms/_transforms.py:417).  Epilogue in registers: 10*log10(max(x, amin))
//           (TA/functional/functional.py:390-391), SpecAugment intervals (:939-953), store,
//           fp64 sum / sum-of-squares.
//   norm    when the clip is finished the same CTA re-reads its (L2-resident) dB block and
//           writes (x - mean) / (std + eps)  (R/src/data/preprocessing.py:111-116).
//
// Algorithmic HBM bytes per clip: 4*min(len, T) read + 4*n_mels*frames written.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/logmel_b200.h"
#include "fft_gen.cuh"

// Build-time experiment switches (tools/build_variants.py); the defaults are the shipped kernel.
#ifndef LM_TW2
#define LM_TW2 1      // 1: twiddle = product of two table entries (10 loads); 0: 31 table loads
#endif
#ifndef LM_ALWAYS_ACTIVE
#define LM_ALWAYS_ACTIVE 1   // 1: warps whose frame is past the clip end compute anyway (no divergent join)
#endif
#ifndef LM_STAGGER_NS
#define LM_STAGGER_NS 0     // > 0: warps of one scheduler start the FFT (warp/4) * ns apart
#endif
#ifndef LM_F32_TILE_STATS
#define LM_F32_TILE_STATS 0 // 1: per-thread per-tile sums in fp32, accumulated across tiles in fp64
#endif
#ifndef LM_WINCALC
#define LM_WINCALC 0    // 1: Hann window by angle addition in registers (2 FFMA2 per pair) instead of 32 LDS.64
#endif
#ifndef LM_TW_P1
#define LM_TW_P1 0      // 1: twiddle multiply before barrier X (in part 1) instead of after it
#endif
#ifndef LM_SPLIT
#define LM_SPLIT 1    // 1: CTA barrier X between the two halves of the FFT; 0: before the FFT
#endif

namespace lm {

constexpr int kWarps = 16;
constexpr int kThreads = kWarps * 32;
constexpr int kScrPitch = 36;                 // transpose rows: 32 + 4 (16 B aligned, LDS.128 conflict-free)
constexpr int kRowFloats = 1168;              // per-warp row: >= 32*36 and == 16 (mod 32) for the MMA loads
constexpr int kPbOff = 528;                   // n_fft=1024: second frame's spectrum inside the row (== 16 mod 32)
constexpr int kMaxMelTiles = 32;              // n_mels <= 256
constexpr int kMaxDk = 96;                    // 16-bin steps of banded filterbank kept on chip (48 KB)

struct MelTable {                             // lives in global memory, copied to shared
    int kb[kMaxMelTiles];                     // first bin of the tile's band (multiple of 4)
    int ndk[kMaxMelTiles];                    // 16-bin steps in the band
    int off[kMaxMelTiles];                    // first step's index into melw (units of 32 float4)
    int warp_tile[kWarps][2];                 // mel tiles owned by each warp (-1 = none)
};

struct KParams {
    // batch
    const float* __restrict__ wave;
    const long long* __restrict__ offset;
    const int* __restrict__ length;
    const lm_aug* __restrict__ aug;
    const float* __restrict__ noise;
    float* __restrict__ out_norm;
    float* __restrict__ out_db;
    float* __restrict__ out_melpow;
    int B;
    int normalize;
    // plan
    int T, hop, frames, n_mels, n_tiles;
    int ns;          // staged floats per tile = (TILE_F-1)*hop + NFFT, rounded up to 4
    int n_dk;        // total 16-bin steps in melw
    int use_tma;
    float db_scale;  // db_multiplier * log10(2): dB = db_scale * log2(x) - db_offset
    float amin, db_offset, floor_db, norm_eps;
    const float* __restrict__ window;   // [NFFT]
    const float2* __restrict__ tw;      // [32*32]  W1024^(n2*k1) = (cos, -sin), index k1*32+n2
    const float2* __restrict__ utw;     // [512]    (cos, sin)(2 pi k / 2048)
    const float4* __restrict__ wphase;  // [32]     (cos p0, cos p1, sin p0, sin p1), p_j = 2 pi (2 lane + j) / n_fft
    const float4* __restrict__ melw;    // [n_dk][32 lanes]: fb/4 in mma B-fragment order
    const MelTable* __restrict__ mel_table;
};

// ---------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk copy (TMA) global -> shared, TF32 mma
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// D += A(16x8, row) * B(8x8, col), TF32 inputs (low 13 mantissa bits ignored), fp32 accumulate
__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
    asm(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// Opaque copy: stops the compiler from hoisting address arithmetic that depends on `v` out of the
// item loop and keeping it alive across the register-heavy FFT (it is recomputed per phase instead).
__device__ __forceinline__ int launder(int v) {
    asm volatile("" : "+r"(v));
    return v;
}
__device__ __forceinline__ uint32_t tf32_hi(float x) { return __float_as_uint(x) & 0xffffe000u; }
__device__ __forceinline__ uint32_t tf32_lo(float x, uint32_t hi) { return __float_as_uint(x - __uint_as_float(hi)); }

// ---------------------------------------------------------------------------------------
// Philox4x32-10 -> N(0,1): throughput-mode noise when no host-drawn noise is supplied
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float philox_normal(uint64_t seed, uint32_t idx) {
    uint32_t c0 = idx >> 2, c1 = 0u, c2 = 0x6c6f676du, c3 = 0x656c0000u;
    uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        c0 = h1 ^ c1 ^ k0; c1 = l1; c2 = h0 ^ c3 ^ k1; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    const uint32_t a = (idx & 2u) ? c2 : c0, b = (idx & 2u) ? c3 : c1;
    const float u1 = (static_cast<float>(a >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float u2 = (static_cast<float>(b >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float r = sqrtf(-2.0f * __logf(u1));
    float s, c;
    __sincosf(6.283185307179586f * u2, &s, &c);
    return (idx & 1u) ? r * s : r * c;
}

// ---------------------------------------------------------------------------------------
// 1024-point complex FFT of one warp, in two halves so that a CTA barrier can sit between them.
// On entry lane n2 holds z[32*n1 + n2] in z[n1] (re, im packed in one 64-bit register); on exit lane
// k1 holds Z[k1 + 32*k2] in (xr[k2], xi[k2]).
//   part 1 (registers only): packed complex radix-32 FFT over the lane-local index;
//   part 2: twiddle, 32x32 transpose through the warp's private row `scr` (real and imaginary planes
//           separately; the LDS.128 reads hand the second FFT register pairs of neighbouring
//           points, the layout its packed stages 1-4 want -- see gen_fft.py) + second radix-32 FFT.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_twiddle(lm_f2 (&z)[32], const float2* __restrict__ tw, int lane) {
#if LM_TW2
    // Twiddle W1024^(lane*k1), k1 = 4a + b, as the product of two table entries W^(4a*lane) * W^(b*lane):
    // 10 shared-memory loads instead of 31 (the shared-memory pipe is the tighter resource here, and
    // 31 loads in flight on top of z do not fit the register file), one extra rounding per twiddle.
    // tw holds (cos, -sin); (r + i m)(wx + i wy) = (r wx - m wy, m wx + r wy).
    auto cmul = [](lm_f2 v, lm_f2 w) {
        return lm_fma2(lm_swap(v), lm_pack(-lm_hi(w), lm_hi(w)), lm_mul2(v, lm_bcast(lm_lo(w))));
    };
    auto ldtw = [&](int k1) {
        const float2 w = tw[k1 * 32 + lane];
        return lm_pack(w.x, w.y);
    };
    const lm_f2 B1 = ldtw(1), B2 = ldtw(2), B3 = ldtw(3);
    z[1] = cmul(z[1], B1);
    z[2] = cmul(z[2], B2);
    z[3] = cmul(z[3], B3);
#pragma unroll
    for (int a = 1; a < 8; ++a) {
        const lm_f2 A = ldtw(4 * a);
        z[4 * a] = cmul(z[4 * a], A);
        z[4 * a + 1] = cmul(z[4 * a + 1], cmul(A, B1));
        z[4 * a + 2] = cmul(z[4 * a + 2], cmul(A, B2));
        z[4 * a + 3] = cmul(z[4 * a + 3], cmul(A, B3));
    }
#else
#pragma unroll
    for (int k1 = 1; k1 < 32; ++k1) {
        const float2 w = tw[k1 * 32 + lane];   // (cos, -sin): (r + i m)(wx + i wy) = (r wx - m wy, m wx + r wy)
        z[k1] = lm_fma2(lm_swap(z[k1]), lm_pack(-w.y, w.y), lm_mul2(z[k1], lm_bcast(w.x)));
    }
#endif
}
__device__ __forceinline__ void warp_cfft1024_part1(lm_f2 (&z)[32], const float2* __restrict__ tw, int lane) {
    lm_fft32_aos(z);
    if (LM_TW_P1) warp_twiddle(z, tw, lane);
}
__device__ __forceinline__ void warp_cfft1024_part2(lm_f2 (&z)[32], float (&xr)[32], float (&xi)[32],
                                                    float* __restrict__ scr, const float2* __restrict__ tw, int lane) {
    if (!LM_TW_P1) warp_twiddle(z, tw, lane);
    lm_f2 pr[16], pi[16];
#pragma unroll
    for (int k1 = 0; k1 < 32; ++k1) scr[k1 * kScrPitch + lane] = lm_lo(z[k1]);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(scr + lane * kScrPitch + 4 * q);
        pr[2 * q] = lm_pack(v.x, v.y);
        pr[2 * q + 1] = lm_pack(v.z, v.w);
    }
    __syncwarp();
#pragma unroll
    for (int k1 = 0; k1 < 32; ++k1) scr[k1 * kScrPitch + lane] = lm_hi(z[k1]);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(scr + lane * kScrPitch + 4 * q);
        pi[2 * q] = lm_pack(v.x, v.y);
        pi[2 * q + 1] = lm_pack(v.z, v.w);
    }
    __syncwarp();
    lm_fft32_soa(pr, pi, xr, xi);
}

// -0.5 cos(2 pi n/32) and 0.5 sin(2 pi n/32): the n1-dependent half of the window's angle addition
__device__ constexpr float kWinCa[32] = {-5.000000000e-01f, -4.903926402e-01f, -4.619397663e-01f, -4.157348062e-01f, -3.535533906e-01f, -2.777851165e-01f, -1.913417162e-01f, -9.754516101e-02f, -3.061616998e-17f, 9.754516101e-02f, 1.913417162e-01f, 2.777851165e-01f, 3.535533906e-01f, 4.157348062e-01f, 4.619397663e-01f, 4.903926402e-01f, 5.000000000e-01f, 4.903926402e-01f, 4.619397663e-01f, 4.157348062e-01f, 3.535533906e-01f, 2.777851165e-01f, 1.913417162e-01f, 9.754516101e-02f, 9.184850994e-17f, -9.754516101e-02f, -1.913417162e-01f, -2.777851165e-01f, -3.535533906e-01f, -4.157348062e-01f, -4.619397663e-01f, -4.903926402e-01f};
__device__ constexpr float kWinSa[32] = {0.000000000e+00f, 9.754516101e-02f, 1.913417162e-01f, 2.777851165e-01f, 3.535533906e-01f, 4.157348062e-01f, 4.619397663e-01f, 4.903926402e-01f, 5.000000000e-01f, 4.903926402e-01f, 4.619397663e-01f, 4.157348062e-01f, 3.535533906e-01f, 2.777851165e-01f, 1.913417162e-01f, 9.754516101e-02f, 6.123233996e-17f, -9.754516101e-02f, -1.913417162e-01f, -2.777851165e-01f, -3.535533906e-01f, -4.157348062e-01f, -4.619397663e-01f, -4.903926402e-01f, -5.000000000e-01f, -4.903926402e-01f, -4.619397663e-01f, -4.157348062e-01f, -3.535533906e-01f, -2.777851165e-01f, -1.913417162e-01f, -9.754516101e-02f};

template <int NFFT>
struct Geo {
    static constexpr int FPW = (NFFT == 2048) ? 1 : 2;   // frames per warp pass
    static constexpr int TILE_F = kWarps * FPW;
    static constexpr int MT = TILE_F / 16;               // 16-frame MMA row blocks per tile
};

// Shared-memory carve-up, shared by host (size) and device (pointers).  Everything of fixed size
// comes first so that its addresses are compile-time constants (no registers spent on pointers);
// the two run-time sized arrays (staging buffers, filterbank) sit at the end.
template <int NFFT>
struct Smem {
    static constexpr size_t kBar = 0;                                    // 2 mbarriers + 2 'TMA pending' flags
    static constexpr size_t kRed = kBar + 32;                            // block-reduction scratch + broadcast
    static constexpr size_t kCtx = kRed + sizeof(double) * 2 * kWarps + 16;   // four ClipCtx slots (ordinal & 3)
    static constexpr size_t kTab = kCtx + 4 * 64;
    static constexpr size_t kStat = kTab + ((sizeof(MelTable) + 15) & ~size_t(15));   // per-thread fp64 (sum, sumsq)
    static constexpr size_t kWph = kStat + sizeof(double) * 2 * kThreads;          // window phase table
    static constexpr size_t kWin = kWph + sizeof(float4) * 32;
    static constexpr size_t kTw = kWin + sizeof(float) * NFFT;
    static constexpr size_t kUtw = kTw + sizeof(float2) * 1024;
    static constexpr size_t kScr = kUtw + ((NFFT == 2048) ? sizeof(float2) * 512 : 0);
    static constexpr size_t kSbuf = kScr + sizeof(float) * kWarps * kRowFloats;
    static_assert(kSbuf % 16 == 0 && kScr % 16 == 0 && kStat % 16 == 0 && kCtx % 16 == 0, "alignment");
    static __host__ __device__ size_t melw_offset(int ns) { return kSbuf + sizeof(float) * 2 * static_cast<size_t>(ns); }
    static __host__ __device__ size_t total(int ns, int n_dk) {
        return melw_offset(ns) + sizeof(float4) * 32 * static_cast<size_t>(n_dk);
    }
};

// Everything the staging and epilogue code needs to know about one clip.  Four slots live in
// shared memory (clip ordinal & 3: staging runs two items ahead of the epilogue) so that none of
// it occupies registers across the FFT.
struct ClipCtx {
    const float* src;    // first sample after the centre crop
    const float* nz;     // host-drawn noise row or nullptr
    uint64_t seed;
    int lc;              // valid samples after pad/crop
    int shift, f0, f1, t0, t1;
    float nscale, gain;
    int plain;
    int pad_;
};
static_assert(sizeof(ClipCtx) <= 64, "ClipCtx slot size");

__device__ __forceinline__ void load_clip(const KParams& p, int clip, ClipCtx* __restrict__ c) {
    const long long off = p.offset[clip];
    const int len = p.length[clip];
    const int crop = len > p.T ? (len - p.T) / 2 : 0;       // centre crop
    c->lc = len < p.T ? len : p.T;
    c->src = p.wave + off + crop;
    int shift = 0, f0 = 0, f1 = 0, t0 = 0, t1 = 0;
    float nscale = 0.0f, gain = 1.0f;
    uint64_t seed = 0;
    if (p.aug != nullptr) {
        const lm_aug a = p.aug[clip];
        shift = a.shift; nscale = a.noise_scale; gain = a.gain;
        f0 = a.f0; f1 = a.f1; t0 = a.t0; t1 = a.t1; seed = a.seed;
    }
    c->shift = shift; c->f0 = f0; c->f1 = f1; c->t0 = t0; c->t1 = t1;
    c->nscale = nscale; c->gain = gain; c->seed = seed;
    c->nz = (p.noise != nullptr && nscale != 0.0f) ? p.noise + static_cast<size_t>(clip) * p.T : nullptr;
    c->plain = (shift == 0) && (nscale == 0.0f) && (gain == 1.0f);
}

template <int NFFT, bool EXTRA_OUT>
__global__ void __launch_bounds__(kThreads, 1) logmel_kernel(const KParams p) {
    using G = Geo<NFFT>;
    constexpr int TILE_F = G::TILE_F;
    constexpr int HALF = NFFT / 2;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    using L = Smem<NFFT>;
    uint64_t* const mbar = reinterpret_cast<uint64_t*>(smem_raw + L::kBar);
    volatile int* const s_pend = reinterpret_cast<volatile int*>(smem_raw + L::kBar + 16);
    double* const red = reinterpret_cast<double*>(smem_raw + L::kRed);
    float* const bcast = reinterpret_cast<float*>(smem_raw + L::kRed + sizeof(double) * 2 * kWarps);
    ClipCtx* const s_ctx = reinterpret_cast<ClipCtx*>(smem_raw + L::kCtx);
    const MelTable* const s_tab = reinterpret_cast<const MelTable*>(smem_raw + L::kTab);
    double2* const s_stat = reinterpret_cast<double2*>(smem_raw + L::kStat);
    float4* const s_wph = reinterpret_cast<float4*>(smem_raw + L::kWph);
    float* const s_win = reinterpret_cast<float*>(smem_raw + L::kWin);
    float2* const s_tw = reinterpret_cast<float2*>(smem_raw + L::kTw);
    float2* const s_utw = reinterpret_cast<float2*>(smem_raw + L::kUtw);
    float* const scr_all = reinterpret_cast<float*>(smem_raw + L::kScr);
    float* const sbuf = reinterpret_cast<float*>(smem_raw + L::kSbuf);
    float4* const s_melw = reinterpret_cast<float4*>(smem_raw + L::melw_offset(p.ns));

    const int tid = threadIdx.x, lane_ = tid & 31, warp_ = tid >> 5;

    const int n_my = (p.B - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    if (n_my <= 0) return;
    const int n_items = n_my * p.n_tiles;

    // ---- constants -> shared memory, once per (persistent) CTA -----------------------------
    for (int i = tid; i < NFFT; i += kThreads) s_win[i] = p.window[i];
    for (int i = tid; i < 1024; i += kThreads) s_tw[i] = p.tw[i];
    if (tid < 32) s_wph[tid] = p.wphase[tid];
    if (NFFT == 2048)
        for (int i = tid; i < 512; i += kThreads) s_utw[i] = p.utw[i];
    for (int i = tid; i < 32 * p.n_dk; i += kThreads) s_melw[i] = p.melw[i];
    for (int i = tid; i < static_cast<int>(sizeof(MelTable) / 4); i += kThreads)
        reinterpret_cast<int*>(smem_raw + L::kTab)[i] = reinterpret_cast<const int*>(p.mel_table)[i];
    for (int i = tid; i < kWarps * kRowFloats; i += kThreads) scr_all[i] = 0.0f;   // pad columns stay finite
    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        fence_mbar_init();
        s_pend[0] = 0;
        s_pend[1] = 0;
    }
    __syncthreads();

    const int T = p.T, hop = p.hop, frames = p.frames, n_mels = p.n_mels;
    const size_t clip_elems = static_cast<size_t>(n_mels) * frames;

    // ---- staging of one item into buffer `buf` -------------------------------------------------
    // bulk part: [e_lo, e_lo + cnt) of the tile is src[j0 + e_lo ...] verbatim (plain clips only)
    auto bulk_range = [&](const ClipCtx* __restrict__ c, int tile_, int& e_lo, int& cnt) {
        e_lo = 0; cnt = 0;
        if (!p.use_tma || !c->plain) return;
        const int tf = tile_ * TILE_F;
        const int nf = (frames - tf) < TILE_F ? (frames - tf) : TILE_F;
        const int need = (nf - 1) * hop + NFFT;
        const int j0 = tf * hop - HALF;
        const int lo = j0 < 0 ? -j0 : 0;
        int hi = c->lc - j0;
        if (hi > need) hi = need;
        if (hi <= lo) return;
        if ((reinterpret_cast<uintptr_t>(c->src + j0 + lo) & 15u) != 0 || (lo & 3) != 0) return;
        e_lo = lo;
        cnt = (hi - lo) & ~3;
    };
    auto stage_bulk = [&](const ClipCtx* __restrict__ c, int tile_, int buf) {   // ONE thread
        int e_lo, cnt;
        bulk_range(c, tile_, e_lo, cnt);
        if (cnt != 0) {
            const int j0 = tile_ * TILE_F * hop - HALF;
            fence_proxy_async();
            mbar_expect_tx(&mbar[buf], static_cast<uint32_t>(cnt) * 4u);
            bulk_g2s(sbuf + static_cast<size_t>(buf) * p.ns + e_lo, c->src + j0 + e_lo, static_cast<uint32_t>(cnt) * 4u,
                     &mbar[buf]);
        }
        s_pend[buf] = (cnt != 0);
    };
    auto stage_gather = [&](const ClipCtx* __restrict__ cc, int tile_, int buf) {   // all threads
        const ClipCtx c = *cc;
        int e_lo, cnt;
        bulk_range(cc, tile_, e_lo, cnt);
        const int tf = tile_ * TILE_F;
        const int nf = (frames - tf) < TILE_F ? (frames - tf) : TILE_F;
        const int need = (nf - 1) * hop + NFFT;
        const int j0 = tf * hop - HALF;
        float* __restrict__ sb = sbuf + static_cast<size_t>(buf) * p.ns;
        // reflect (torch.stft center=True) -> roll -> pad/crop -> gain, + noise; slots past `need`
        // feed only frames >= `frames` and are zeroed
        const int rest = p.ns - cnt;   // slots the bulk copy does not cover: [0, e_lo) and [e_lo + cnt, ns)
        for (int idx = tid; idx < rest; idx += kThreads) {
            const int e = idx < e_lo ? idx : idx + cnt;
            int j = j0 + e;
            if (j < 0) j = -j;
            else if (j >= T) j = 2 * (T - 1) - j;
            float v = 0.0f;
            if (e < need && j >= 0 && j < T) {
                int i = j - c.shift;               // torch.roll: out[j] = in[(j - shift) mod T]
                if (i < 0) i += T;
                else if (i >= T) i -= T;
                if (i < c.lc) v = __ldg(c.src + i) * c.gain;
                if (c.nscale != 0.0f) {
                    const float z = c.nz ? __ldg(c.nz + i) : philox_normal(c.seed, static_cast<uint32_t>(i));
                    v = fmaf(z, c.nscale, v);
                }
            }
            sb[e] = v;
        }
    };

    // ---- prologue: items 0 and 1 are staged before the loop; item it+2 is staged during item it ------
    const int clip0 = blockIdx.x, cstride = gridDim.x;
    if (tid == 0) {
        load_clip(p, clip0, &s_ctx[0]);
        if (p.n_tiles == 1 && n_items > 1) load_clip(p, clip0 + cstride, &s_ctx[1]);
        stage_bulk(&s_ctx[0], 0, 0);
        if (n_items > 1) stage_bulk(&s_ctx[p.n_tiles == 1 ? 1 : 0], p.n_tiles == 1 ? 0 : 1, 1);
    }
    __syncthreads();
    stage_gather(&s_ctx[0], 0, 0);
    if (n_items > 1) stage_gather(&s_ctx[p.n_tiles == 1 ? 1 : 0], p.n_tiles == 1 ? 0 : 1, 1);
    __syncthreads();

    uint32_t parity0 = 0, parity1 = 0;     // mbarrier phase per staging buffer (CTA-uniform)
    s_stat[tid] = make_double2(0.0, 0.0);  // this thread's running (sum, sum of squares) of the clip's dB values
    int tile = 0, ord = 0;                 // tile index and clip ordinal of item `it`

#pragma unroll 1
    for (int it = 0; it < n_items; ++it) {
        const int buf = it & 1;
        const int tf = tile * TILE_F;                  // first frame of the tile
        const int clip = clip0 + ord * cstride;
        const float* __restrict__ sb = sbuf + static_cast<size_t>(buf) * p.ns;
        if (s_pend[buf]) {
            mbar_wait(&mbar[buf], buf ? parity1 : parity0);
            if (buf) parity1 ^= 1u; else parity0 ^= 1u;
        }

#if !LM_SPLIT
        __syncthreads();   // (X) placed before the FFT: experiment baseline
#endif
        // ---- FFT part 1 (registers + reads of the staged samples only) ----------------------------------
        lm_f2 z[32];
        const bool active = LM_ALWAYS_ACTIVE ? true : ((NFFT == 2048) ? (tf + warp_ < frames) : (tf + 2 * warp_ < frames));
        if (LM_STAGGER_NS > 0) __nanosleep((warp_ >> 2) * LM_STAGGER_NS);
        if (active) {
            const int lane = launder(lane_), warp = launder(warp_);
            if (NFFT == 2048) {
                const float2* __restrict__ s2 = reinterpret_cast<const float2*>(sb + warp * hop);
                const float2* __restrict__ w2 = reinterpret_cast<const float2*>(s_win);
#if LM_WINCALC
                // hann[64 n1 + 2 lane + j] = 0.5 - 0.5 cos(2 pi n1/32 + phi_j): angle addition with the
                // per-lane (cos phi_j, sin phi_j) pairs; the constants in n1 are literals
                const float4 wp4 = s_wph[lane];
                const lm_f2 CP = lm_pack(wp4.x, wp4.y), SP = lm_pack(wp4.z, wp4.w);
                (void)w2;
#pragma unroll
                for (int n1 = 0; n1 < 32; ++n1) {
                    const float2 v = s2[32 * n1 + lane];
                    const float ca = kWinCa[n1], sa = kWinSa[n1];
                    const lm_f2 w = lm_fma2(lm_bcast(ca), CP, lm_fma2(lm_bcast(sa), SP, lm_bcast(0.5f)));
                    z[n1] = lm_mul2(lm_pack(v.x, v.y), w);
                }
#else
#pragma unroll
                for (int n1 = 0; n1 < 32; ++n1) {
                    const float2 v = s2[32 * n1 + lane];
                    const float2 w = w2[32 * n1 + lane];
                    z[n1] = lm_mul2(lm_pack(v.x, v.y), lm_pack(w.x, w.y));
                }
#endif
            } else {
                // n_fft = 1024: two frames per warp as one complex signal z = a + i b
                const float* __restrict__ sa = sb + (2 * warp) * hop;
                const float* __restrict__ sbb = sa + hop;
#pragma unroll
                for (int n1 = 0; n1 < 32; ++n1) {
                    const float w = s_win[32 * n1 + lane];
                    z[n1] = lm_mul2(lm_pack(sa[32 * n1 + lane], sbb[32 * n1 + lane]), lm_bcast(w));
                }
            }
            warp_cfft1024_part1(z, s_tw, lane);
        }
#if LM_SPLIT
        __syncthreads();   // (X) every warp is done with the mel phase of the previous item (rows are
                           //     free) and with this item's staged samples (buffer `buf` is free)
#endif

        // ---- FFT part 2: transpose, second FFT, untangle -> 4|X|^2 in the warp's row ----------------------------
        if (active) {
            const int lane = launder(lane_), warp = launder(warp_);
            float* const scr = scr_all + warp * kRowFloats;
            float xr[32], xi[32];
            warp_cfft1024_part2(z, xr, xi, scr, s_tw, lane);
            const int srcl = (32 - lane) & 31;
            const bool l0 = (lane == 0);
            if (NFFT == 2048) {
                // Real-FFT untangle.  Bin k = lane + 32*k2 pairs with 1024-k, which lives in lane
                // (32-lane)&31 at slot 31-k2 (lane 0: slot (32-k2)&31) and comes over by warp shuffle.
                // Two pairs per packed op: k2 = i (lo half) and k2 = i+16 (hi half), i = 0..7; their
                // partners are the other lane's slots 31-i and 15-i, and the hi twiddle is the lo one
                // turned by pi/2: (c, s)(k+512) = (-s, c)(k).  The row was last read inside part 2
                // (followed by __syncwarp): 4|X|^2 goes straight into it, bin-major.
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float s_lr = l0 ? xr[(32 - i) & 31] : xr[31 - i];
                    const float s_li = l0 ? xi[(32 - i) & 31] : xi[31 - i];
                    const float s_hr = l0 ? xr[(16 - i) & 31] : xr[15 - i];
                    const float s_hi = l0 ? xi[(16 - i) & 31] : xi[15 - i];
                    const float b_lr = __shfl_sync(0xffffffffu, s_lr, srcl);
                    const float b_li = __shfl_sync(0xffffffffu, s_li, srcl);
                    const float b_hr = __shfl_sync(0xffffffffu, s_hr, srcl);
                    const float b_hi = __shfl_sync(0xffffffffu, s_hi, srcl);
                    const lm_f2 Ar = lm_pack(xr[i], xr[i + 16]), Ai = lm_pack(xi[i], xi[i + 16]);
                    const lm_f2 Br = lm_pack(b_lr, b_hr), Bi = lm_pack(b_li, b_hi);
                    const lm_f2 Er = lm_add2(Ar, Br), Ei = lm_sub2(Ai, Bi), Or = lm_add2(Ai, Bi), Oi = lm_sub2(Br, Ar);
                    const float2 cs = s_utw[lane + 32 * i];
                    const lm_f2 C = lm_pack(cs.x, -cs.y), S = lm_pack(cs.y, cs.x), nS = lm_pack(-cs.y, -cs.x);
                    const lm_f2 Tr = lm_fma2(C, Or, lm_mul2(S, Oi));
                    const lm_f2 Ti = lm_fma2(C, Oi, lm_mul2(nS, Or));
                    const lm_f2 Ur = lm_add2(Er, Tr), Ui = lm_add2(Ei, Ti), Vr = lm_sub2(Er, Tr), Vi = lm_sub2(Ei, Ti);
                    const lm_f2 PU = lm_fma2(Ur, Ur, lm_mul2(Ui, Ui)), PV = lm_fma2(Vr, Vr, lm_mul2(Vi, Vi));
                    scr[lane + 32 * i] = lm_lo(PU);
                    scr[lane + 32 * i + 512] = lm_hi(PU);
                    scr[1024 - lane - 32 * i] = lm_lo(PV);
                    scr[512 - lane - 32 * i] = lm_hi(PV);
                }
                {   // lane 0 only: bins 256 and 768 pair with each other (slots 8 and 24), twiddle pi/4
                    const float ar = xr[8], ai = xi[8], br = xr[24], bi = xi[24];
                    const float er = ar + br, ei = ai - bi, orr = ai + bi, oi = br - ar;
                    const float c = 0.70710678118654752440f;
                    const float tr = c * (orr + oi), ti = c * (oi - orr);
                    const float ur = er + tr, ui = ei + ti, vr = er - tr, vi = ei - ti;
                    if (l0) {
                        scr[256] = fmaf(ur, ur, ui * ui);
                        scr[768] = fmaf(vr, vr, vi * vi);
                    }
                }
            } else {
                // A = Z[k], B = Z[1024-k]:  |Xa|^2 = |A + conj B|^2 / 4, |Xb|^2 = |A - conj B|^2 / 4
                const float z16r = xr[16], z16i = xi[16];
                float* Pa = scr;
                float* Pb = scr + kPbOff;
#pragma unroll
                for (int k2 = 0; k2 < 16; ++k2) {
                    float br = __shfl_sync(0xffffffffu, xr[31 - k2], srcl);
                    float bi = __shfl_sync(0xffffffffu, xi[31 - k2], srcl);
                    if (l0) { br = xr[(32 - k2) & 31]; bi = xi[(32 - k2) & 31]; }
                    const float ar = xr[k2], ai = xi[k2];
                    const float ur = ar + br, ui = ai - bi, vr = ar - br, vi = ai + bi;
                    Pa[lane + 32 * k2] = fmaf(ur, ur, ui * ui);
                    Pb[lane + 32 * k2] = fmaf(vr, vr, vi * vi);
                }
                if (l0) {   // k = 512 pairs with itself: A = B
                    Pa[512] = 4.0f * z16r * z16r;
                    Pb[512] = 4.0f * z16i * z16i;
                }
            }
        }
        // ---- item it+2: its TMA part goes into the buffer consumed before (X).  Issued here, where no
        //      FFT registers are live, by one thread; the clip's context slot is filled when its first
        //      tile comes up. ------------------------------------------------------------------------------
        const bool has2 = (it + 2 < n_items);
        int tile2 = tile + 2, ord2 = ord;
        while (tile2 >= p.n_tiles) { tile2 -= p.n_tiles; ++ord2; }
        if (has2 && tid == 0) {
            if (tile2 == 0) load_clip(p, clip0 + ord2 * cstride, &s_ctx[ord2 & 3]);
            stage_bulk(&s_ctx[ord2 & 3], tile2, buf);
        }
        __syncthreads();   // (Y) all power rows of the tile are in shared memory

        // ---- mel phase: tensor cores, one 8-mel column block per warp ---------------------------------
        {
            const int lane = launder(lane_), warp = launder(warp_);
            const int g = lane >> 2, tg = lane & 3;
            const int nf = (frames - tf) < TILE_F ? (frames - tf) : TILE_F;
            const ClipCtx* __restrict__ cx = &s_ctx[ord & 3];
            double s_acc = s_stat[tid].x, q_acc = s_stat[tid].y;
            float* __restrict__ out = p.out_norm + static_cast<size_t>(clip) * clip_elems;
            float* __restrict__ odb = (EXTRA_OUT && p.out_db) ? p.out_db + static_cast<size_t>(clip) * clip_elems : nullptr;
            float* __restrict__ omp = (EXTRA_OUT && p.out_melpow) ? p.out_melpow + static_cast<size_t>(clip) * clip_elems : nullptr;
#pragma unroll 1
            for (int slot = 0; slot < 2; ++slot) {
                const int mt = s_tab->warp_tile[warp][slot];
                if (mt < 0) break;
                const int kb = s_tab->kb[mt], ndk = s_tab->ndk[mt];
                const float4* __restrict__ wp = s_melw + static_cast<size_t>(s_tab->off[mt]) * 32 + lane;
#pragma unroll 1
                for (int mb = 0; mb < G::MT; ++mb) {
                    if (mb * 16 >= nf) break;
                    // frame row f of the tile lives in warp row f/FPW (+ kPbOff for the odd frame)
                    const int fa = mb * 16 + g, fb_ = fa + 8;
                    const float* __restrict__ ra =
                        scr_all + (fa / G::FPW) * kRowFloats + (fa % G::FPW) * kPbOff + kb + 4 * tg;
                    const float* __restrict__ rb =
                        scr_all + (fb_ / G::FPW) * kRowFloats + (fb_ % G::FPW) * kPbOff + kb + 4 * tg;
                    float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f}, acc2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
                    for (int d = 0; d < ndk; ++d) {
                        const float4 pa = *reinterpret_cast<const float4*>(ra + 16 * d);
                        const float4 pb = *reinterpret_cast<const float4*>(rb + 16 * d);
                        const float4 w = wp[32 * d];
                        const uint32_t ah0 = tf32_hi(pa.x), ah1 = tf32_hi(pa.y), ah2 = tf32_hi(pa.z), ah3 = tf32_hi(pa.w);
                        const uint32_t bh0 = tf32_hi(pb.x), bh1 = tf32_hi(pb.y), bh2 = tf32_hi(pb.z), bh3 = tf32_hi(pb.w);
                        const uint32_t wh0 = tf32_hi(w.x), wh1 = tf32_hi(w.y), wh2 = tf32_hi(w.z), wh3 = tf32_hi(w.w);
                        // k-step 1: logical k = tg -> bin 4tg, k = tg+4 -> bin 4tg+1; k-step 2: bins 4tg+2, 4tg+3
                        mma_tf32(acc0, ah0, bh0, ah1, bh1, wh0, wh1);
                        mma_tf32(acc1, tf32_lo(pa.x, ah0), tf32_lo(pb.x, bh0), tf32_lo(pa.y, ah1), tf32_lo(pb.y, bh1), wh0, wh1);
                        mma_tf32(acc2, ah0, bh0, ah1, bh1, tf32_lo(w.x, wh0), tf32_lo(w.y, wh1));
                        mma_tf32(acc0, ah2, bh2, ah3, bh3, wh2, wh3);
                        mma_tf32(acc1, tf32_lo(pa.z, ah2), tf32_lo(pb.z, bh2), tf32_lo(pa.w, ah3), tf32_lo(pb.w, bh3), wh2, wh3);
                        mma_tf32(acc2, ah2, bh2, ah3, bh3, tf32_lo(w.z, wh2), tf32_lo(w.w, wh3));
                    }
                    // epilogue: c0:(frame g, mel 2tg) c1:(g, 2tg+1) c2:(g+8, 2tg) c3:(g+8, 2tg+1)
                    const int m0 = mt * 8 + 2 * tg, fl0 = mb * 16 + g;
                    const int tt0 = tf + fl0, tt1 = tt0 + 8;
                    const int cf0 = cx->f0, cf1 = cx->f1, ct0 = cx->t0, ct1 = cx->t1;
                    const bool mk_m0 = (m0 >= cf0) && (m0 < cf1), mk_m1 = (m0 + 1 >= cf0) && (m0 + 1 < cf1);
                    const bool mk_t0 = (tt0 >= ct0) && (tt0 < ct1), mk_t1 = (tt1 >= ct0) && (tt1 < ct1);
                    const bool ok_f0 = fl0 < nf, ok_f1 = fl0 + 8 < nf;
                    const bool ok_m0 = m0 < n_mels, ok_m1 = m0 + 1 < n_mels;
                    const int o00 = m0 * frames + tt0;
                    float ls = 0.0f, lq = 0.0f;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const bool ok = ((c & 2) ? ok_f1 : ok_f0) && ((c & 1) ? ok_m1 : ok_m0);
                        if (ok) {
                            const bool masked = ((c & 1) ? mk_m1 : mk_m0) || ((c & 2) ? mk_t1 : mk_t0);
                            const float mp = acc0[c] + (acc1[c] + acc2[c]);
                            float v = (mp <= p.amin) ? p.floor_db : fmaf(p.db_scale, __log2f(mp), -p.db_offset);
                            if (masked) v = 0.0f;
                            const int o = o00 + ((c & 1) ? frames : 0) + ((c & 2) ? 8 : 0);
                            out[o] = v;
                            if (EXTRA_OUT) {
                                if (odb) odb[o] = v;
                                if (omp) omp[o] = mp;
                            }
                            if (LM_F32_TILE_STATS) {
                                ls += v;
                                lq = fmaf(v, v, lq);
                            } else {
                                const double dv = static_cast<double>(v);
                                s_acc += dv;
                                q_acc = fma(dv, dv, q_acc);
                            }
                        }
                    }
                    if (LM_F32_TILE_STATS) {
                        s_acc += static_cast<double>(ls);
                        q_acc += static_cast<double>(lq);
                    }
                }
            }
            s_stat[tid] = make_double2(s_acc, q_acc);
        }

        // ---- gather part of item it+2 (its TMA part is already in flight) -----------------------------
        if (has2) stage_gather(&s_ctx[ord2 & 3], tile2, buf);

        // ---- per-clip normalisation --------------------------------------------------------------------
        if (tile + 1 == p.n_tiles) {
            if (p.normalize) {
                float* __restrict__ out = p.out_norm + static_cast<size_t>(clip) * clip_elems;
                double s_acc = s_stat[tid].x, q_acc = s_stat[tid].y;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    s_acc += __shfl_xor_sync(0xffffffffu, s_acc, o);
                    q_acc += __shfl_xor_sync(0xffffffffu, q_acc, o);
                }
                if (lane_ == 0) { red[warp_] = s_acc; red[kWarps + warp_] = q_acc; }
                __syncthreads();   // also orders every thread's dB stores before the re-read below
                if (tid == 0) {
                    double s = 0.0, q = 0.0;
                    for (int w = 0; w < kWarps; ++w) { s += red[w]; q += red[kWarps + w]; }
                    const double n = static_cast<double>(clip_elems);
                    const double mean = s / n;
                    double var = (q - s * mean) / (n - 1.0);   // unbiased, as torch.std
                    if (!(var > 0.0)) var = 0.0;
                    bcast[0] = static_cast<float>(mean);
                    bcast[1] = static_cast<float>(sqrt(var)) + p.norm_eps;
                }
                __syncthreads();
                const float mean = bcast[0], inv = 1.0f / bcast[1];
                float4* __restrict__ o4 = reinterpret_cast<float4*>(out);
                const int n4 = (reinterpret_cast<uintptr_t>(out) & 15u) == 0 ? static_cast<int>(clip_elems >> 2) : 0;
                for (int i = tid; i < n4; i += kThreads) {
                    float4 v = __ldcg(o4 + i);
                    v.x = (v.x - mean) * inv;
                    v.y = (v.y - mean) * inv;
                    v.z = (v.z - mean) * inv;
                    v.w = (v.w - mean) * inv;
                    o4[i] = v;
                }
                for (int i = (n4 << 2) + tid; i < static_cast<int>(clip_elems); i += kThreads)
                    out[i] = (__ldcg(out + i) - mean) * inv;
            }
            s_stat[tid] = make_double2(0.0, 0.0);
            tile = 0;
            ++ord;
        } else {
            ++tile;
        }
    }
}

}  // namespace lm
