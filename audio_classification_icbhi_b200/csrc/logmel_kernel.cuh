// logmel_kernel.cuh -- the fused log-mel kernel for sm_100a (B200).
//
// One persistent CTA per SM (16 warps).  A CTA takes whole clips (blockIdx.x, += gridDim.x) and,
// per clip, walks tiles of 16 frames (32 for n_fft = 1024):
//
//   stage   the tile's samples -> shared memory, once per sample although every sample feeds
//           4 frames.  Interior tiles of un-augmented clips are one cp.async.bulk (TMA, 1-D)
//           issued a tile ahead into the other buffer; edge tiles / augmented clips go through
//           the gather path that does pad/crop (R/src/data/preprocessing.py:70-83), noise and
//           roll (:85-93) and torch.stft's reflect padding as index arithmetic.
//   frame   one warp = one frame.  x Hann, 2048 real -> 1024 complex points held as
//           32 registers/lane; register radix-32 FFT over the lane-local index, twiddle,
//           32x32 transpose through a private shared-memory scratch, second radix-32 FFT,
//           then the real-FFT untangle with the partner bin fetched by warp shuffle;
//           |X|^2 lands in the scratch (never in HBM).
//           (TA/functional/functional.py:123-145: torch.stft + abs().pow(2))
//   mel+dB  sparse banded filterbank rows from shared memory, 10*log10(max(x, amin))
//           (TA/transforms/_transforms.py:417, TA/functional/functional.py:390-391)
//   flush   the [n_mels x 16] dB tile is written with the SpecAugment intervals applied
//           (TA/functional/functional.py:939-953) and fp64 sum / sum-of-squares are kept.
//   norm    when the clip is finished the same CTA re-reads its (L2-resident) dB block and
//           writes (x - mean) / (std + eps)  (R/src/data/preprocessing.py:111-116).
//
// Algorithmic HBM bytes per clip: 4*min(len, T) read + 4*n_mels*frames written.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/logmel_b200.h"
#include "fft_gen.cuh"

namespace lm {

constexpr int kWarps = 16;
constexpr int kThreads = kWarps * 32;
constexpr int kScrPitch = 36;                 // 32 + 4: rows 16 B aligned, LDS.128 conflict-free
constexpr int kScrFloats = 32 * kScrPitch;    // 1152 >= 1025 power bins
constexpr int kMaxMelW = 6144;                // filterbank non-zeros kept on chip

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
    int mel_nnz;     // floats in melw (padded to 4)
    int use_tma;
    float db_mult, amin, db_offset, floor_db, norm_eps;
    const float* __restrict__ window;   // [NFFT]
    const float2* __restrict__ tw;      // [32*32]  W1024^(n2*k1) = (cos, -sin), index k1*32+n2
    const float2* __restrict__ utw;     // [512]    (cos, sin)(2 pi k / 2048)
    const float* __restrict__ melw;     // concatenated filter rows, pre-scaled by 1/4
    const int* __restrict__ mel_meta;   // [3*n_mels]: start bin, length, offset into melw
};

// ---------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk copy (TMA) global -> shared
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
// 1024-point complex FFT of one warp: lane n2 holds z[32*n1 + n2] in slot n1 on entry,
// lane k1 holds Z[k1 + 32*k2] in slot k2 on exit.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_cfft1024(float (&xr)[32], float (&xi)[32], float* __restrict__ scr,
                                              const float2* __restrict__ tw, int lane) {
    lm_fft32(xr, xi);
#pragma unroll
    for (int k1 = 1; k1 < 32; ++k1) {
        const float2 w = tw[k1 * 32 + lane];   // (cos, -sin)
        const float r = xr[k1], i = xi[k1];
        xr[k1] = fmaf(r, w.x, -i * w.y);
        xi[k1] = fmaf(r, w.y, i * w.x);
    }
    // 32x32 transpose, real plane then imaginary plane
#pragma unroll
    for (int k1 = 0; k1 < 32; ++k1) scr[k1 * kScrPitch + lane] = xr[k1];
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(scr + lane * kScrPitch + 4 * q);
        xr[4 * q] = v.x; xr[4 * q + 1] = v.y; xr[4 * q + 2] = v.z; xr[4 * q + 3] = v.w;
    }
    __syncwarp();
#pragma unroll
    for (int k1 = 0; k1 < 32; ++k1) scr[k1 * kScrPitch + lane] = xi[k1];
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(scr + lane * kScrPitch + 4 * q);
        xi[4 * q] = v.x; xi[4 * q + 1] = v.y; xi[4 * q + 2] = v.z; xi[4 * q + 3] = v.w;
    }
    __syncwarp();
    lm_fft32(xr, xi);
}

// mel rows + dB for one frame whose (4x) power spectrum is in P
__device__ __forceinline__ void mel_db_frame(const KParams& p, const float* __restrict__ P,
                                             const float* __restrict__ melw, const int* __restrict__ meta,
                                             float* __restrict__ dbt, int pitch, int f_local, int lane,
                                             size_t out_base /* clip*n_mels*frames + t */, bool write_pow) {
    for (int m = lane; m < p.n_mels; m += 32) {
        const int st = meta[m], ln = meta[p.n_mels + m], of = meta[2 * p.n_mels + m];
        float acc = 0.0f;
        for (int i = 0; i < ln; ++i) acc = fmaf(melw[of + i], P[st + i], acc);
        const float db = (acc <= p.amin) ? p.floor_db : fmaf(p.db_mult, log10f(acc), -p.db_offset);
        dbt[m * pitch + f_local] = db;
        if (write_pow) p.out_melpow[out_base + static_cast<size_t>(m) * p.frames] = acc;
    }
}

template <int NFFT>
struct Geo {
    static constexpr int FPW = (NFFT == 2048) ? 1 : 2;   // frames per warp pass
    static constexpr int TILE_F = kWarps * FPW;
    static constexpr int PITCH = TILE_F + 1;
};

// Shared-memory carve-up, shared by host (size) and device (pointers).
template <int NFFT>
struct Smem {
    static __host__ __device__ size_t align16(size_t x) { return (x + 15) & ~size_t(15); }
    size_t o_bar, o_red, o_sbuf, o_scr, o_dbt, o_win, o_tw, o_utw, o_melw, o_meta, total;
    __host__ __device__ Smem(int ns, int n_mels, int mel_nnz) {
        size_t o = 0;
        o_bar = o; o += 16;
        o_red = o; o += sizeof(double) * 2 * kWarps + 16;
        o = align16(o);
        o_sbuf = o; o += sizeof(float) * 2 * static_cast<size_t>(ns);
        o_scr = o; o += sizeof(float) * kWarps * kScrFloats;
        o_dbt = o; o += align16(sizeof(float) * static_cast<size_t>(n_mels) * Geo<NFFT>::PITCH);
        o_win = o; o += sizeof(float) * NFFT;
        o_tw = o; o += sizeof(float2) * 1024;
        o_utw = o; o += (NFFT == 2048) ? sizeof(float2) * 512 : 0;
        o_melw = o; o += align16(sizeof(float) * static_cast<size_t>(mel_nnz));
        o_meta = o; o += align16(sizeof(int) * 3 * static_cast<size_t>(n_mels));
        total = o;
    }
};

template <int NFFT>
__global__ void __launch_bounds__(kThreads, 1) logmel_kernel(const KParams p) {
    using G = Geo<NFFT>;
    constexpr int TILE_F = G::TILE_F;
    constexpr int PITCH = G::PITCH;
    constexpr int HALF = NFFT / 2;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const Smem<NFFT> L(p.ns, p.n_mels, p.mel_nnz);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw + L.o_bar);
    double* red = reinterpret_cast<double*>(smem_raw + L.o_red);
    float* bcast = reinterpret_cast<float*>(smem_raw + L.o_red + sizeof(double) * 2 * kWarps);
    float* sbuf = reinterpret_cast<float*>(smem_raw + L.o_sbuf);
    float* scr_all = reinterpret_cast<float*>(smem_raw + L.o_scr);
    float* dbt = reinterpret_cast<float*>(smem_raw + L.o_dbt);
    float* s_win = reinterpret_cast<float*>(smem_raw + L.o_win);
    float2* s_tw = reinterpret_cast<float2*>(smem_raw + L.o_tw);
    float2* s_utw = reinterpret_cast<float2*>(smem_raw + L.o_utw);
    float* s_melw = reinterpret_cast<float*>(smem_raw + L.o_melw);
    int* s_meta = reinterpret_cast<int*>(smem_raw + L.o_meta);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* scr = scr_all + warp * kScrFloats;

    // ---- constants -> shared memory, once per (persistent) CTA -----------------------------
    for (int i = tid; i < NFFT; i += kThreads) s_win[i] = p.window[i];
    for (int i = tid; i < 1024; i += kThreads) s_tw[i] = p.tw[i];
    if (NFFT == 2048)
        for (int i = tid; i < 512; i += kThreads) s_utw[i] = p.utw[i];
    for (int i = tid; i < p.mel_nnz; i += kThreads) s_melw[i] = p.melw[i];
    for (int i = tid; i < 3 * p.n_mels; i += kThreads) s_meta[i] = p.mel_meta[i];
    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();

    uint32_t parity0 = 0, parity1 = 0;     // mbarrier phase per staging buffer (CTA-uniform)
    bool pending0 = false, pending1 = false;

    const int T = p.T, hop = p.hop, frames = p.frames, n_mels = p.n_mels;
    const size_t clip_elems = static_cast<size_t>(n_mels) * frames;

#pragma unroll 1
    for (int clip = blockIdx.x; clip < p.B; clip += gridDim.x) {
        // ---- per-clip scalars (uniform) ---------------------------------------------------
        const long long off = p.offset[clip];
        const int len = p.length[clip];
        const int crop = len > T ? (len - T) / 2 : 0;      // centre crop
        const int lc = len < T ? len : T;                  // valid samples after pad/crop
        const float* __restrict__ src = p.wave + off + crop;
        int shift = 0, f0 = 0, f1 = 0, t0m = 0, t1m = 0;
        float nscale = 0.0f, gain = 1.0f;
        uint64_t seed = 0;
        if (p.aug != nullptr) {
            const lm_aug a = p.aug[clip];
            shift = a.shift; nscale = a.noise_scale; gain = a.gain;
            f0 = a.f0; f1 = a.f1; t0m = a.t0; t1m = a.t1; seed = a.seed;
        }
        const float* __restrict__ nz = (p.noise != nullptr && nscale != 0.0f)
                                           ? p.noise + static_cast<size_t>(clip) * T : nullptr;
        const bool plain = (shift == 0) && (nscale == 0.0f) && (gain == 1.0f);
        float* __restrict__ out = p.out_norm + static_cast<size_t>(clip) * clip_elems;
        float* __restrict__ odb = p.out_db ? p.out_db + static_cast<size_t>(clip) * clip_elems : nullptr;

        double s_acc = 0.0, q_acc = 0.0;

        auto tile_need = [&](int tile) {   // samples the tile's valid frames touch
            const int tf = tile * TILE_F;
            const int nf = (frames - tf) < TILE_F ? (frames - tf) : TILE_F;
            return (nf - 1) * hop + NFFT;
        };
        auto tile_contig = [&](int tile) { // all of them plain interior samples?
            const int j0 = tile * TILE_F * hop - HALF;
            const int need4 = (tile_need(tile) + 3) & ~3;
            return plain && j0 >= 0 && (j0 + need4) <= lc;
        };
        auto tile_tma_ok = [&](int tile) {
            if (!p.use_tma || !tile_contig(tile)) return false;
            const int j0 = tile * TILE_F * hop - HALF;
            return (reinterpret_cast<uintptr_t>(src + j0) & 15u) == 0;
        };
        auto issue_tma = [&](int tile, int buf) {  // one thread
            const int j0 = tile * TILE_F * hop - HALF;
            const uint32_t bytes = static_cast<uint32_t>(((tile_need(tile) + 3) & ~3) * 4);
            fence_proxy_async();
            mbar_expect_tx(&mbar[buf], bytes);
            bulk_g2s(sbuf + static_cast<size_t>(buf) * p.ns, src + j0, bytes, &mbar[buf]);
        };

#pragma unroll 1
        for (int tile = 0; tile < p.n_tiles; ++tile) {
            const int buf = tile & 1;
            const int tf = tile * TILE_F;                  // first frame of the tile
            const int j0 = tf * hop - HALF;                // first sample (reflect-padded domain)
            float* __restrict__ sb = sbuf + static_cast<size_t>(buf) * p.ns;
            const int need = tile_need(tile);

            // ---- stage --------------------------------------------------------------------
            if (tile_tma_ok(tile)) {
                bool& pend = buf ? pending1 : pending0;
                uint32_t& par = buf ? parity1 : parity0;
                if (!pend && tid == 0) issue_tma(tile, buf);
                mbar_wait(&mbar[buf], par);
                par ^= 1u;
                pend = false;
            } else if (tile_contig(tile)) {
                const float* __restrict__ g = src + j0;
                if ((reinterpret_cast<uintptr_t>(g) & 15u) == 0) {
                    const int n4 = (need + 3) >> 2;
                    for (int e = tid; e < n4; e += kThreads)
                        reinterpret_cast<float4*>(sb)[e] = __ldg(reinterpret_cast<const float4*>(g) + e);
                } else {
                    for (int e = tid; e < need; e += kThreads) sb[e] = __ldg(g + e);
                }
            } else {
                // gather: reflect (torch.stft center=True) -> roll -> pad/crop -> gain, + noise
                // (slots past `need` feed only frames >= `frames`; zero them so the two-frame
                //  n_fft=1024 packing never mixes stale shared memory into a valid frame)
                for (int e = tid; e < p.ns; e += kThreads) {
                    int j = j0 + e;
                    if (j < 0) j = -j;
                    else if (j >= T) j = 2 * (T - 1) - j;
                    float v = 0.0f;
                    if (e < need && j >= 0 && j < T) {
                        int i = j - shift;               // torch.roll: out[j] = in[(j - shift) mod T]
                        if (i < 0) i += T;
                        else if (i >= T) i -= T;
                        if (i < lc) v = __ldg(src + i) * gain;
                        if (nscale != 0.0f) {
                            const float z = nz ? __ldg(nz + i) : philox_normal(seed, static_cast<uint32_t>(i));
                            v = fmaf(z, nscale, v);
                        }
                    }
                    sb[e] = v;
                }
            }
            __syncthreads();   // (A) tile staged; previous flush finished

            // ---- prefetch the next tile with TMA while this one is computed ------------------
            if (tile + 1 < p.n_tiles && tile_tma_ok(tile + 1)) {
                if (tid == 0) issue_tma(tile + 1, buf ^ 1);
                if (buf) pending0 = true; else pending1 = true;
            }

            // ---- frames ---------------------------------------------------------------------
            if (NFFT == 2048) {
                const int t = tf + warp;
                if (t < frames) {
                    float xr[32], xi[32];
                    const float2* __restrict__ s2 = reinterpret_cast<const float2*>(sb + warp * hop);
                    const float2* __restrict__ w2 = reinterpret_cast<const float2*>(s_win);
#pragma unroll
                    for (int n1 = 0; n1 < 32; ++n1) {
                        const float2 v = s2[32 * n1 + lane];
                        const float2 w = w2[32 * n1 + lane];
                        xr[n1] = v.x * w.x;
                        xi[n1] = v.y * w.y;
                    }
                    warp_cfft1024(xr, xi, scr, s_tw, lane);
                    // real-FFT untangle: pair (k, 1024-k), partner lane (32-lane)&31 via shuffle
                    const int srcl = (32 - lane) & 31;
                    const float z16r = xr[16], z16i = xi[16];
                    // scr was last read inside warp_cfft1024 (followed by __syncwarp): the
                    // (4x) power spectrum goes straight into it, bin-major.
#pragma unroll
                    for (int k2 = 0; k2 < 16; ++k2) {
                        float br = __shfl_sync(0xffffffffu, xr[31 - k2], srcl);
                        float bi = __shfl_sync(0xffffffffu, xi[31 - k2], srcl);
                        if (lane == 0) { br = xr[(32 - k2) & 31]; bi = xi[(32 - k2) & 31]; }
                        const float ar = xr[k2], ai = xi[k2];
                        const float er = ar + br, ei = ai - bi, orr = ai + bi, oi = br - ar;
                        const float2 cs = s_utw[lane + 32 * k2];
                        const float tr = fmaf(cs.x, orr, cs.y * oi);
                        const float ti = fmaf(cs.x, oi, -cs.y * orr);
                        const float ur = er + tr, ui = ei + ti, vr = er - tr, vi = ei - ti;
                        scr[lane + 32 * k2] = fmaf(ur, ur, ui * ui);
                        scr[1024 - lane - 32 * k2] = fmaf(vr, vr, vi * vi);
                    }
                    if (lane == 0) scr[512] = 4.0f * fmaf(z16r, z16r, z16i * z16i);
                    __syncwarp();
                    mel_db_frame(p, scr, s_melw, s_meta, dbt, PITCH, warp, lane,
                                 static_cast<size_t>(clip) * clip_elems + t, p.out_melpow != nullptr);
                    __syncwarp();
                }
            } else {
                // n_fft = 1024: two frames per warp as one complex signal z = a + i b
                const int ta = tf + 2 * warp;
                if (ta < frames) {
                    float xr[32], xi[32];
                    const float* __restrict__ sa = sb + (2 * warp) * hop;
                    const float* __restrict__ sbb = sa + hop;
#pragma unroll
                    for (int n1 = 0; n1 < 32; ++n1) {
                        const float w = s_win[32 * n1 + lane];
                        xr[n1] = sa[32 * n1 + lane] * w;
                        xi[n1] = sbb[32 * n1 + lane] * w;
                    }
                    warp_cfft1024(xr, xi, scr, s_tw, lane);
                    // A = Z[k], B = Z[1024-k]:  |Xa|^2 = |A + conj B|^2 / 4, |Xb|^2 = |A - conj B|^2 / 4
                    const int srcl = (32 - lane) & 31;
                    const float z16r = xr[16], z16i = xi[16];
                    float* Pa = scr;
                    float* Pb = scr + 576;   // 513 bins each
#pragma unroll
                    for (int k2 = 0; k2 < 16; ++k2) {
                        float br = __shfl_sync(0xffffffffu, xr[31 - k2], srcl);
                        float bi = __shfl_sync(0xffffffffu, xi[31 - k2], srcl);
                        if (lane == 0) { br = xr[(32 - k2) & 31]; bi = xi[(32 - k2) & 31]; }
                        const float ar = xr[k2], ai = xi[k2];
                        const float ur = ar + br, ui = ai - bi, vr = ar - br, vi = ai + bi;
                        Pa[lane + 32 * k2] = fmaf(ur, ur, ui * ui);
                        Pb[lane + 32 * k2] = fmaf(vr, vr, vi * vi);
                    }
                    if (lane == 0) {   // k = 512 pairs with itself: A = B
                        Pa[512] = 4.0f * z16r * z16r;
                        Pb[512] = 4.0f * z16i * z16i;
                    }
                    __syncwarp();
                    mel_db_frame(p, Pa, s_melw, s_meta, dbt, PITCH, 2 * warp, lane,
                                 static_cast<size_t>(clip) * clip_elems + ta, p.out_melpow != nullptr);
                    if (ta + 1 < frames)
                        mel_db_frame(p, Pb, s_melw, s_meta, dbt, PITCH, 2 * warp + 1, lane,
                                     static_cast<size_t>(clip) * clip_elems + ta + 1, p.out_melpow != nullptr);
                    __syncwarp();
                }
            }
            __syncthreads();   // (B) dB tile complete

            // ---- flush: masks, store, statistics -----------------------------------------------
            {
                const int nf = (frames - tf) < TILE_F ? (frames - tf) : TILE_F;
                for (int idx = tid; idx < n_mels * TILE_F; idx += kThreads) {
                    const int m = idx / TILE_F, f = idx - m * TILE_F;
                    if (f < nf) {
                        const int t = tf + f;
                        float v = dbt[m * PITCH + f];
                        if ((m >= f0 && m < f1) || (t >= t0m && t < t1m)) v = 0.0f;
                        const size_t o = static_cast<size_t>(m) * frames + t;
                        out[o] = v;
                        if (odb) odb[o] = v;
                        const double d = static_cast<double>(v);
                        s_acc += d;
                        q_acc = fma(d, d, q_acc);
                    }
                }
            }
        }   // tiles

        // ---- per-clip normalisation ------------------------------------------------------------
        if (p.normalize) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                s_acc += __shfl_xor_sync(0xffffffffu, s_acc, o);
                q_acc += __shfl_xor_sync(0xffffffffu, q_acc, o);
            }
            if (lane == 0) { red[warp] = s_acc; red[kWarps + warp] = q_acc; }
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
            const float mean = bcast[0], denom = bcast[1];
            float4* __restrict__ o4 = reinterpret_cast<float4*>(out);
            const int n4 = (reinterpret_cast<uintptr_t>(out) & 15u) == 0 ? static_cast<int>(clip_elems >> 2) : 0;
            for (int i = tid; i < n4; i += kThreads) {
                float4 v = __ldcg(o4 + i);
                v.x = __fdiv_rn(v.x - mean, denom);
                v.y = __fdiv_rn(v.y - mean, denom);
                v.z = __fdiv_rn(v.z - mean, denom);
                v.w = __fdiv_rn(v.w - mean, denom);
                o4[i] = v;
            }
            for (int i = (n4 << 2) + tid; i < static_cast<int>(clip_elems); i += kThreads)
                out[i] = __fdiv_rn(__ldcg(out + i) - mean, denom);
            __syncthreads();   // red/bcast reused by the next clip
        }
    }   // clips
}

}  // namespace lm
