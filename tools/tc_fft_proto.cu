// tc_fft_proto.cu -- the tensor-core STFT core (csrc/logmel_tc_core.cuh) on its own: numerics against a float64 DFT
// and cycles per frame per SM, before it goes behind the C ABI.  Two 8-warp groups per CTA, one CTA per SM, as in
// the log-mel kernel; a tile is 8 consecutive frames (hop 512) of one long synthetic signal.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o tools/tc_fft_proto tools/tc_fft_proto.cu
// Run:   tools/tc_fft_proto            (results: profiles/r2/tc_fft_proto.txt)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../audio_classification_icbhi_b200/csrc/logmel_tc_core.cuh"
#include "../audio_classification_icbhi_b200/csrc/logmel_tc_tables.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

using namespace lmtc;

constexpr int kNs = 7 * 512 + 2048;   // staged samples per tile

struct ProtoSmem {
    static constexpr size_t kG = 0;
    static constexpr size_t kTw1 = kG + 2 * kGBytes;
    static constexpr size_t kUtwC = kTw1 + kTw1Rows * 32 * 8;
    static constexpr size_t kUtwS = kUtwC + kUtwRows * kUtwPitch * 4;
    static constexpr size_t kWin = kUtwS + kUtwRows * kUtwPitch * 4;
    static constexpr size_t kMisc = kWin + 1024 * 4;            // mbarriers, tmem base, pscale
    static constexpr size_t kGroup0 = (kMisc + 256 + 1023) & ~size_t(1023);
    static constexpr size_t kSb = 0, kA2 = kSb + kNs * 4, kScr = kA2 + kA2Bytes, kGroupBytes = ((kScr + kScratchFloats * 4) + 1023) & ~size_t(1023);
    static constexpr size_t kTotal = kGroup0 + 2 * kGroupBytes;
};

__device__ __forceinline__ void group_bar(int group) { asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(kGroupThreads) : "memory"); }

__global__ void __launch_bounds__(512, 1) proto_kernel(const float* __restrict__ signal, int n_items, int write_p, float* __restrict__ p_out,
                                                      const uint4* __restrict__ g_img, const float* __restrict__ tw1, const float* __restrict__ utw_c,
                                                      const float* __restrict__ utw_s, const float* __restrict__ window,
                                                      unsigned long long* __restrict__ cycles, unsigned long long* __restrict__ phases = nullptr, int mode = 0) {
    extern __shared__ __align__(1024) uint8_t smem[];
    using L = ProtoSmem;
    const int tid = threadIdx.x, lane = tid & 31, group = tid >> 8, gtid = tid & 255, gw = gtid >> 5;
    for (int i = tid; i < 2 * kGBytes / 16; i += 512) reinterpret_cast<uint4*>(smem + L::kG)[i] = g_img[i];
    for (int i = tid; i < kTw1Rows * 64; i += 512) reinterpret_cast<float*>(smem + L::kTw1)[i] = tw1[i];
    for (int i = tid; i < kUtwRows * kUtwPitch; i += 512) {
        reinterpret_cast<float*>(smem + L::kUtwC)[i] = utw_c[i];
        reinterpret_cast<float*>(smem + L::kUtwS)[i] = utw_s[i];
    }
    for (int i = tid; i < 1024; i += 512) reinterpret_cast<float*>(smem + L::kWin)[i] = window[i];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kMisc);
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(smem + L::kMisc + 16);
    float* pscale = reinterpret_cast<float*>(smem + L::kMisc + 32) + group * 8;
    if (tid == 0) {
        mbar_init(s32(&bars[0]), 1);
        mbar_init(s32(&bars[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) tmem_alloc(s32(tmem_base_s), 512);
    uint8_t* gbase = smem + L::kGroup0 + group * L::kGroupBytes;
    float* sb = reinterpret_cast<float*>(gbase + L::kSb);
    for (int i = gtid; i < kA2Bytes / 4; i += 256) reinterpret_cast<float*>(gbase + L::kA2)[i] = 0.f;
    proxy_fence();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    Ctx cx;
    cx.tm = *tmem_base_s + group * kTmemColsPerGroup;
    cx.bar = s32(&bars[group]);
    cx.a2 = s32(gbase + L::kA2);
    cx.g_hi = make_desc(s32(smem + L::kG), 128, 1024);
    cx.g_lo = make_desc(s32(smem + L::kG + kGBytes), 128, 1024);
    cx.prow = reinterpret_cast<float*>(gbase + L::kA2);
    cx.scratch = reinterpret_cast<float*>(gbase + L::kScr);
    cx.pscale = pscale;
    cx.tw1 = reinterpret_cast<const float2*>(smem + L::kTw1);
    cx.utw_c = reinterpret_cast<const float*>(smem + L::kUtwC);
    cx.utw_s = reinterpret_cast<const float*>(smem + L::kUtwS);
    const float* s_win = reinterpret_cast<const float*>(smem + L::kWin);
    const bool leader = elect_one();
    uint32_t parity = 0;

    long long ph[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tl;
#define PH(i) do { const long long tn_ = clock64(); ph[i] += tn_ - tl; tl = tn_; } while (0)
    const long long t0 = clock64();
    tl = t0;
    const bool no_mma = (mode & 2) != 0;
    for (int item = blockIdx.x * 2 + group; item < n_items; item += 2 * gridDim.x) {
        if ((mode & 1) && group == 1) break;
        const float4* src = reinterpret_cast<const float4*>(signal + static_cast<size_t>(item) * 4096);
        for (int i = gtid; i < kNs / 4; i += 256) reinterpret_cast<float4*>(sb)[i] = __ldg(src + i);
        group_bar(group);
        PH(0);   // load + barrier
        b0_frame(cx, sb + gw * 512, s_win, gw, lane);
        PH(1);   // B0
        tc_fence_before();
        group_bar(group);
        if (gw == 0 && !no_mma) issue_stage1(cx, leader);
        PH(2);   // barrier + issue
        if (!no_mma) { mbar_wait(cx.bar, parity); parity ^= 1u; }
        tc_fence_after();
        PH(3);   // wait S1
        {
            uint32_t w[32];
            b1_half(cx, gw, lane, 0, w);
            b1_store(cx, gw, lane, w);
        }
        PH(4);   // B1a
        tc_fence_before();
        group_bar(group);
        if (gw == 0 && !no_mma) issue_stage2(cx, leader, 0);
        PH(5);   // barrier + issue
        {
            uint32_t w[32];
            b1_half(cx, gw, lane, 1, w);
            PH(6);   // B1b compute
            if (!no_mma) { mbar_wait(cx.bar, parity); parity ^= 1u; }   // MMA-A has read the operand buffer
            PH(7);   // wait S2A
            b1_store(cx, gw, lane, w);
        }
        tc_fence_before();
        group_bar(group);
        if (gw == 0 && !no_mma) issue_stage2(cx, leader, 1);
        PH(8);   // store + barrier + issue
        if (!no_mma) { mbar_wait(cx.bar, parity); parity ^= 1u; }
        tc_fence_after();
        PH(9);   // wait S2B
        b2_rows(cx, gw, lane);
        PH(10);  // B2
        tc_fence_before();
        group_bar(group);
        fixup(cx, gw, lane);
        group_bar(group);
        PH(11);  // fix-up + barriers
        if (write_p) {
            float* dst = p_out + static_cast<size_t>(item) * 8 * 1024;
            for (int i = gtid; i < 8 * 1024; i += 256) dst[i] = cx.prow[(i >> 10) * kPPitch + (i & 1023)];
        }
        group_bar(group);
    }
    const long long t1 = clock64();
    if (gtid == 0) cycles[blockIdx.x * 2 + group] = static_cast<unsigned long long>(t1 - t0);
    if (blockIdx.x == 0 && lane == 0 && (gw == 0 || gw == 5) && group == 0 && phases)
        for (int i = 0; i < 12; ++i) phases[(gw ? 12 : 0) + i] = static_cast<unsigned long long>(ph[i]);
    tc_fence_before();
    __syncthreads();
    if (tid < 32) tmem_free(*tmem_base_s, 512);
}

int main(int argc, char** argv) {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("device: %s, %d SMs\n", prop.name, sms);

    std::vector<uint8_t> g_hi, g_lo;
    lmtc_host::build_dft32(g_hi, g_lo);
    std::vector<uint8_t> g_img(g_hi);
    g_img.insert(g_img.end(), g_lo.begin(), g_lo.end());
    std::vector<float> tw1, utw_c, utw_s, window(2048);
    lmtc_host::build_tw1(tw1);
    lmtc_host::build_utw(utw_c, utw_s);
    for (int n = 0; n < 2048; ++n) window[n] = static_cast<float>(0.5 - 0.5 * std::cos(6.283185307179586476925286766559 * n / 2048.0));

    const int check_items = 6, perf_items = sms * 2 * 32;
    const size_t n_samp = static_cast<size_t>(perf_items) * 4096 + 2048;
    std::vector<float> sig(n_samp);
    std::mt19937 rng(1234);
    std::normal_distribution<float> nd(0.f, 0.1f);
    for (auto& v : sig) v = nd(rng);
    // item 1: quiet noise (1e-3 of the rest); item 2: a tone between bins plus weak noise; item 3: loud uniform; item 4: silence then a click
    for (int i = 4096; i < 4096 + 5632; ++i) sig[i] *= 1e-2f;
    for (int i = 2 * 4096 + 1536; i < 3 * 4096 + 1536; ++i) sig[i] = 0.5f * std::sin(0.0813f * i) + 1e-3f * sig[i];
    for (int i = 3 * 4096 + 1536; i < 4 * 4096 + 1536; ++i) sig[i] = 2.f * (static_cast<float>(rng() & 0xffff) / 65535.f) - 1.f;
    for (int i = 4 * 4096 + 1536; i < 5 * 4096 + 1536; ++i) sig[i] = 0.f;
    sig[4 * 4096 + 3000] = 0.8f;

    float *d_sig, *d_p, *d_tw1, *d_uc, *d_us, *d_win;
    uint4* d_g;
    unsigned long long* d_cyc;
    CK(cudaMalloc(&d_sig, n_samp * 4));
    CK(cudaMalloc(&d_p, static_cast<size_t>(check_items) * 8 * 1024 * 4));
    CK(cudaMalloc(&d_g, g_img.size()));
    CK(cudaMalloc(&d_tw1, tw1.size() * 4));
    CK(cudaMalloc(&d_uc, utw_c.size() * 4));
    CK(cudaMalloc(&d_us, utw_s.size() * 4));
    CK(cudaMalloc(&d_win, 2048 * 4));
    CK(cudaMalloc(&d_cyc, sms * 2 * 8));
    unsigned long long* d_ph;
    CK(cudaMalloc(&d_ph, 24 * 8));
    CK(cudaMemcpy(d_sig, sig.data(), n_samp * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_g, g_img.data(), g_img.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_tw1, tw1.data(), tw1.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_uc, utw_c.data(), utw_c.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_us, utw_s.data(), utw_s.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_win, window.data(), 2048 * 4, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(proto_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(ProtoSmem::kTotal)));
    printf("shared memory per CTA: %zu bytes\n", ProtoSmem::kTotal);

    // ---- numerics -----------------------------------------------------------------------------------
    proto_kernel<<<2, 512, ProtoSmem::kTotal>>>(d_sig, check_items, 1, d_p, d_g, d_tw1, d_uc, d_us, d_win, d_cyc);
    CK(cudaDeviceSynchronize());
    std::vector<float> P(static_cast<size_t>(check_items) * 8 * 1024);
    CK(cudaMemcpy(P.data(), d_p, P.size() * 4, cudaMemcpyDeviceToHost));
    const double two_pi = 6.283185307179586476925286766559;
    std::vector<double> ct(2048), st(2048);
    for (int i = 0; i < 2048; ++i) { ct[i] = std::cos(two_pi * i / 2048.0); st[i] = std::sin(two_pi * i / 2048.0); }
    for (int item = 0; item < check_items; ++item) {
        double worst_rel = 0, worst_peak = 0;
        int bad = 0;
        for (int f = 0; f < 8; ++f) {
            const float* x = sig.data() + static_cast<size_t>(item) * 4096 + f * 512;
            std::vector<double> xw(2048), ref(1024);
            for (int n = 0; n < 1024; ++n) {   // the kernel's window identity: w[n + 1024] = 1 - w[n], products rounded to fp32
                xw[n] = static_cast<double>(x[n] * window[n]);
                xw[n + 1024] = static_cast<double>(fmaf(-x[n + 1024], window[n], x[n + 1024]));
            }
            double peak = 0;
            for (int k = 1; k < 1024; ++k) {
                double re = 0, im = 0;
                for (int n = 0; n < 2048; ++n) { const int a = (n * k) & 2047; re += xw[n] * ct[a]; im -= xw[n] * st[a]; }
                ref[k] = 4.0 * (re * re + im * im);
                peak = std::max(peak, ref[k]);
            }
            for (int k = 1; k < 1024; ++k) {
                const double got = P[(static_cast<size_t>(item) * 8 + f) * 1024 + k];
                const double err = std::fabs(got - ref[k]);
                if (!(err == err)) { ++bad; continue; }
                worst_rel = std::max(worst_rel, err / std::max(ref[k], 1e-6 * peak));
                worst_peak = std::max(worst_peak, err / std::max(peak, 1e-300));
                if (err > 1e-4 * std::max(ref[k], 1e-6 * peak)) ++bad;
            }
        }
        printf("check item %d: max |P - ref| / max(ref, 1e-6 peak) = %.3e, / frame peak = %.3e, bins beyond 1e-4: %d\n", item, worst_rel, worst_peak, bad);
    }

    // ---- cycles -------------------------------------------------------------------------------------
    for (int rep = 0; rep < 6; ++rep) {
        const int mode = rep < 3 ? 0 : rep - 2;   // 1: one group only, 2: no MMAs, 3: both
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        proto_kernel<<<sms, 512, ProtoSmem::kTotal>>>(d_sig, perf_items, 0, d_p, d_g, d_tw1, d_uc, d_us, d_win, d_cyc, d_ph, mode);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        std::vector<unsigned long long> cyc(sms * 2);
        CK(cudaMemcpy(cyc.data(), d_cyc, cyc.size() * 8, cudaMemcpyDeviceToHost));
        double mx = 0, av = 0;
        for (auto c : cyc) { mx = std::max<double>(mx, c); av += c; }
        av /= cyc.size();
        const double frames_per_sm = static_cast<double>(perf_items) * 8 / sms;
        if (rep >= 2) {
            printf("   mode %d (bit 0: only group 0 works; bit 1: no MMAs)\n", mode);
            std::vector<unsigned long long> ph(24);
            CK(cudaMemcpy(ph.data(), d_ph, 24 * 8, cudaMemcpyDeviceToHost));
            const char* nm[12] = {"load+bar", "B0", "bar+issue S1", "wait S1", "B1a", "bar+issue S2A", "B1b compute", "wait S2A", "B1b store+bar+issue", "wait S2B", "B2", "fixup+bars"};
            const double tiles = perf_items / (2.0 * sms);
            for (int i = 0; i < 12; ++i) printf("   phase %-20s: %8.0f cycles per tile (warp 0)   %8.0f (warp 5)\n", nm[i], ph[i] / tiles, ph[12 + i] / tiles);
        }
        printf("perf rep %d: %d tiles, %.3f ms, %.1f cycles per frame per SM (slowest group %.1f), = %.2f M frames/s = %.2f M 157-frame clips/s for the FFT core alone\n",
               rep, perf_items, ms, av / frames_per_sm, mx / frames_per_sm, perf_items * 8 / ms / 1e3, perf_items * 8 / 157.0 / ms / 1e3);
    }
    return 0;
}
