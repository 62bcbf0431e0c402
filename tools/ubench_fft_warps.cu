// ubench_fft_warps.cu -- how fast do N free-running warps per SM turn staged frames into power rows?
// The frame transform of the log-mel kernel (window + 2048-point real FFT + 4|X|^2, csrc/logmel_kernel.cuh), one warp
// per frame, WITHOUT the group barriers, the mel phase, staging and normalisation around it: the upper bound for a
// warp-specialised arrangement (FFT warps | mel warps).  Prints cycles per frame per SM for 4 ... 16 warps.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ubench_fft_warps ubench_fft_warps.cu
#include <cstdio>
#include <vector>
#include <cmath>
#include "../audio_classification_icbhi_b200/csrc/logmel_kernel.cuh"
using namespace lm;

// NOISE: what the warps beyond the first `nfft_warps` run beside the FFT warps (until the FFT warps are done):
//   1 LOP3 chains (ALU pipe)   2 LDS.128 stream (shared-memory pipe)   3 HMMA.1688 TF32 chains (tensor pipe)   4 scalar FFMA chains   5 spin on an mbarrier that never completes
__device__ __forceinline__ void side_stream(int kind, volatile int* stop, const float4* smem4, uint64_t* bar, float* out) {
    const int lane = threadIdx.x & 31;
    uint32_t q[8]; float f[8]; float d[4][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { q[i] = threadIdx.x * 77 + i; f[i] = threadIdx.x * 1e-3f + i; }
#pragma unroll
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
    float4 acc = make_float4(0, 0, 0, 0);
    long long n = 0;
    while (*stop == 0) {
        ++n;
        if (kind == 1) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("lop3.b32 %0, %0, %1, 0x9e3779b9, 0x96;" : "+r"(q[i]) : "r"(q[(i + 1) & 7]));
        } else if (kind == 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float4 v;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                             : "r"(smem_u32(smem4 + lane + 32 * ((i + (int)n) & 7))) : "memory");
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        } else if (kind == 3) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 4; ++i) mma_tf32(d[i], q[0], q[1], q[2], q[3], q[4], q[5]);
        } else if (kind == 4) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(0.999f), "f"(1e-3f));
        } else {
            asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(0u) : "memory");
        }
    }
    float r = acc.x + acc.y + acc.z + acc.w;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += f[i] + __uint_as_float(q[i] & 0xff);
#pragma unroll
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r += d[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r + (float)n;
}

template <int MAXT>
__global__ void __launch_bounds__(MAXT, 1) fftk(const float* __restrict__ g_win, const float2* __restrict__ g_tw,
                                               const float2* __restrict__ g_utw, const float* __restrict__ g_samples,
                                               float* out, int iters, int nfft_warps = 99, int noise = 0) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* s_win = reinterpret_cast<float*>(smem_raw);                 // 1024
    float2* s_tw = reinterpret_cast<float2*>(s_win + 1024);            // 32 * kTwRows
    float2* s_utw = s_tw + 32 * kTwRows;                               // 512
    float* sb = reinterpret_cast<float*>(s_utw + 512);                 // 5632 staged samples (8 frames, hop 512)
    float* rows = sb + 5632;                                           // one row per warp
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 1024; i += blockDim.x) s_win[i] = g_win[i];
    for (int i = tid; i < 32 * kTwRows; i += blockDim.x) s_tw[i] = g_tw[i];
    for (int i = tid; i < 512; i += blockDim.x) s_utw[i] = g_utw[i];
    for (int i = tid; i < 5632; i += blockDim.x) sb[i] = g_samples[i];
    for (int i = tid; i < (int)(blockDim.x / 32) * kRowFloats; i += blockDim.x) rows[i] = 0.f;
    __shared__ int s_stop, s_done;
    __shared__ uint64_t s_bar;
    if (tid == 0) { s_stop = 0; s_done = 0; mbar_init(&s_bar, 1); }
    __syncthreads();
    if (warp >= nfft_warps) {
        side_stream(noise, &s_stop, reinterpret_cast<const float4*>(sb), &s_bar, out);
        return;
    }
    float* const scr = rows + warp * kRowFloats;
    const int hop = 512;
    float keep = 0.f;
    for (int it = 0; it < iters; ++it) {
        const int f = (it + warp) & 7;
        lm_f2 z[32];
        {
            const float2* __restrict__ s2 = reinterpret_cast<const float2*>(sb + f * hop);
            const float2* __restrict__ w2 = reinterpret_cast<const float2*>(s_win);
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const float2 v1 = s2[32 * r + lane];
                const float2 v2 = s2[32 * (r + 16) + lane];
                const float2 w = w2[32 * r + lane];
                const lm_f2 V1 = lm_pack(v1.x, v1.y), V2 = lm_pack(v2.x, v2.y), W = lm_pack(w.x, w.y);
                z[r] = lm_fma2(lm_sub2(V1, V2), W, V2);
                z[r + 16] = lm_fma2(lm_add2(V1, V2), W, lm_pack(-v2.x, -v2.y));
            }
            lm_fft32_aos_from2(z);
        }
        float xr[32], xi[32];
        warp_cfft1024_part2(z, xr, xi, scr, s_tw, lane);
        const int srcl = (32 - lane) & 31;
        const bool l0 = (lane == 0);
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
        {
            const float ar = xr[8], ai = xi[8], br = xr[24], bi = xi[24];
            const float er = ar + br, ei = ai - bi, orr = ai + bi, oi = br - ar;
            const float c = 0.70710678118654752440f;
            const float tr = c * (orr + oi), ti = c * (oi - orr);
            const float ur = er + tr, ui = ei + ti, vr = er - tr, vi = ei - ti;
            if (l0) { scr[256] = fmaf(ur, ur, ui * ui); scr[768] = fmaf(vr, vr, vi * vi); }
        }
        __syncwarp();
        keep += scr[(lane * 33 + it) & 1023];   // the row is "consumed" (one load) before the next frame overwrites it
        __syncwarp();
    }
    out[blockIdx.x * blockDim.x + tid] = keep;
    if (lane == 0 && atomicAdd(&s_done, 1) == min(nfft_warps, (int)(blockDim.x >> 5)) - 1) *reinterpret_cast<volatile int*>(&s_stop) = 1;
}

template <int MAXT> void run(const char* label, const float* win, const float2* tw, const float2* utw, const float* smp, float* out, int sms) {
    const int iters = 2000;
    for (int nw : {4, 8, 12, 16}) {
        if (nw * 32 > MAXT) continue;
        const size_t smem = sizeof(float) * 1024 + sizeof(float2) * (32 * kTwRows + 512) + sizeof(float) * (5632 + nw * kRowFloats);
        cudaFuncSetAttribute(fftk<MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        fftk<MAXT><<<sms, nw * 32, smem>>>(win, tw, utw, smp, out, 10);
        cudaDeviceSynchronize();
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        fftk<MAXT><<<sms, nw * 32, smem>>>(win, tw, utw, smp, out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double cyc_per_frame_sm = ms * 1.965e6 / (double(iters) * nw);
        printf("%s %2d warps per SM: %7.1f cycles per frame per SM   (%6.0f cycles per frame per warp)   %s\n", label, nw, cyc_per_frame_sm,
               cyc_per_frame_sm * nw, cudaGetErrorString(cudaGetLastError()));
    }
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    std::vector<float> win(2048), smp(5632);
    std::vector<float2> tw(32 * kTwRows), utw(512);
    const double two_pi = 6.283185307179586476925286766559;
    for (int n = 0; n < 2048; ++n) win[n] = (float)(0.5 - 0.5 * cos(two_pi * n / 2048.0));
    for (int r = 0; r < kTwRows; ++r) {
        const int k1 = (kTwRows == 31) ? r + 1 : (r < 3 ? r + 1 : 4 * (r - 2));
        for (int n2 = 0; n2 < 32; ++n2) { const double a = two_pi * (k1 * n2) / 1024.0; tw[r * 32 + n2] = make_float2((float)cos(a), (float)-sin(a)); }
    }
    for (int k = 0; k < 512; ++k) { const double a = two_pi * k / 2048.0; utw[k] = make_float2((float)cos(a), (float)sin(a)); }
    for (int i = 0; i < 5632; ++i) smp[i] = (float)sin(0.37 * i) * 0.1f + 0.01f * (float)((i * 2654435761u) >> 20) / 4096.f;
    float *d_win, *d_smp, *d_out; float2 *d_tw, *d_utw;
    cudaMalloc(&d_win, 8192); cudaMalloc(&d_smp, 5632 * 4); cudaMalloc(&d_tw, tw.size() * 8); cudaMalloc(&d_utw, 4096); cudaMalloc(&d_out, sms * 512 * 4);
    cudaMemcpy(d_win, win.data(), 8192, cudaMemcpyHostToDevice); cudaMemcpy(d_smp, smp.data(), 5632 * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_tw, tw.data(), tw.size() * 8, cudaMemcpyHostToDevice); cudaMemcpy(d_utw, utw.data(), 4096, cudaMemcpyHostToDevice);
    printf("frame transform only (window, 2048-point real FFT, 4|X|^2 -> shared-memory row), no barriers; the shipped kernel runs at ~818 cycles per frame per SM all told\n");
    run<512>("128 regs", d_win, d_tw, d_utw, d_smp, d_out, sms);
    // 8 FFT warps + 8 warps of a side stream: what does a concurrent instruction stream cost the FFT warps?
    const char* kinds[6] = {"", "LOP3 chains (ALU)", "LDS.128 stream", "HMMA.1688 TF32 chains", "scalar FFMA chains", "mbarrier.try_wait spin"};
    for (int kind = 1; kind <= 5; ++kind) {
        const int iters = 2000, nw = 16;
        const size_t smem = sizeof(float) * 1024 + sizeof(float2) * (32 * kTwRows + 512) + sizeof(float) * (5632 + nw * kRowFloats);
        cudaFuncSetAttribute(fftk<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        fftk<512><<<sms, 512, smem>>>(d_win, d_tw, d_utw, d_smp, d_out, 10, 8, kind);
        cudaDeviceSynchronize();
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        fftk<512><<<sms, 512, smem>>>(d_win, d_tw, d_utw, d_smp, d_out, iters, 8, kind);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("8 FFT warps + 8 warps of %-24s: %7.1f cycles per frame per SM   %s\n", kinds[kind], ms * 1.965e6 / (double(iters) * 8), cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
