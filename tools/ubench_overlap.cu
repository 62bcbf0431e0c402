// Microbenchmark: do packed FP32 math (FFMA2) and shared-memory traffic (LDS.64 / STS.32) overlap on one SM?
// Independent instruction streams, 16 warps per SM (4 per scheduler), as in the log-mel kernel.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_overlap ubench_overlap.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

// MODE bit 0: 16 FFMA2 per iteration; bit 1: 8 LDS.64 per iteration; bit 2: 8 STS.32 per iteration;
//      bit 3: 8 LDS.128 per iteration; bit 4: 32 scalar FFMA per iteration (the flops of 16 FFMA2)
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, int iters, float s) {
    __shared__ float2 sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = make_float2(i, -i);
    __syncthreads();
    u64 a[16]; u64 sv, cv;
    asm("mov.b64 %0, {%1,%1};" : "=l"(sv) : "f"(s));
    asm("mov.b64 %0, {%1,%1};" : "=l"(cv) : "f"(1e-3f));
#pragma unroll
    for (int i = 0; i < 16; ++i) { float v = threadIdx.x * 1e-3f + i; asm("mov.b64 %0, {%1,%1};" : "=l"(a[i]) : "f"(v)); }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float2* base = sm + warp * 256 + lane;          // each warp its own 2 KB, lanes contiguous: 2 wavefronts per LDS.64
    float* sbase = reinterpret_cast<float*>(sm) + warp * 512 + lane;
    unsigned sink = 0;
    for (int it = 0; it < iters; ++it) {
        float2 v[8];
        if (MODE & 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                u64 t;
                asm volatile("ld.shared.b64 %0, [%1];" : "=l"(t) : "r"(static_cast<unsigned>(__cvta_generic_to_shared(base + 32 * i))));
                asm("mov.b64 {%0,%1}, %2;" : "=f"(v[i].x), "=f"(v[i].y) : "l"(t));
            }
        }
        float4 q[8];
        if (MODE & 8) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(q[i].x), "=f"(q[i].y), "=f"(q[i].z), "=f"(q[i].w)
                             : "r"(static_cast<unsigned>(__cvta_generic_to_shared(reinterpret_cast<const float4*>(sm) + warp * 128 + ((lane + 32 * i) & 127)))));
        }
        if (MODE & 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fma2(a[i], sv, cv);
        }
        if (MODE & 16) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float x, y;
                asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(a[i]));
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x) : "f"(s), "f"(1e-3f));
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(y) : "f"(s), "f"(1e-3f));
                asm("mov.b64 %0, {%1,%2};" : "=l"(a[i]) : "f"(x), "f"(y));
            }
        }
        if (MODE & 8) {
#pragma unroll
            for (int i = 0; i < 8; ++i) sink ^= __float_as_uint(q[i].x) ^ __float_as_uint(q[i].y) ^ __float_as_uint(q[i].z) ^ __float_as_uint(q[i].w);
        }
        if (MODE & 4) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(sbase + 32 * i))), "r"(sink + i) : "memory");
        }
        if (MODE & 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) sink ^= __float_as_uint(v[i].x) ^ __float_as_uint(v[i].y);
        }
    }
    float r = __uint_as_float(sink & 0xff);
#pragma unroll
    for (int i = 0; i < 16; ++i) { float x, y; asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(a[i])); r += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <class F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out; cudaMalloc(&out, sms * 512 * sizeof(float));
    const int iters = 20000;
    float t1 = timeit([&] { k<1><<<sms, 512>>>(out, iters, 0.999f); });
    float t2 = timeit([&] { k<2><<<sms, 512>>>(out, iters, 0.999f); });
    float t4 = timeit([&] { k<4><<<sms, 512>>>(out, iters, 0.999f); });
    float t3 = timeit([&] { k<3><<<sms, 512>>>(out, iters, 0.999f); });
    float t5 = timeit([&] { k<5><<<sms, 512>>>(out, iters, 0.999f); });
    float t7 = timeit([&] { k<7><<<sms, 512>>>(out, iters, 0.999f); });
    // per SM and iteration: 16 warps x 16 FFMA2 x 2 cycles / 4 schedulers = 128 FMA-pipe cycles; 16 x 8 LDS.64 = 256 wavefronts; 16 x 8 STS.32 = 128 wavefronts
    float t8 = timeit([&] { k<8><<<sms, 512>>>(out, iters, 0.999f); });
    float t9 = timeit([&] { k<9><<<sms, 512>>>(out, iters, 0.999f); });
    float t16 = timeit([&] { k<16><<<sms, 512>>>(out, iters, 0.999f); });
    float t18 = timeit([&] { k<18><<<sms, 512>>>(out, iters, 0.999f); });
    printf("LDS.128 only    %.3f ms  (%.1f cycles/iter; 128 LDS.128 = 512 x 128 B per SM and iteration)\n", t8, t8 * 1.965e6 / iters);
    printf("FFMA2 + LDS.128 %.3f ms  (%.1f cycles/iter; sum %.1f)\n", t9, t9 * 1.965e6 / iters, (t1 + t8) * 1.965e6 / iters);
    printf("FFMA x2 only    %.3f ms  (%.1f cycles/iter; FMA-pipe floor 128)\n", t16, t16 * 1.965e6 / iters);
    printf("FFMA x2 + LDS64 %.3f ms  (%.1f cycles/iter; sum %.1f)\n", t18, t18 * 1.965e6 / iters, (t16 + t2) * 1.965e6 / iters);
    printf("FFMA2 only      %.3f ms  (%.1f cycles/iter at 1.965 GHz; FMA-pipe floor 128)\n", t1, t1 * 1.965e6 / iters);
    printf("LDS.64 only     %.3f ms  (%.1f cycles/iter; wavefront floor 256)\n", t2, t2 * 1.965e6 / iters);
    printf("STS.32 only     %.3f ms  (%.1f cycles/iter; wavefront floor 128)\n", t4, t4 * 1.965e6 / iters);
    printf("FFMA2 + LDS     %.3f ms  (%.1f cycles/iter; max 256, sum %.1f)\n", t3, t3 * 1.965e6 / iters, (t1 + t2) * 1.965e6 / iters);
    printf("FFMA2 + STS     %.3f ms  (%.1f cycles/iter; max 128, sum %.1f)\n", t5, t5 * 1.965e6 / iters, (t1 + t4) * 1.965e6 / iters);
    printf("FFMA2 + LDS+STS %.3f ms  (%.1f cycles/iter; max 384, sum %.1f)\n", t7, t7 * 1.965e6 / iters, (t1 + t2 + t4) * 1.965e6 / iters);
    return 0;
}
