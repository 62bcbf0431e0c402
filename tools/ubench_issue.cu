// ubench_issue.cu -- does a packed FFMA2 take one issue slot or two?  (What bounds an instruction mix that is ~45 % packed fp32.)
// 16 warps per SM (4 per scheduler).  Per warp and iteration:
//   F  32 FFMA2 (16 independent chains x 2)       FMA-pipe floor 2 cycles each per scheduler
//   S  64 scalar FFMA (32 chains x 2)             the same flops
//   I  64 LOP3 (16 independent chains x 4)        ALU pipe
//   M  32 shared-memory loads LDS.64 (conflict-free), addresses change per iteration
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_issue ubench_issue.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ uint32_t lop(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }

template <int MODE>   // bit 0 F, bit 1 S, bit 2 I, bit 3 M
__global__ void __launch_bounds__(512, 1) k(float* out, int iters, float s, int one) {
    extern __shared__ __align__(16) float2 sm[];   // 64 KB + slack
    for (int i = threadIdx.x; i < 8192 + 64; i += blockDim.x) sm[i] = make_float2(i, -i);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 a[16]; u64 sv, cv;
    asm("mov.b64 %0, {%1,%1};" : "=l"(sv) : "f"(s));
    asm("mov.b64 %0, {%1,%1};" : "=l"(cv) : "f"(1e-3f));
    float f[32];
    uint32_t q[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { float v = threadIdx.x * 1e-3f + i; asm("mov.b64 %0, {%1,%1};" : "=l"(a[i]) : "f"(v)); q[i] = threadIdx.x * 77 + i; }
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = threadIdx.x * 1e-3f + i;
    unsigned sink = 0;
    const uint32_t lbase = static_cast<uint32_t>(__cvta_generic_to_shared(sm + warp * 512 + lane));
    for (int it = 0; it < iters; ++it) {
        float2 v[32];
        if (MODE & 8) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v[i].x), "=f"(v[i].y) : "r"(lbase + 256u * (i & 15) + (((it * one + (i >> 4)) & 7) << 3)) : "memory");
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            if (MODE & 1) {
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = fma2(a[i], sv, cv);
            }
            if (MODE & 2) {
#pragma unroll
                for (int i = 0; i < 32; ++i) f[i] = fma1(f[i], s, 1e-3f);
            }
            if (MODE & 4) {
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int i = 0; i < 16; ++i) q[i] = lop(q[i], q[(i + 1) & 15], 0x9e3779b9u);
            }
        }
        if (MODE & 8) {
#pragma unroll
            for (int i = 0; i < 32; ++i) sink ^= __float_as_uint(v[i].x) ^ __float_as_uint(v[i].y);
        }
    }
    float r = __uint_as_float(sink & 0xff);
#pragma unroll
    for (int i = 0; i < 16; ++i) { float x, y; asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(a[i])); r += x + y + __uint_as_float(q[i] & 0xff); }
#pragma unroll
    for (int i = 0; i < 32; ++i) r += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <class F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
template <int M> float run(float* out, int sms, int iters, int threads) {
    cudaFuncSetAttribute(k<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 512);
    return timeit([&] { k<M><<<sms, threads, 65536 + 512>>>(out, iters, 0.999f, 1); });
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out; cudaMalloc(&out, sms * 512 * sizeof(float));
    const int iters = 4000;
    const double cyc = 1.965e6 / iters;
    for (int threads : {512, 256}) {
        printf("%d warps per SM; cycles per iteration per SM.  F = 32 FFMA2, S = 64 FFMA, I = 64 LOP3, M = 32 LDS.64 (per warp)\n", threads / 32);
        printf("  F     %7.1f\n", run<1>(out, sms, iters, threads) * cyc);
        printf("  S     %7.1f\n", run<2>(out, sms, iters, threads) * cyc);
        printf("  I     %7.1f\n", run<4>(out, sms, iters, threads) * cyc);
        printf("  M     %7.1f\n", run<8>(out, sms, iters, threads) * cyc);
        printf("  F+I   %7.1f\n", run<5>(out, sms, iters, threads) * cyc);
        printf("  S+I   %7.1f\n", run<6>(out, sms, iters, threads) * cyc);
        printf("  F+M   %7.1f\n", run<9>(out, sms, iters, threads) * cyc);
        printf("  S+M   %7.1f\n", run<10>(out, sms, iters, threads) * cyc);
        printf("  I+M   %7.1f\n", run<12>(out, sms, iters, threads) * cyc);
        printf("  F+I+M %7.1f\n", run<13>(out, sms, iters, threads) * cyc);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
