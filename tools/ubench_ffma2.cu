// Microbenchmark: issue/throughput of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_ffma2 ubench_ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

template <int ILP>
__global__ void k_scalar(float* out, int iters, float s) {
    float a[ILP];
    for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) a[i] = fma1(a[i], s, 1e-3f);
    }
    float r = 0; for (int i = 0; i < ILP; ++i) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int ILP>
__global__ void k_packed(float* out, int iters, float s) {
    u64 a[ILP]; u64 sv, cv;
    asm("mov.b64 %0, {%1,%1};" : "=l"(sv) : "f"(s));
    asm("mov.b64 %0, {%1,%1};" : "=l"(cv) : "f"(1e-3f));
    for (int i = 0; i < ILP; ++i) { float v = threadIdx.x * 1e-3f + i; asm("mov.b64 %0, {%1,%1};" : "=l"(a[i]) : "f"(v)); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) a[i] = fma2(a[i], sv, cv);
    }
    float r = 0; for (int i = 0; i < ILP; ++i) { float x, y; asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(a[i])); r += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int ILP>   // packed FFMA2 interleaved with an equal number of shared-memory loads
__global__ void k_mixed(float* out, int iters, float s) {
    __shared__ float sm[1024];
    sm[threadIdx.x] = threadIdx.x; __syncthreads();
    u64 a[ILP]; u64 sv, cv; float acc = 0;
    asm("mov.b64 %0, {%1,%1};" : "=l"(sv) : "f"(s));
    asm("mov.b64 %0, {%1,%1};" : "=l"(cv) : "f"(1e-3f));
    for (int i = 0; i < ILP; ++i) { float v = threadIdx.x * 1e-3f + i; asm("mov.b64 %0, {%1,%1};" : "=l"(a[i]) : "f"(v)); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) { a[i] = fma2(a[i], sv, cv); acc += sm[(threadIdx.x + i * 32 + it) & 1023]; }
    }
    float r = acc; for (int i = 0; i < ILP; ++i) { float x, y; asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(a[i])); r += x + y; }
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
    float* out; cudaMalloc(&out, sms * 1024 * sizeof(float));
    const int iters = 20000; const int ILP = 8;
    for (int threads : {128, 256, 512, 1024}) {
        float ms1 = timeit([&] { k_scalar<ILP><<<sms, threads>>>(out, iters, 0.999f); });
        float ms2 = timeit([&] { k_packed<ILP><<<sms, threads>>>(out, iters, 0.999f); });
        float ms3 = timeit([&] { k_mixed<ILP><<<sms, threads>>>(out, iters, 0.999f); });
        double n = (double)iters * ILP * threads;   // thread-instructions per SM
        printf("threads/SM %4d: FFMA %.3f ms (%.1f thread-inst/ns/SM)  FFMA2 %.3f ms (%.1f packed-inst/ns/SM = %.1f fma/ns/SM)  FFMA2+LDS %.3f ms\n",
               threads, ms1, n / (ms1 * 1e6), ms2, n / (ms2 * 1e6), 2 * n / (ms2 * 1e6), ms3);
    }
    printf("(at 1.965 GHz, 128 thread-inst/clk/SM = 251.5 thread-inst/ns/SM)\n");
    return 0;
}
