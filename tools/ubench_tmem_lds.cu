// ubench_tmem_lds.cu -- can TMEM serve as a constant-operand cache next to a busy shared-memory pipe?
//
// The log-mel kernel reads ~150 of its ~520 shared-memory wavefronts per frame from tables that never change
// (mel A fragments, window, twiddles).  TMEM is read with tcgen05.ld, which does not go through the LSU data pipe.
// This measures, with 16 warps per SM as in the kernel (cycles per iteration per SM):
//   L   8 LDS.128 per warp                       (512 wavefronts)
//   T   8 tcgen05.ld.32x32b.x8 + 1 wait per warp  (the same 64 registers per lane)
//   F   32 FFMA2 per warp                        (256 FMA-pipe cycles per scheduler -> 256)
// alone and combined, plus the single-warp latency of a dependent tcgen05.ld -> wait chain.
// (tools/ubench_overlap.cu's LDS numbers are wrong: its loads had no memory clobber and ptxas hoisted them
// out of the unrolled loop -- this file replaces them.)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_tmem_lds ubench_tmem_lds.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void ldtm8(uint32_t (&r)[8], uint32_t taddr) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void sttm8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void ldtm_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void sttm_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// MODE bit 0: F, bit 1: L, bit 2: T
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, int iters, float s, int check, int one) {
    extern __shared__ __align__(16) float4 sm[];   // 64 KB: 8 distinct LDS.128 per warp
    __shared__ uint32_t tbase_s;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = make_float4(i, -i, 2 * i, 3 * i);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(smem_u32(&tbase_s), 512);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tbase_s;
    const uint32_t tq = tbase + ((static_cast<uint32_t>(warp & 3) * 32u) << 16);   // this warp's lane quadrant
    if (warp < 4) {   // fill columns 0..63 of every lane: word = lane * 1000 + column
        for (int c = 0; c < 128; c += 8) {
            uint32_t v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (warp * 32 + lane) * 1000 + c + j;
            sttm8(tq + c, v);
        }
        sttm_wait();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    u64 a[16]; u64 sv, cv;
    asm("mov.b64 %0, {%1,%1};" : "=l"(sv) : "f"(s));
    asm("mov.b64 %0, {%1,%1};" : "=l"(cv) : "f"(1e-3f));
#pragma unroll
    for (int i = 0; i < 16; ++i) { float v = threadIdx.x * 1e-3f + i; asm("mov.b64 %0, {%1,%1};" : "=l"(a[i]) : "f"(v)); }
    unsigned sink = 0;
    int bad = 0;
    const uint32_t lbase = smem_u32(sm + warp * 256 + lane);
    for (int it = 0; it < iters; ++it) {
        uint32_t t[8][8];
        if (MODE & 4) {
#pragma unroll
            for (int i = 0; i < 8; ++i) ldtm8(t[i], tq + 8 * i + (((it * one) & 1) << 6));   // iteration-dependent: identical asm statements of an unrolled loop get merged
        }
        float4 q[8];
        if (MODE & 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(q[i].x), "=f"(q[i].y), "=f"(q[i].z), "=f"(q[i].w)
                             : "r"(lbase + 512u * i + (((it * one) & 1) << 4)) : "memory");
        }
        if (MODE & 1) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = fma2(a[i], sv, cv);
        }
        if (MODE & 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) sink ^= __float_as_uint(q[i].x) ^ __float_as_uint(q[i].y) ^ __float_as_uint(q[i].z) ^ __float_as_uint(q[i].w);
        }
        if (MODE & 4) {
            ldtm_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    sink ^= t[i][j];
                    if (check && t[i][j] != static_cast<uint32_t>(((warp & 3) * 32 + lane) * 1000 + 8 * i + j + (((it * one) & 1) << 6))) ++bad;
                }
        }
    }
    float r = __uint_as_float(sink & 0xff) + bad * 1e6f;
#pragma unroll
    for (int i = 0; i < 16; ++i) { float x, y; asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(a[i])); r += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

// one warp, dependent chain: address of the next load comes from the previous one's data (always 0 offset)
__global__ void __launch_bounds__(128, 1) lat(long long* out, int iters) {
    __shared__ uint32_t tbase_s;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(smem_u32(&tbase_s), 512);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tbase_s;
    if (warp == 0) {
        uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int c = 0; c < 64; c += 8) sttm8(tbase + c, z);
        sttm_wait();
        uint32_t off = 0;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            uint32_t t[8];
            ldtm8(t, tbase + off);
            ldtm_wait();
            off = t[0] & 8u;
        }
        const long long t1 = clock64();
        if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = off; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

template <class F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
template <int M> void optin() { cudaFuncSetAttribute(k<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 64); }
int main() {
    optin<1>(); optin<2>(); optin<3>(); optin<4>(); optin<5>(); optin<6>(); optin<7>();
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out; cudaMalloc(&out, sms * 512 * sizeof(float));
    const int iters = 5000;
    const double cyc = 1.965e6 / iters;
    // correctness of the TMEM fill / read first
    k<4><<<sms, 512, 65536 + 64>>>(out, 4, 0.999f, 1, 1);
    float h[512]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    int bad = 0; for (int i = 0; i < 512; ++i) if (h[i] > 1e5f) ++bad;
    printf("TMEM table read-back: %s (%s)\n", bad ? "MISMATCH" : "ok", cudaGetErrorString(cudaGetLastError()));
    const char* names[8] = {"", "F   ", "L   ", "F+L ", "T   ", "F+T ", "L+T ", "F+L+T"};
    float t[8] = {0};
    t[1] = timeit([&] { k<1><<<sms, 512, 65536 + 64>>>(out, iters, 0.999f, 0, 1); });
    t[2] = timeit([&] { k<2><<<sms, 512, 65536 + 64>>>(out, iters, 0.999f, 0, 1); });
    t[3] = timeit([&] { k<3><<<sms, 512, 65536 + 64>>>(out, iters, 0.999f, 0, 1); });
    t[4] = timeit([&] { k<4><<<sms, 512, 65536 + 64>>>(out, iters, 0.999f, 0, 1); });
    t[5] = timeit([&] { k<5><<<sms, 512, 65536 + 64>>>(out, iters, 0.999f, 0, 1); });
    t[6] = timeit([&] { k<6><<<sms, 512, 65536 + 64>>>(out, iters, 0.999f, 0, 1); });
    t[7] = timeit([&] { k<7><<<sms, 512, 65536 + 64>>>(out, iters, 0.999f, 0, 1); });
    printf("per SM and iteration, 16 warps: F = 32 FFMA2 per warp (floor 256), L = 8 LDS.128 per warp (512 wavefronts), T = 8 tcgen05.ld.x8 per warp (the same 64 registers)\n");
    for (int m = 1; m < 8; ++m) printf("%s %8.3f ms  %7.1f cycles/iter\n", names[m], t[m], t[m] * cyc);
    long long* lo; cudaMalloc(&lo, 16);
    lat<<<1, 128>>>(lo, 10000); cudaDeviceSynchronize();
    long long hl[2]; cudaMemcpy(hl, lo, 16, cudaMemcpyDeviceToHost);
    printf("tcgen05.ld.x8 -> wait -> dependent address: %.1f cycles per round trip (%s)\n", hl[0] / 10000.0, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
