// ubench_hmma.cu -- issue rate of the legacy warp-level MMAs (mma.sync) on sm_100a: how many cycles per instruction per scheduler?
// 16 warps per SM (4 per scheduler), 8 independent accumulator chains per warp, operands in registers.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_hmma ubench_hmma.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int KIND>
__global__ void __launch_bounds__(512, 1) k(float* out, int iters) {
    float d[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) d[i][j] = threadIdx.x * 1e-3f + i + j;
    uint32_t a0 = threadIdx.x * 3 + 1, a1 = a0 * 7, a2 = a0 * 11, a3 = a0 * 13, b0 = a0 * 17, b1 = a0 * 19;
    a0 &= 0x3f7fe000u; a1 &= 0x3f7fe000u; a2 &= 0x3f7fe000u; a3 &= 0x3f7fe000u; b0 &= 0x3f7fe000u; b1 &= 0x3f7fe000u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (KIND == 1)
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3]) : "r"(a0), "r"(a1), "r"(b0));
            else if (KIND == 2)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (KIND == 3)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3]) : "r"(a0), "r"(a1), "r"(b0));
        }
    }
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) r += d[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int KIND> void run(const char* name, double macs, float* out, int sms) {
    const int iters = 4000;
    for (int threads : {512, 128}) {
        k<KIND><<<sms, threads>>>(out, 10); cudaDeviceSynchronize();
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0); k<KIND><<<sms, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double per_sched = double(iters) * 8 * (threads / 32) / 4;     // instructions per scheduler
        const double cyc = ms * 1.965e6 / per_sched;
        printf("%-28s %2d warps/SM: %6.2f cycles per instruction per scheduler = %7.1f MAC/clk/SM   %s\n", name, threads / 32, cyc, macs * 4 / cyc, cudaGetErrorString(cudaGetLastError()));
    }
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out; cudaMalloc(&out, sms * 512 * sizeof(float));
    run<0>("mma.sync m16n8k8  tf32", 1024, out, sms);
    run<1>("mma.sync m16n8k4  tf32", 512, out, sms);
    run<2>("mma.sync m16n8k16 bf16", 2048, out, sms);
    run<3>("mma.sync m16n8k16 f16", 2048, out, sms);
    run<4>("mma.sync m16n8k8  bf16", 1024, out, sms);
    return 0;
}
