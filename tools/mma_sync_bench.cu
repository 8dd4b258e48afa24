// Microbenchmark: throughput of the legacy warp-level tensor path (mma.sync) on
// sm_100a -- bf16 m16n8k16 and tf32 m16n8k8, fp32 accumulate -- to decide
// whether a split-precision (3 x bf16 / 3 x tf32) attention kernel pays; and
// bf16 / fp16 m16n8k8 (does a half-depth k-step cost half? head dim 40 = 2.5 k16 steps).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_sync_bench tools/mma_sync_bench.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

template <int KIND>   // 0: bf16 m16n8k16, 1: tf32 m16n8k8, 2: bf16 m16n8k8, 3: fp16 m16n8k16, 4: fp16 m16n8k8
__global__ void __launch_bounds__(256) bench(int iters, float* sink, long long* cycles) {
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    uint32_t a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 * 3, b1 = b0 + 1;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (KIND == 0) {
                asm volatile(
                    "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 "
                    "{%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                    : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
                    : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            } else if (KIND == 2) {
                asm volatile(
                    "mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 "
                    "{%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                    : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
                    : "r"(a0), "r"(a1), "r"(b0));
            } else if (KIND == 3) {
                asm volatile(
                    "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 "
                    "{%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                    : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
                    : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            } else if (KIND == 4) {
                asm volatile(
                    "mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 "
                    "{%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                    : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
                    : "r"(a0), "r"(a1), "r"(b0));
            } else {
                asm volatile(
                    "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 "
                    "{%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                    : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
                    : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 4; ++j) s += acc[i][j];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float* sink; long long* cycles;
    cudaMalloc(&sink, sms * 4 * 256 * sizeof(float));
    cudaMallocManaged(&cycles, sms * 4 * sizeof(long long));
    const int iters = 4096;
    const char* names[5] = {"bf16 m16n8k16", "tf32 m16n8k8 ", "bf16 m16n8k8 ", "fp16 m16n8k16", "fp16 m16n8k8 "};
    for (int kind = 0; kind < 5; ++kind) {
        for (int ctas_per_sm = 1; ctas_per_sm <= 4; ctas_per_sm *= 2) {
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                switch (kind) {
                    case 0: bench<0><<<sms * ctas_per_sm, 256>>>(iters, sink, cycles); break;
                    case 1: bench<1><<<sms * ctas_per_sm, 256>>>(iters, sink, cycles); break;
                    case 2: bench<2><<<sms * ctas_per_sm, 256>>>(iters, sink, cycles); break;
                    case 3: bench<3><<<sms * ctas_per_sm, 256>>>(iters, sink, cycles); break;
                    default: bench<4><<<sms * ctas_per_sm, 256>>>(iters, sink, cycles); break;
                }
                cudaEventRecord(e1);
                cudaDeviceSynchronize();
            }
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            const double flops_per_mma = (kind == 0 || kind == 3) ? 2.0 * 16 * 8 * 16 : 2.0 * 16 * 8 * 8;
            const double mmas = (double)sms * ctas_per_sm * 8 /*warps*/ * iters * 8;
            printf("%s warps/SM %2d: %.1f TFLOP/s  (%.2f clk per MMA per SM sub-partition, %lld clk)\n",
                   names[kind], ctas_per_sm * 8,
                   mmas * flops_per_mma / (ms * 1e-3) / 1e12,
                   (double)cycles[0] / (iters * 8.0 * ctas_per_sm * 2), cycles[0]);
        }
    }
    return 0;
}
